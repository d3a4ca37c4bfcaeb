"""Diagnosis of the ray-geometry gradients: per field relative L2 error against oracle autograd, for all
levels and for the coarse levels only (where a one-ulp difference of a coordinate cannot change the cell)."""
import sys
import torch
sys.path.insert(0, '.')
from oracle import zipnerf_oracle as zo
from nerf_lidar_b200 import synthetic, ops
from tests.test_gpu_encode import _model, _setup

GEOM = ('origins', 'directions', 'base_x', 'base_y')
sd = synthetic.init_state_dict(seed=12, table_std=0.5)
model = _model(sd)
for which, S in (('nerf', 32), ('prop0', 64), ('prop1', 64)):
    pre = 'nerf_mlp.' if which == 'nerf' else f'prop_mlp_{which[-1]}.'
    enc = model.get_submodule(pre[:-1]).encoder
    L, C = enc.num_levels, enc.level_dim
    batch, t, deg = _setup(7, S, True)
    N = t.shape[0]
    for lv_max in (L, 6, 3):
        g = torch.randn(N * S, L, C, generator=torch.Generator().manual_seed(3))
        g[:, lv_max:] = 0
        g = g.reshape(N * S, L * C)
        leaf = {k: batch[k].clone().requires_grad_(True) for k in GEOM}
        means, stds = zo.cast_rays(t, leaf['origins'], leaf['directions'], batch['radii'], leaf['base_x'], leaf['base_y'], deg)
        feat = zo.encode_features(means, stds, sd[pre + 'encoder.embeddings'], sd[pre + 'encoder.offsets'], sd[pre + 'encoder.grid_sizes'], C)
        feat.backward(g.reshape(N, S, -1))
        cu = {k: batch[k].cuda().requires_grad_(True) for k in GEOM}
        rays = ops.RayBundle({**{k: v.cuda() for k, v in batch.items()}, **cu})
        ops.nerf_encode(t.cuda(), deg.cuda(), enc, rays, 0.35).backward(g.cuda())
        errs = {k: float((cu[k].grad.cpu().double() - leaf[k].grad.double()).norm() / leaf[k].grad.double().norm()) for k in GEOM}
        print(which, 'levels <', lv_max, {k: f'{v:.2e}' for k, v in errs.items()}, flush=True)
