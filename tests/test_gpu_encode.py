"""Fused cast_rays + contract + hash-grid + erf-mean kernels (NeRF level and
proposal levels) against the oracle chain."""
import numpy as np
import pytest
import torch

from oracle import zipnerf_oracle as zo
from nerf_lidar_b200 import synthetic
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def _setup(seed, S, rand):
    batch = synthetic.to_torch(synthetic.make_train_batch(256, seed=seed))
    N = batch['origins'].shape[0]
    g = torch.Generator().manual_seed(seed)
    s = torch.sort(torch.rand(N, S + 1, generator=g), -1).values
    t = zo.s_to_t(s, batch['near'], batch['far'])
    deg = torch.rand(N, S, 7, generator=g) if rand else None
    return batch, t, deg


@pytest.mark.parametrize('rand', [False, True])
def test_nerf_encode_forward_backward(rand, full_state_dict_visible):
    from nerf_lidar_b200 import models, configs, ops
    sd = full_state_dict_visible
    batch, t, deg = _setup(5, 32, rand)
    means, stds = zo.cast_rays(t, batch['origins'], batch['directions'], batch['radii'], batch['base_x'],
                               batch['base_y'], deg)
    emb = sd['nerf_mlp.encoder.embeddings']
    want = zo.encode_features(means, stds, emb, sd['nerf_mlp.encoder.offsets'], sd['nerf_mlp.encoder.grid_sizes'], 4)
    cfg = configs.nuscenes_single()
    model = models.Model(cfg).cuda()
    model.load_state_dict(sd, strict=False)
    rays = ops.RayBundle({k: v.cuda() for k, v in batch.items()})
    feat = ops.nerf_encode(t.cuda(), None if deg is None else deg.cuda(), model.nerf_mlp.encoder, rays, 0.35)
    N, S = t.shape[0], 32
    # points within 1 ulp of a cell boundary may land in the neighbouring cell (the
    # interpolant is continuous), so compare values, not cells
    assert_close(feat.reshape(N, S, 40), want, 2e-5, 'nerf features')
    # backward: linear in the table -> <g, F(e)> = <dF^T g, e>
    g = torch.randn(N * S, 40, generator=torch.Generator().manual_seed(1)).cuda()
    feat.backward(g)
    ge = model.nerf_mlp.encoder.embeddings.grad
    lhs = float((g.double() * feat.detach().double()).sum())
    rhs = float((ge.double() * model.nerf_mlp.encoder.embeddings.detach().double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * float((g.double() * feat.detach().double()).abs().sum())


@pytest.mark.parametrize('lvl,L', [(0, 6), (1, 8)])
def test_prop_level_forward_backward(lvl, L, full_state_dict_visible):
    from nerf_lidar_b200 import models, configs, ops
    sd = {k: v.clone() for k, v in full_state_dict_visible.items()}
    pre = f'prop_mlp_{lvl}.'
    batch, t, deg = _setup(6 + lvl, 64, True)
    N, S = t.shape[0], 64
    emb = sd[pre + 'encoder.embeddings'].clone().requires_grad_(True)
    p = dict(sd)
    p[pre + 'encoder.embeddings'] = emb
    for k in list(p):
        if k.startswith(pre + 'density_layer'):
            p[k] = p[k].clone().requires_grad_(True)
    means, stds = zo.cast_rays(t, batch['origins'], batch['directions'], batch['radii'], batch['base_x'],
                               batch['base_y'], deg)

    # differentiable oracle features: gather with autograd through the table
    from oracle import grid_oracle as go
    z, s2 = zo.contract_mean_std(means.reshape(-1, 3), stds.reshape(-1))
    x01 = (z / 2 + 1) / 2
    offs = sd[pre + 'encoder.offsets'].numpy().astype(np.int64)
    feats = []
    for l in range(L):
        idx, w, valid, *_ = go.corner_setup(x01, l, 1.0, 16, offs)
        f = (emb[idx + int(offs[l]), 0] * w).sum(-1)
        feats.append(torch.where(valid, f, torch.zeros_like(f)))
    feat = torch.stack(feats, -1).reshape(N, S, 7, L, 1)
    sdv = (s2 / 2).reshape(N, S, 7)
    wj = torch.erf(1 / torch.clamp(torch.sqrt(8 * sdv[..., None] ** 2 * sd[pre + 'encoder.grid_sizes'] ** 2), min=1e-10))
    feat = (feat * wj[..., None]).mean(-3).flatten(-2, -1)
    dens_want = zo.prop_mlp(p, pre, feat)
    gd = torch.randn(N, S, generator=torch.Generator().manual_seed(2))
    dens_want.backward(gd)

    cfg = configs.nuscenes_single()
    model = models.Model(cfg).cuda()
    model.load_state_dict(sd, strict=False)
    mlp = model.get_submodule(f'prop_mlp_{lvl}')
    rays = ops.RayBundle({k: v.cuda() for k, v in batch.items()})
    dens = ops.prop_level(t.cuda(), deg.cuda(), mlp, rays, 0.35)
    assert_close(dens, dens_want, 2e-5, 'prop density')
    dens.backward(gd.cuda())
    assert_close(mlp.density_layer[0].weight.grad, p[pre + 'density_layer.0.weight'].grad, 2e-4, 'gW0')
    assert_close(mlp.density_layer[0].bias.grad, p[pre + 'density_layer.0.bias'].grad, 2e-4, 'gb0')
    assert_close(mlp.density_layer[2].weight.grad, p[pre + 'density_layer.2.weight'].grad, 2e-4, 'gW1')
    assert_close(mlp.density_layer[2].bias.grad, p[pre + 'density_layer.2.bias'].grad, 2e-4, 'gb1')
    assert_close(mlp.encoder.embeddings.grad, emb.grad, 2e-4, 'grad table')
