"""Whole-step ray-geometry gradients against oracle autograd: fused bf16 NerfMLP vs torch fp32 heads."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from oracle import train_oracle as to
from oracle import zipnerf_oracle as zo
from nerf_lidar_b200 import synthetic, configs, models, train
from tests.helpers import use_torch_heads

GEOM = ('origins', 'directions', 'base_x', 'base_y', 'viewdirs')
for table_std in (0.2, 0.02):
    sd = synthetic.init_state_dict(seed=23, table_std=table_std)
    batch = synthetic.to_torch(synthetic.make_train_batch(512, seed=23))
    n = batch['origins'].shape[0]
    rin = [{k: torch.from_numpy(v) for k, v in r.items()} for r in synthetic.make_rand_inputs(n, seed=23)]
    step, num_patch = 600, 0
    ref = to.RefTrainer(sd)
    train_frac = float(np.clip((step - 1) / (25000 - 1), 0, 1))
    leaf = {k: batch[k].clone().requires_grad_(True) for k in GEOM}
    rend, hist = zo.model_forward(ref.p, {**batch, **leaf}, rin, train_frac, True, training=False)
    ls_ref = to.losses(batch, rend, hist, step, num_patch)
    sum(ls_ref.values()).backward()
    for heads in ('torch32', 'fused'):
        cfg = configs.nuscenes_single()
        model = models.Model(cfg, training=True).cuda()
        model.load_state_dict(sd, strict=False)
        if heads == 'torch32':
            use_torch_heads(model, torch.float32)
        tr = train.Trainer(model, cfg)
        cb = {k: v.cuda() for k, v in batch.items()}
        cu = {k: cb[k].clone().requires_grad_(True) for k in GEOM}
        crin = [{k: v.cuda() for k, v in r.items()} for r in rin]
        r2, h2 = model(True, {**cb, **cu}, train_frac, True, rand_inputs=crin)
        ls = train.compute_losses(cb, r2, h2, cfg, step, num_patch)
        sum(ls.values()).backward()
        errs = {}
        for k in GEOM:
            got, want = cu[k].grad.cpu().double(), leaf[k].grad.double()
            errs[k] = f'{float((got - want).norm() / want.norm()):.2e} (|want| {float(want.norm()):.2e})'
        print(table_std, heads, errs, flush=True)
