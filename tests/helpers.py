"""Shared helpers of the parity tests."""
import os

import numpy as np
import torch

from nerf_lidar_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

CASES = {
    'eval_init': dict(batch_size=64, seed=11, table_std=1e-4, rand=False, train_frac=1.0),
    'train_visible': dict(batch_size=64, seed=12, table_std=0.5, rand=True, train_frac=0.5),
}


def load_case(name, device='cpu'):
    case = CASES[name]
    golden = dict(np.load(os.path.join(GOLDEN_DIR, f'{name}.npz')))
    sd = synthetic.init_state_dict(seed=case['seed'], table_std=case['table_std'])
    batch = synthetic.to_torch(synthetic.make_train_batch(case['batch_size'], seed=case['seed']))
    n = batch['origins'].shape[0]
    rin = None
    if case['rand']:
        rin = [{k: torch.from_numpy(v) for k, v in r.items()} for r in synthetic.make_rand_inputs(n, seed=case['seed'])]
    if device != 'cpu':
        sd = {k: v.to(device) for k, v in sd.items()}
        batch = {k: v.to(device) for k, v in batch.items()}
        if rin is not None:
            rin = [{k: v.to(device) for k, v in r.items()} for r in rin]
    return case, golden, sd, batch, rin


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def assert_close(a, b, rtol, name='', atol=0.0):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    assert a.shape == b.shape, f'{name}: shape {a.shape} vs {b.shape}'
    scale = np.abs(b).max() + 1e-30
    err = np.abs(a.astype(np.float64) - b.astype(np.float64)).max()
    assert err <= rtol * scale + atol, f'{name}: max|d|={err:.3e} scale={scale:.3e} (rtol {rtol})'
    return err / scale


import contextlib


@contextlib.contextmanager
def poisoned_empty(enable=True):
    """torch.empty / torch.empty_like return NaN-filled (floating) or 0x7f-filled (integer) tensors: a kernel that
    reads a byte it (or a predecessor) did not write shows up as a NaN / a changed result."""
    if not enable:
        yield
        return
    real_empty, real_like = torch.empty, torch.empty_like

    def fill(t):
        if t.is_floating_point():
            t.fill_(float('nan'))
        elif t.dtype == torch.uint8:
            t.fill_(0x7f)
        elif t.dtype != torch.bool:
            t.fill_(0x7f7f7f7f if t.dtype in (torch.int32, torch.int64) else 0x7f)
        return t

    def empty(*a, **k):
        return fill(real_empty(*a, **k))

    def empty_like(*a, **k):
        return fill(real_like(*a, **k))
    torch.empty, torch.empty_like = empty, empty_like
    try:
        yield
    finally:
        torch.empty, torch.empty_like = real_empty, real_like


def torch_heads(mlp, feat, viewdirs, S, dtype=torch.float32):
    """Plain-torch evaluation of the NerfMLP dense head (Z/internal/models.py:996-997,1116-1251) on the
    module's parameters -- the reference the fused tcgen05 kernels are compared with.  dtype=float32 is the
    reference's arithmetic; bfloat16 rounds the operands like the kernels do (fp32 accumulation)."""
    import math
    import torch.nn.functional as F
    N = viewdirs.shape[0]

    def lin(layer, x):
        if dtype == torch.float32:
            return F.linear(x, layer.weight, layer.bias)
        return F.linear(x.to(dtype), layer.weight.to(dtype), None).float() + layer.bias

    x = lin(mlp.density_layer[2], torch.relu(lin(mlp.density_layer[0], feat)))
    density = F.softplus(x[..., 0] + mlp.density_bias).reshape(N, S)
    sem = torch.softmax(lin(mlp.sem_layer[2], torch.relu(lin(mlp.sem_layer[0], x))), dim=-1).reshape(N, S, mlp.class_num)
    inten = lin(mlp.intensity_layer[2], torch.relu(lin(mlp.intensity_layer[0], x))).reshape(N, S, 1)
    de = mlp.dir_enc(viewdirs)
    de = de[:, None, :].expand(N, S, de.shape[-1]).reshape(N * S, -1)
    h_in = torch.cat([x, de], dim=-1)
    h = h_in
    for i in range(mlp.net_depth_viewdirs):
        h = torch.relu(lin(mlp.get_submodule(f'lin_second_stage_{i}'), h))
        if i == mlp.skip_layer_dir:
            h = torch.cat([h, h_in], dim=-1)
    rgb = torch.sigmoid(mlp.rgb_premultiplier * lin(mlp.rgb_layer, h) + mlp.rgb_bias)
    rgb = (rgb * (1 + 2 * mlp.rgb_padding) - mlp.rgb_padding).reshape(N, S, 3)
    return dict(density=density, rgb=rgb, semantic=sem, intensity=inten)


def use_torch_heads(model, dtype=torch.float32):
    """Routes the model's NerfMLP head through `torch_heads` (tests that need the reference's fp32 arithmetic
    around the other kernels: teacher-forced stages, autograd comparisons of the whole step)."""
    mlp = model.nerf_mlp
    mlp.heads = lambda feat, viewdirs, S: torch_heads(mlp, feat, viewdirs, S, dtype)
    return model
