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
