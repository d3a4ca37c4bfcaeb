"""b1 boundary: the `_gridencoder` ABI on the B200 kernels against the oracle, on
the three tables of nuscenes_single.gin.  Indices bit-exact, fp32 values 1e-5."""
import numpy as np
import pytest
import torch

from oracle import grid_oracle as go
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu

TABLES = {'prop0': (6, 1), 'prop1': (8, 1), 'nerf': (10, 4)}


def _points(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, generator=g)
    # edge cases: exact 0 / 1 / cell boundaries / out of range
    x[0] = torch.tensor([0., 0., 0.])
    x[1] = torch.tensor([1., 1., 1.])
    x[2] = torch.tensor([0.5, 0.25, 0.75])
    x[3] = torch.tensor([1.0000001, 0.5, 0.5])
    x[4] = torch.tensor([-1e-7, 0.5, 0.5])
    x[5] = torch.tensor([(7 - 0.5) / 15, (3 - 0.5) / 31, (100 - 0.5) / 8191])
    return x


@pytest.mark.parametrize('name', list(TABLES))
def test_corner_indices_bit_exact(name):
    from nerf_lidar_b200 import _lib
    L, C = TABLES[name]
    offs, _ = go.make_offsets(3, L, 2.0, 16, 21)
    x = _points(4096, 1)
    xd, od = x.cuda(), torch.from_numpy(offs).cuda()
    idx = torch.empty(L, x.shape[0], 8, dtype=torch.int32, device='cuda')
    _lib.check(_lib.load().nlb_grid_corner_indices(xd.data_ptr(), od.data_ptr(), idx.data_ptr(), x.shape[0], 3, L,
                                                   1.0, 16, 0, 0, _lib.stream()))
    got = idx.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    for l in range(L):
        want, _, valid, *_ = go.corner_setup(x, l, 1.0, 16, offs.astype(np.int64))
        want = want.numpy()
        want[~valid.numpy()] = 0xFFFFFFFF
        assert np.array_equal(got[l], want), f'level {l}'


@pytest.mark.parametrize('name', list(TABLES))
def test_forward_backward_vs_oracle(name):
    from nerf_lidar_b200.gridencoder import GridEncoder
    L, C = TABLES[name]
    desired = 16 * 2 ** (L - 1)
    enc = GridEncoder(3, L, C, base_resolution=16, desired_resolution=desired, log2_hashmap_size=21).cuda()
    g = torch.Generator().manual_seed(3)
    emb = (torch.rand(enc.embeddings.shape, generator=g) * 2 - 1)
    enc.embeddings.data.copy_(emb)
    x = _points(2048, 2) * 2 - 1  # GridEncoder takes [-1, 1]
    xd = x.cuda().requires_grad_(True)
    out = enc(xd, bound=1)
    want, dy_dx = go.grid_encode_forward((x + 1) / 2, emb, enc.offsets.cpu(), 1.0, 16, calc_dy_dx=True)
    want_bl = want.permute(1, 0, 2).reshape(x.shape[0], L * C)
    assert_close(out, want_bl, 1e-5, f'{name} forward')
    gout = torch.randn(x.shape[0], L * C, generator=g)
    out.backward(gout.cuda())
    g_lbc = gout.reshape(x.shape[0], L, C).permute(1, 0, 2).contiguous()
    ge, gi = go.grid_encode_backward(g_lbc, (x + 1) / 2, emb, enc.offsets.cpu(), 1.0, 16, dy_dx)
    assert_close(enc.embeddings.grad, ge, 1e-5, f'{name} grad_embeddings')
    assert_close(xd.grad, gi / 2, 2e-4, f'{name} grad_inputs')  # d(x+1)/2


def test_reference_error_behaviour():
    from nerf_lidar_b200 import _gridencoder as be
    x = torch.rand(8, 3)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        be.grid_encode_forward(x, x, x.int(), x, 8, 3, 1, 1, 1.0, 16, None, 0, False, 0)
    xd = torch.rand(8, 3, device='cuda')
    offs = torch.tensor([0, 4920], dtype=torch.int32, device='cuda')
    emb = torch.rand(4920, 3, device='cuda')
    with pytest.raises(RuntimeError, match='C must be 1, 2, 4, or 8'):
        be.grid_encode_forward(xd, emb, offs, torch.empty(1, 8, 3, device='cuda'), 8, 3, 3, 1, 1.0, 16, None, 0, False, 0)
    # empty input is a no-op
    be.grid_encode_forward(xd[:0].contiguous(), emb[:, :1].contiguous(), offs, torch.empty(1, 0, 1, device='cuda'),
                           0, 3, 1, 1, 1.0, 16, None, 0, False, 0)


def test_total_variation_runs_and_is_finite():
    from nerf_lidar_b200.gridencoder import GridEncoder
    enc = GridEncoder(3, 4, 2, base_resolution=16, desired_resolution=128, log2_hashmap_size=15).cuda()
    enc.embeddings.data.uniform_(-1, 1)
    enc.embeddings.grad = torch.zeros_like(enc.embeddings)
    enc.grad_total_variation(weight=1e-3, B=10000)
    g = enc.embeddings.grad
    assert torch.isfinite(g).all() and float(g.abs().sum()) > 0
