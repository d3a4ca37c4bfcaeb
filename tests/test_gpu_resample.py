"""Step-function resampling kernel against the oracle.

Inverse-CDF sampling is ill-conditioned in float32 where a wide interval carries a
tiny weight: a 1e-7 difference in the CDF moves the sample by 1e-7 / (bin mass) of
the bin width.  The reference's own fp32 result therefore differs from the exact
answer by the same amount, so the bar is: (1) the knot search and interpolation
are bit-exact / 1e-6 given the same CDF (`sorted_interp`), (2) the kernel is as close
to the float64 evaluation of the oracle as the oracle's float32 evaluation is, and
1e-5 on well-conditioned rays."""
import numpy as np
import pytest
import torch

from oracle import zipnerf_oracle as zo
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def test_sorted_interp_indices_exact():
    from nerf_lidar_b200 import ops
    g = torch.Generator().manual_seed(0)
    N, n, nx = 257, 191, 64
    w = torch.rand(N, n - 1, generator=g)
    w[:, 10:20] = 0  # ties in the CDF (zero-weight bins)
    w = w / w.sum(-1, keepdim=True)
    cw = torch.cat([torch.zeros(N, 1), torch.cumsum(w[:, :-1], -1).clamp_max(1), torch.ones(N, 1)], -1)
    t = torch.sort(torch.rand(N, n, generator=g), -1).values
    u = torch.sort(torch.rand(N, nx, generator=g), -1).values
    u[:, 0] = 0.0
    u[:, 1] = cw[:, 15]  # exactly on a (tied) knot
    want, widx = zo.interp_sorted(u, cw, t)
    got, gidx = ops.sorted_interp(u.cuda(), cw.cuda(), t.cuda(), return_index=True)
    assert np.array_equal(gidx.cpu().numpy(), widx.numpy().astype(np.int32))
    assert_close(got, want, 1e-6, 'sorted_interp')


def _case(n_in, S, rand, power):
    g = torch.Generator().manual_seed(n_in + S + power)
    N = 300
    sd = torch.sort(torch.rand(N, n_in + 1, generator=g), -1).values
    sd[:, 0], sd[:, -1] = 0.0, 1.0
    sd[5, 10:14] = sd[5, 10]  # zero-width intervals
    w = torch.rand(N, n_in, generator=g) ** power
    w[7, :20] = 0
    w = w / w.sum(-1, keepdim=True)
    near = torch.full((N, 1), 2 / 60.)
    far = torch.full((N, 1), 500 / 60.)
    jit = torch.rand(N, 1, generator=g) if rand else None
    return sd, w, near, far, jit


@pytest.mark.parametrize('rand', [False, True])
@pytest.mark.parametrize('n_in,S', [(64, 64), (64, 32), (256, 64)])
@pytest.mark.parametrize('power', [1, 4])
def test_resample_level_vs_oracle(rand, n_in, S, power):
    from nerf_lidar_b200 import ops
    sd, w, near, far, jit = _case(n_in, S, rand, power)
    prod = n_in
    want_s, want_idx, dil_s, logits = zo.resample_level(sd, w, 1, S, prod, 0.5, jit)
    want_t = zo.s_to_t(want_s, near, far)
    d = lambda x: None if x is None else x.double()
    true_s, _, _, _ = zo.resample_level(d(sd), d(w), 1, S, prod, 0.5, d(jit))
    true_t = zo.s_to_t(true_s, d(near), d(far))
    dilation = 0.0025 + 0.5 / prod
    anneal = (10 * 0.5) / (9 * 0.5 + 1)
    got_s, got_t, got_idx = ops.resample_level(sd.cuda(), w.cuda(), near.cuda(), far.cuda(), S, True, dilation,
                                               anneal, None if jit is None else jit.cuda(), rand, return_index=True)
    got_s, got_t = got_s.cpu(), got_t.cpu()
    e_ref_s = float((want_s.double() - true_s).abs().max())
    e_ker_s = float((got_s.double() - true_s).abs().max())
    e_ref_t = float((want_t.double() - true_t).abs().max())
    e_ker_t = float((got_t.double() - true_t).abs().max())
    assert e_ker_s <= 4 * e_ref_s + 1e-6, (e_ker_s, e_ref_s)
    assert e_ker_t <= 4 * e_ref_t + 1e-5, (e_ker_t, e_ref_t)
    # direct comparison: the median is at rounding level, the tail is conditioning
    assert float((got_s - want_s).abs().median()) <= 5e-7
    assert_close(got_s, want_s, 1e-4 if power == 1 else 1e-3, 'sdist')
    # sample ("interval") indices: identical, except where the centre u lies within rounding of a CDF knot --
    # the kernel sums the CDF with a warp scan, torch.cumsum sequentially, so the two CDFs differ by a few ulp of
    # 1.0 and a centre that close to a knot may land on either side (the interpolated value is continuous there)
    gi, wi = got_idx.cpu().numpy().astype(np.int64), want_idx.numpy().astype(np.int64)
    cw = zo.cdf_from_logits(logits).numpy()
    u = zo.sample_u(S, jit, sd.shape[0]).expand(sd.shape[0], S).numpy()
    rows, cols = np.nonzero(gi != wi)
    assert rows.size <= 5e-3 * gi.size, f'sample-index mismatch rate {rows.size / gi.size}'
    lo, hi = np.minimum(gi, wi)[rows, cols], np.maximum(gi, wi)[rows, cols]
    gap = np.maximum(np.abs(u[rows, cols] - cw[rows, lo + 1]), np.abs(u[rows, cols] - cw[rows, hi]))
    assert rows.size == 0 or gap.max() <= 4 * 1.2e-7, f'{rows.size} mismatching centres, farthest {gap.max():.3e} from its knot'
    assert torch.all(got_s[:, 1:] >= got_s[:, :-1])


def test_level0_is_regular():
    from nerf_lidar_b200 import ops
    N = 64
    near = torch.full((N, 1), 2 / 60., device='cuda')
    far = torch.full((N, 1), 500 / 60., device='cuda')
    s, t = ops.resample_level(None, None, near, far, 64, False, 0.5025, 1.0, None, False)
    want_s, _ = zo.sample_intervals(torch.tensor([[0., 1.]]).expand(N, 2), torch.zeros(N, 1), 64, None, 0., 1.)
    assert_close(s, want_s, 1e-6, 'level0 sdist')
    assert_close(t, zo.s_to_t(want_s, near.cpu(), far.cpu()), 1e-5, 'level0 tdist')


def test_bad_arguments_raise():
    from nerf_lidar_b200 import ops
    near = torch.ones(4, 1, device='cuda')
    with pytest.raises(RuntimeError, match='num_samples must be > 1'):
        ops.resample_level(None, None, near, near * 2, 1, False, 0.1, 1.0, None, False)
