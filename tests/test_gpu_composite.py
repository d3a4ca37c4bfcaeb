"""Compositing kernels against the oracle: fp32 within 1e-5 relative; backward
against torch autograd of the oracle."""
import pytest
import torch

from oracle import zipnerf_oracle as zo
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def _inputs(N, S, K, seed):
    g = torch.Generator().manual_seed(seed)
    dens = torch.rand(N, S, generator=g) * 40
    dens[3] = 0
    dens[4] *= 1000
    s = torch.sort(torch.rand(N, S + 1, generator=g), -1).values
    near, far = torch.full((N, 1), 2 / 60.), torch.full((N, 1), 500 / 60.)
    t = zo.s_to_t(s, near, far)
    dirs = torch.randn(N, 3, generator=g)
    rgb = torch.rand(N, S, 3, generator=g)
    sem = torch.softmax(torch.randn(N, S, K, generator=g), -1)
    inten = torch.randn(N, S, 1, generator=g)
    return dens, t, dirs, far, rgb, sem, inten


@pytest.mark.parametrize('S,N', [(32, 515), (64, 515), (256, 515), (32, 33001), (64, 32771)])
def test_forward_vs_oracle(S, N):
    """Small ray counts run the warp-per-ray kernel, >= 32768 rays the thread-per-ray kernel (K = 19)."""
    from nerf_lidar_b200 import ops
    K = 19
    dens, t, dirs, far, rgb, sem, inten = _inputs(N, S, K, S)
    w, _, _ = zo.alpha_weights(dens, t, dirs, True)
    want = zo.composite(rgb, w, t, far, 1.0, sem, inten, True)
    c = lambda x: x.cuda()
    got = ops.composite(c(dens), c(t), c(dirs), c(far), c(rgb), c(sem), c(inten), 1.0, True, True)
    assert_close(got['weights'], w, 1e-5, 'weights')
    for k in ('rgb', 'depth', 'acc', 'semantic', 'intensity', 'distance_mean'):
        assert_close(got[k], want[k], 1e-5, k)
    pct = got['distance_percentiles']
    assert_close(pct[:, 0], want['distance_percentile_5'], 1e-5, 'p5')
    assert_close(pct[:, 1], want['distance_median'], 1e-5, 'median')
    assert_close(pct[:, 2], want['distance_percentile_95'], 1e-5, 'p95')
    assert torch.allclose(got['weights'].sum(-1), torch.ones(N, device='cuda'), atol=1e-5)


@pytest.mark.parametrize('N,S', [(100, 64), (515, 64), (131, 37), (33001, 64), (32771, 37), (40000, 32)])
def test_proposal_level_without_colour(N, S):
    """Proposal levels (no colour / class / intensity inputs): every output against the oracle; >= 32768 rays
    run the thread-per-ray kernels (direct 16-byte loads when S % 4 == 0, the staged variant otherwise), incl.
    ragged ray counts and a sample count that is not a multiple of the chunk."""
    from nerf_lidar_b200 import ops
    dens, t, dirs, far, *_ = _inputs(N, S, 19, 5 + S)
    w, _, _ = zo.alpha_weights(dens, t, dirs, True)
    want = zo.composite(torch.zeros(N, S, 3), w, t, far, 1.0, None, None, True)
    got = ops.composite(dens.cuda(), t.cuda(), dirs.cuda(), far.cuda())
    assert got['semantic'] is None and got['intensity'] is None
    assert_close(got['weights'], w, 1e-5, 'weights')
    assert_close(got['rgb'], want['rgb'], 1e-5, 'rgb', atol=1e-6)
    for k in ('depth', 'acc', 'distance_mean'):
        assert_close(got[k], want[k], 1e-5, k)
    pct = got['distance_percentiles']
    assert_close(pct[:, 0], want['distance_percentile_5'], 1e-5, 'p5')
    assert_close(pct[:, 1], want['distance_median'], 1e-5, 'median')
    assert_close(pct[:, 2], want['distance_percentile_95'], 1e-5, 'p95')


@pytest.mark.parametrize('S', [32, 64])
def test_backward_vs_autograd(S):
    from nerf_lidar_b200 import ops
    N, K = 200, 19
    dens, t, dirs, far, rgb, sem, inten = _inputs(N, S, K, 77 + S)
    dens = dens * 0.2
    g = torch.Generator().manual_seed(9)
    cw, crgb, cdep, csem, cint, cacc = (torch.randn(N, S, generator=g), torch.randn(N, 3, generator=g),
                                        torch.randn(N, generator=g), torch.randn(N, K, generator=g),
                                        torch.randn(N, generator=g), torch.randn(N, generator=g))

    def loss(out, weights):
        return ((weights * cw.to(weights.device)).sum() + (out['rgb'] * crgb.to(weights.device)).sum()
                + (out['depth'] * cdep.to(weights.device)).sum() + (out['semantic'] * csem.to(weights.device)).sum()
                + (out['intensity'] * cint.to(weights.device)).sum() + (out['acc'] * cacc.to(weights.device)).sum())

    a = [x.clone().requires_grad_(True) for x in (dens, rgb, sem, inten)]
    w, _, _ = zo.alpha_weights(a[0], t, dirs, True)
    loss(zo.composite(a[1], w, t, far, 1.0, a[2], a[3], True), w).backward()
    b = [x.clone().cuda().requires_grad_(True) for x in (dens, rgb, sem, inten)]
    out = ops.composite(b[0], t.cuda(), dirs.cuda(), far.cuda(), b[1], b[2], b[3], 1.0, True, True)
    loss(out, out['weights']).backward()
    for name, x, y in zip(('density', 'rgb', 'semantic', 'intensity'), b, a):
        assert_close(x.grad, y.grad, 2e-4, 'grad ' + name)
