"""Fused distortion / anti-interlevel loss kernels against the oracle (autograd)."""
import pytest
import torch

from oracle import train_oracle as to
from oracle import zipnerf_oracle as zo
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def _hist(N, S, seed):
    g = torch.Generator().manual_seed(seed)
    s = torch.sort(torch.rand(N, S + 1, generator=g), -1).values
    s[:, 0], s[:, -1] = 0.0, 1.0
    dens = torch.rand(N, S, generator=g) * 30
    near, far = torch.full((N, 1), 2 / 60.), torch.full((N, 1), 500 / 60.)
    w, _, _ = zo.alpha_weights(dens, zo.s_to_t(s, near, far), torch.ones(N, 3), True)
    return s, w


def test_distortion_value_and_gradient():
    from nerf_lidar_b200 import ops
    N, S = 333, 32
    s, w = _hist(N, S, 1)
    wr = w.clone().requires_grad_(True)
    want = to.distortion([dict(sdist=s, weights=wr)], mult=1.0)
    want.backward()
    wc = w.clone().cuda().requires_grad_(True)
    got = ops.distortion_per_ray(s.cuda(), wc).mean()
    got.backward()
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
    assert_close(wc.grad, wr.grad, 1e-5, 'd distortion / d w')


@pytest.mark.parametrize('Sp,pulse', [(64, 0.03), (64, 0.003)])
def test_interlevel_value_and_gradient(Sp, pulse):
    from nerf_lidar_b200 import ops
    N, Sc = 257, 32
    c, w = _hist(N, Sc, 2)
    cp, wp = _hist(N, Sp, 3)
    wpr = wp.clone().requires_grad_(True)
    hist = [dict(sdist=cp, weights=wpr), dict(sdist=c, weights=w)]
    want = to.anti_interlevel(hist, pulse_width=(pulse,), mult=1.0)
    want.backward()
    wpc = wp.clone().cuda().requires_grad_(True)
    got = ops.interlevel_per_ray(c.cuda(), w.cuda(), cp.cuda(), wpc, pulse).sum() / (N * Sp)
    got.backward()
    assert abs(float(got) - float(want)) <= 2e-4 * abs(float(want)), (float(got), float(want))
    assert_close(wpc.grad, wpr.grad, 5e-4, 'd interlevel / d wp')
