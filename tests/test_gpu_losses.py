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


@pytest.mark.parametrize('step,lidar_sup', [(6000, True), (1000, True), (4000, False)])
def test_fused_supervision_losses_vs_torch(step, lidar_sup):
    """csrc/render_losses.cu (data, depth incl. the 0.9-quantile gate, semantic CE, intensity,
    edge-aware smoothness) against the plain-torch restatement: values and gradients w.r.t.
    the rendered rgb / depth / semantic / intensity."""
    from nerf_lidar_b200 import configs, synthetic, train
    cfg = configs.nuscenes_single()
    cfg.lidar_supervision = lidar_sup
    B = 8192
    num_patch = (B // 4) // 1024
    batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=9)).items()}
    N = batch['origins'].shape[0]
    g = torch.Generator(device='cuda').manual_seed(1)

    def leaves():
        rgb = torch.rand(N, 3, device='cuda', generator=g).requires_grad_(True)
        depth = (batch['depth'] + torch.randn(N, device='cuda', generator=g) * 0.3).abs().add(0.05).requires_grad_(True)
        sem = torch.softmax(torch.randn(N, 19, device='cuda', generator=g) * 2, dim=-1).requires_grad_(True)
        inten = torch.rand(N, device='cuda', generator=g).requires_grad_(True)
        return rgb, depth, sem, inten

    a = leaves()
    b = [t.detach().clone().requires_grad_(True) for t in a]
    hist = [dict(sdist=torch.zeros(1), weights=torch.zeros(1))]
    cfg.anti_interlevel_loss_mult = 0.
    cfg.distortion_loss_mult = 0.
    outs = []
    for fn, (rgb, depth, sem, inten) in ((train.compute_losses_torch, a), (train.compute_losses, b)):
        rend = [dict(rgb=rgb, depth=depth, semantic=sem, intensity=inten)]
        ls = fn(batch, rend, hist, cfg, step, num_patch)
        sum(ls.values()).backward()
        outs.append(ls)
    assert set(outs[0]) == set(outs[1])
    for k in outs[0]:
        want, got = float(outs[0][k]), float(outs[1][k])
        assert abs(got - want) <= 2e-5 * max(abs(want), 1e-6), (k, got, want)
    for name, ta, tb in zip(('rgb', 'depth', 'semantic', 'intensity'), a, b):
        assert_close(tb.grad, ta.grad, 2e-5, 'grad ' + name)
