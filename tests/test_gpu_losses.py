"""Fused distortion / anti-interlevel loss kernels against the oracle (autograd)."""
import numpy as np
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


@pytest.mark.parametrize('step,lidar_sup,only_lidar', [(6000, True, False), (1000, True, False), (4000, False, False),
                                                       (6000, True, True)])
def test_fused_supervision_losses_vs_oracle(step, lidar_sup, only_lidar):
    """csrc/render_losses.cu (data, depth incl. the 0.9-quantile gate, semantic CE, intensity, edge-aware
    smoothness, all under the dataset mask of Z/train.py:286-327) against the oracle's restatement of the
    reference training loop (oracle/train_oracle.losses, itself pinned live against Z/internal/train_utils.py):
    values and gradients w.r.t. the rendered rgb / depth / semantic / intensity."""
    from nerf_lidar_b200 import configs, synthetic, train
    cfg = configs.nuscenes_single()
    cfg.lidar_supervision = lidar_sup
    cfg.only_lidar_supervison = only_lidar
    B = 8192
    num_patch = (B // 4) // 1024
    cpu = synthetic.to_torch(synthetic.make_train_batch(B, seed=9))
    assert 0.05 < float((cpu['mask'] == 0).float().mean()) < 0.2        # the mask is not trivial
    assert float((cpu['mask'][:2048] == 0).float().mean()) > 0.05       # ... on the patch rays either
    batch = {k: v.cuda() for k, v in cpu.items()}
    N = cpu['origins'].shape[0]
    g = torch.Generator().manual_seed(1)
    rgb = torch.rand(N, 3, generator=g)
    depth = (cpu['depth'] + torch.randn(N, generator=g) * 0.3).abs().add(0.05)
    sem = torch.softmax(torch.randn(N, 19, generator=g) * 2, dim=-1)
    inten = torch.rand(N, generator=g)
    a = [t.clone().requires_grad_(True) for t in (rgb, depth, sem, inten)]
    b = [t.clone().cuda().requires_grad_(True) for t in (rgb, depth, sem, inten)]
    want = to.losses(cpu, [dict(rgb=a[0], depth=a[1], semantic=a[2], intensity=a[3])], None, step, num_patch,
                     end_step=cfg.end_step, start_step=cfg.start_step, lidar_supervision=lidar_sup,
                     only_lidar_supervision=only_lidar, pose_refine=cfg.pose_refine, regularisers=False)
    sum(want.values()).backward()
    cfg.anti_interlevel_loss_mult = 0.
    cfg.distortion_loss_mult = 0.
    got = train.compute_losses(batch, [dict(rgb=b[0], depth=b[1], semantic=b[2], intensity=b[3])],
                               [dict(sdist=torch.zeros(1), weights=torch.zeros(1))], cfg, step, num_patch)
    sum(got.values()).backward()
    assert set(want) == set(got)
    for k in want:
        assert abs(float(got[k]) - float(want[k])) <= 2e-5 * max(abs(float(want[k])), 1e-6), (k, float(got[k]), float(want[k]))
    for name, ta, tb in zip(('rgb', 'depth', 'semantic', 'intensity'), a, b):
        assert_close(tb.grad, ta.grad, 2e-5, 'grad ' + name)


def test_patch_layout_violation_is_loud():
    """The patch kernel smooths the LEADING num_patch * P * P rays (datasets.py:356-366); a patch_mask that says
    otherwise must not be smoothed silently (the reference gathers by patch_mask == 1)."""
    from nerf_lidar_b200 import configs, synthetic, train
    cfg = configs.nuscenes_single()
    cfg.anti_interlevel_loss_mult = 0.
    cfg.distortion_loss_mult = 0.
    batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(4096, seed=3)).items()}
    N = batch['origins'].shape[0]
    rend = [dict(rgb=torch.rand(N, 3, device='cuda'), depth=torch.rand(N, device='cuda') + 0.1,
                 semantic=torch.softmax(torch.randn(N, 19, device='cuda'), -1), intensity=torch.rand(N, device='cuda'))]
    hist = [dict(sdist=torch.zeros(1), weights=torch.zeros(1))]
    ok = train.compute_losses(batch, rend, hist, cfg, 6000, 1)
    assert np.isfinite(float(ok['d_smo'])) and float(ok['d_smo']) > 0
    batch['patch_mask'] = batch['patch_mask'].roll(7)
    bad = train.compute_losses(batch, rend, hist, cfg, 6000, 1)
    assert np.isnan(float(bad['d_smo'])) and np.isnan(float(bad['s_smo']))


@pytest.mark.parametrize('step,num_patch_on,hash_decay', [(6000, True, True), (3000, True, False), (6000, False, True)])
def test_fused_loss_assembly_vs_summed_dictionary(step, num_patch_on, hash_decay):
    """train.compute_losses_fused (ops.interlevel_total + ops.main_loss: the dictionary's sums from k_weighted_sums,
    every gradient seed from k_scale_tensors) against `compute_losses` summed and differentiated by torch: every
    entry, main / prop / total, and the gradients w.r.t. the rendered rgb / depth / semantic / intensity, the
    final-level weights (distortion) and both proposal levels' weights (anti-interlevel); also with a non-unit
    seed, which the reference never uses but autograd allows."""
    from nerf_lidar_b200 import configs, synthetic, train
    cfg = configs.nuscenes_single()
    B = 4096
    num_patch = (B // 4) // 1024 if num_patch_on else 0
    cpu = synthetic.to_torch(synthetic.make_train_batch(B, seed=11))
    batch = {k: v.cuda() for k, v in cpu.items()}
    N = cpu['origins'].shape[0]
    g = torch.Generator().manual_seed(2)
    rgb = torch.rand(N, 3, generator=g)
    depth = (cpu['depth'] + torch.randn(N, generator=g) * 0.3).abs().add(0.05)
    sem = torch.softmax(torch.randn(N, 19, generator=g) * 2, dim=-1)
    inten = torch.rand(N, generator=g)
    s2, w2 = _hist(N, 32, 3)
    s0, w0 = _hist(N, 64, 4)
    s1, w1 = _hist(N, 64, 5)
    hd = torch.tensor(0.0123, device='cuda')
    reg = torch.tensor(0.5, device='cuda')

    def leaves():
        return [t.clone().cuda().requires_grad_(True) for t in (rgb, depth, sem, inten, w0, w1, w2)]

    def setup(l):
        rend = [dict(), dict(), dict(rgb=l[0], depth=l[1], semantic=l[2], intensity=l[3])]
        if hash_decay:
            rend[-1]['hash_decay'] = hd
        hist = [dict(sdist=s0.cuda(), weights=l[4]), dict(sdist=s1.cuda(), weights=l[5]), dict(sdist=s2.cuda(), weights=l[6])]
        return rend, hist

    for seed in (1.0, 0.37):
        a, b = leaves(), leaves()
        rend, hist = setup(a)
        want = train.compute_losses(batch, rend, hist, cfg, step, num_patch)
        want['latent_reg'] = reg
        w_main = sum(v for k, v in want.items() if k != 'interlevel')
        w_prop = want['interlevel']
        w_total = w_main.detach() + w_prop.detach() + (hd if hash_decay else 0.)
        torch.autograd.backward([w_main, w_prop], [torch.tensor(seed, device='cuda')] * 2)
        rend, hist = setup(b)
        vals, main, prop = train.compute_losses_fused(batch, rend, hist, cfg, step, num_patch, {'latent_reg': reg})
        torch.autograd.backward([main, prop], [torch.tensor(seed, device='cuda')] * 2)
        assert set(vals) == set(want) | {'loss'} | ({'hash_decay'} if hash_decay else set())
        rel = lambda x, y: abs(float(x) - float(y)) / max(abs(float(y)), 1e-12)
        for k in want:
            assert rel(vals[k], want[k]) < 2e-6, (k, float(vals[k]), float(want[k]))
        assert rel(main, w_main) < 2e-6 and rel(prop, w_prop) < 2e-6 and rel(vals['loss'], w_total) < 2e-6
        if hash_decay:
            assert float(vals['hash_decay']) == float(hd)
        assert not any(v.requires_grad for v in vals.values())
        for name, ta, tb in zip(('rgb', 'depth', 'semantic', 'intensity', 'w_prop0', 'w_prop1', 'w_final'), a, b):
            assert tb.grad is not None and tb.grad.shape == ta.grad.shape, name
            assert_close(tb.grad, ta.grad, 2e-6, 'grad ' + name)


def test_weighted_sums_and_scale_tensors():
    """The two generic launches of the loss assembly against torch: weighted sums with and without per-element
    weights, a term that re-uses an earlier output, outputs no term names staying untouched, zero coefficients not
    reading their (NaN) input; scaled copies with one and two sources and device scalars."""
    from nerf_lidar_b200 import ops
    torch.manual_seed(0)
    x, y, w = torch.randn(10007, device='cuda'), torch.randn(5, 77, device='cuda'), torch.rand(10007, device='cuda')
    nan = torch.full((3,), float('nan'), device='cuda')
    out = torch.full((5,), -1.0, device='cuda')
    ops.weighted_sums([(x, None, 0.5, 0), (y, None, 2.0, 1), (x, w, 1.0, 2), (0, None, 3.0, 1), (nan, None, 0.0, 3)], out)
    want = torch.stack([0.5 * x.sum(), 2.0 * y.sum() + 1.5 * x.sum(), (x * w).sum(), x.new_zeros(()), x.new_tensor(-1.0)])
    assert_close(out, want, 1e-5, 'weighted sums', atol=1e-4)
    g, s1, s2 = torch.tensor(0.3, device='cuda'), torch.tensor([2.0], device='cuda'), torch.tensor([-4.0], device='cuda')
    d1, d2 = torch.empty_like(x), torch.empty_like(y)
    ops.scale_tensors([(x, w, d1, g, s1, s2, 1.0, 0.5), (y, None, d2, None, s1, None, 3.0, 0.)])
    assert_close(d1, x * 0.6 + w * (0.5 * 0.3 * -4.0), 1e-6, 'two sources')
    assert_close(d2, y * 6.0, 1e-6, 'one source')
    with pytest.raises(RuntimeError):
        ops.weighted_sums([(x, None, 1.0, 7)], out)
