"""Size-independent properties of the hot path at BASELINE's full batch size (10 240 rays
through the model; the oracle is too slow there): sortedness and range of the resampled
intervals, partition of unity of the compositing weights under an opaque background,
linearity of the compositing in the per-sample colours, and the optimizer's fixed point."""
import ctypes as C

import pytest
import torch

from nerf_lidar_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def full_forward():
    from nerf_lidar_b200 import configs, models
    cfg = configs.nuscenes_single()
    model = models.Model(cfg).cuda()
    model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=7, table_std=0.3).items()}, strict=False)
    batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(8192, seed=7)).items()}
    with torch.no_grad():
        rend, hist = model(True, batch, 0.4, True)
    return model, batch, rend, hist


def test_intervals_sorted_and_in_range(full_forward):
    _, batch, _, hist = full_forward
    near, far = batch['near'].reshape(-1, 1), batch['far'].reshape(-1, 1)
    for lvl, h in enumerate(hist):
        s, t = h['sdist'], h['tdist']
        assert s.shape[0] == 10240 and s.shape == t.shape
        assert bool((s[:, 1:] >= s[:, :-1]).all()), f'level {lvl}: sdist not sorted'
        assert float(s.min()) >= 0.0 and float(s.max()) <= 1.0
        assert bool((t[:, 1:] >= t[:, :-1]).all()), f'level {lvl}: tdist not sorted'
        assert bool((t >= near * (1 - 1e-5)).all()) and bool((t <= far * (1 + 1e-5)).all())


def test_weights_partition_of_unity(full_forward):
    """Opaque background (nuscenes_single.gin): the last interval absorbs what is left."""
    _, _, rend, hist = full_forward
    for h, r in zip(hist, rend):
        w = h['weights']
        assert bool((w >= 0).all()) and bool(torch.isfinite(w).all())
        assert float((w.sum(-1) - 1).abs().max()) <= 2e-5
        assert float((r['acc'] - w.sum(-1)).abs().max()) <= 2e-5
    for k in ('rgb', 'depth', 'semantic', 'intensity'):
        assert bool(torch.isfinite(rend[-1][k]).all()), k
    sem = rend[-1]['semantic']
    assert float((sem.sum(-1) - 1).abs().max()) <= 1e-4   # weighted mean of per-sample softmax outputs


def test_composite_linear_in_colours(full_forward):
    from nerf_lidar_b200 import ops
    _, batch, _, hist = full_forward
    h = hist[-1]
    N, S = h['density'].shape
    g = torch.Generator(device='cuda').manual_seed(3)
    c1, c2 = torch.rand(N, S, 3, device='cuda', generator=g), torch.rand(N, S, 3, device='cuda', generator=g)
    comp = lambda c: ops.composite(h['density'], h['tdist'], batch['directions'], batch['far'], c, None, None, 0.0,
                                   True, False)['rgb']
    lhs = comp(0.3 * c1 + 0.7 * c2)
    rhs = 0.3 * comp(c1) + 0.7 * comp(c2)
    assert float((lhs - rhs).abs().max()) <= 1e-5


def test_adam_fixed_point():
    """No gradient, no decay, zero moments: the fused table pass leaves the parameters bit-identical."""
    from nerf_lidar_b200 import _lib
    offs = (C.c_int32 * 3)(0, 4920, 40864)
    p = torch.randn(40864, 1, device='cuda')
    p0 = p.clone()
    g, m, v = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    _lib.check(_lib.load().nlb_adam_table_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), offs, 2, 1, 0.0,
                                               0.01, 0.9, 0.99, 1e-15, 1, 1.0, None, _lib.stream()))
    assert torch.equal(p, p0)
