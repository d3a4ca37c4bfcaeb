"""Pose-refinement window (Z/train.py:200-240): gradients of the hot path w.r.t. the ray geometry.

The reference gets them from `dy_dx` + `kernel_input_backward` (gridencoder.cu:201-244,343-369) and autograd
through the erf weights, the contraction and cast_rays; here they come from `nlb_encode_input_backward` /
`nlb_prop_input_backward`, the |d| term of the compositing and the view-direction columns of the NerfMLP.
Checked against autograd through the oracle (whose grid encoder returns input gradients exactly like
`_grid_encode.backward`, pinned against the reference binary in test_gpu_ref_kernel.py)."""
import numpy as np
import pytest
import torch

from oracle import train_oracle as to
from oracle import zipnerf_oracle as zo
from nerf_lidar_b200 import synthetic
from tests.helpers import assert_close
from tests.test_gpu_encode import _model, _setup

pytestmark = pytest.mark.gpu

GEOM = ('origins', 'directions', 'base_x', 'base_y')


def _rel_l2(got, want):
    return float((got.double() - want.double()).norm() / (want.double().norm() + 1e-30))


def _oracle_features(sd, pre, C, batch, t, deg):
    leaf = {k: batch[k].clone().requires_grad_(True) for k in GEOM}
    means, stds = zo.cast_rays(t, leaf['origins'], leaf['directions'], batch['radii'], leaf['base_x'],
                               leaf['base_y'], deg)
    feat = zo.encode_features(means, stds, sd[pre + 'encoder.embeddings'], sd[pre + 'encoder.offsets'],
                              sd[pre + 'encoder.grid_sizes'], C)
    return leaf, feat


def _cuda_rays(ops, batch):
    cu = {k: batch[k].cuda().requires_grad_(True) for k in GEOM}
    return cu, ops.RayBundle({**{k: v.cuda() for k, v in batch.items()}, **cu})


@pytest.mark.parametrize('which', ['nerf', 'prop0', 'prop1'])
def test_encode_ray_gradients_vs_oracle(which, full_state_dict_visible):
    """Oracle: autograd through cast_rays -> contract -> encoder (dy_dx) -> erf-weighted mean (-> PropMLP).

    Like the forward (test_gpu_encode.py's module note) the fine levels are ulp-sensitive: at resolution 8192 one
    fp32 ulp of a coordinate is 5e-4 of a cell, and the trilinear slope along one axis is linear in the other
    two fractions.  (1) a feature gradient on the coarse levels only (resolution <= 512) pins the whole chain --
    trilinear derivative, erf-weight derivative, contraction Jacobian, cast_rays -- to fp32 rounding; (2) the
    real op with all levels is held to the forward's whole-chain bar."""
    from nerf_lidar_b200 import ops
    sd = full_state_dict_visible
    S = 32 if which == 'nerf' else 64
    batch, t, deg = _setup(7, S, True)
    N = t.shape[0]
    model = _model(sd)
    pre = 'nerf_mlp.' if which == 'nerf' else f'prop_mlp_{which[-1]}.'
    mlp = model.get_submodule(pre[:-1])
    enc = mlp.encoder
    L, C = enc.num_levels, enc.level_dim
    gen = torch.Generator().manual_seed(3)
    # (1) coarse levels, through the encode op (any table shape)
    g = torch.randn(N * S, L, C, generator=gen)
    g[:, 6:] = 0
    g = g.reshape(N * S, L * C)
    leaf, feat = _oracle_features(sd, pre, C, batch, t, deg)
    feat.backward(g.reshape(N, S, -1))
    cu, rays = _cuda_rays(ops, batch)
    ops.nerf_encode(t.cuda(), deg.cuda(), enc, rays, 0.35).backward(g.cuda())
    for k in GEOM:
        assert torch.isfinite(cu[k].grad).all(), k
        err = _rel_l2(cu[k].grad.cpu(), leaf[k].grad)
        assert err < 1e-4, (which, 'coarse levels', k, err)
    # (2) every level, through the op the model calls
    leaf, feat = _oracle_features(sd, pre, C, batch, t, deg)
    cu, rays = _cuda_rays(ops, batch)
    if which == 'nerf':
        g = torch.randn(N * S, L * C, generator=gen)
        feat.backward(g.reshape(N, S, -1))
        ops.nerf_encode(t.cuda(), deg.cuda(), enc, rays, 0.35).backward(g.cuda())
    else:
        g = torch.randn(N, S, generator=gen)
        zo.prop_mlp(sd, pre, feat).backward(g)
        ops.prop_level(t.cuda(), deg.cuda(), mlp, rays, 0.35).backward(g.cuda())
    for k in GEOM:
        assert torch.isfinite(cu[k].grad).all(), k
        err = _rel_l2(cu[k].grad.cpu(), leaf[k].grad)
        assert err < 2e-2, (which, k, err)


@pytest.mark.parametrize('C,L,log2,rand', [(2, 7, 15, True), (8, 4, 14, False), (1, 5, 12, True)])
def test_encode_ray_gradients_other_tables_and_ragged_rows(C, L, log2, rand):
    """Other table shapes (the ObjMLP's L7 x C2 grid, level_dim 8, an odd level count), a sample count that is not
    a multiple of 32 (a warp's samples then span several rays: the per-lane atomics path), a ragged last block,
    rand=False (no rotation noise), and the empty batch; low resolutions so every level is in the fp32-exact
    regime of the test above."""
    from nerf_lidar_b200 import ops
    from nerf_lidar_b200._lib import NlbRayGrads, check, load, ptr, stream
    from nerf_lidar_b200.gridencoder import GridEncoder
    from tests.test_gpu_encode import _slice_batch
    import ctypes as Ct
    n_rays, S = 53, 9
    batch, _, _ = _setup(17, S, True)
    batch = _slice_batch(batch, n_rays)
    g0 = torch.Generator().manual_seed(100 * C + L)
    s_ = torch.sort(torch.rand(n_rays, S + 1, generator=g0), -1).values
    t = zo.s_to_t(s_, batch['near'], batch['far'])
    deg = torch.rand(n_rays, S, 7, generator=g0) if rand else None
    enc = GridEncoder(input_dim=3, num_levels=L, level_dim=C, base_resolution=16, desired_resolution=256,
                      log2_hashmap_size=log2).cuda()
    with torch.no_grad():
        enc.embeddings.copy_(torch.rand(enc.embeddings.shape, generator=g0) * 2 - 1)
    emb, offs, gs = enc.embeddings.detach().cpu(), enc.offsets.cpu(), enc.grid_sizes.cpu()
    leaf = {k: batch[k].clone().requires_grad_(True) for k in GEOM}
    means, stds = zo.cast_rays(t, leaf['origins'], leaf['directions'], batch['radii'], leaf['base_x'], leaf['base_y'], deg)
    # the oracle's encode assumes base resolution 16 and per-level scale from the grid sizes it is given
    from oracle import grid_oracle as go
    z, sd_ = zo.contract_mean_std(means.reshape(-1, 3), stds.reshape(-1))
    x01 = (z / 2 + 1) / 2
    flat = go._GridEncodeFn.apply(x01, emb, offs, float(np.log2(enc.per_level_scale)), 16, True, 0, False, 0)
    w = torch.erf(1 / torch.clamp(torch.sqrt(8 * (sd_ / 2).reshape(n_rays, S, 7)[..., None] ** 2 * gs ** 2), min=1e-10))
    feat = (flat.reshape(n_rays, S, 7, L, C) * w[..., None]).mean(-3).flatten(-2, -1)
    g = torch.randn(n_rays * S, L * C, generator=g0)
    feat.backward(g.reshape(n_rays, S, -1))
    cu, rays = _cuda_rays(ops, batch)
    degc = None if deg is None else deg.cuda()
    ops.nerf_encode(t.cuda(), degc, enc, rays, 0.35).backward(g.cuda())
    for k in GEOM:
        err = _rel_l2(cu[k].grad.cpu(), leaf[k].grad)
        assert err < 2e-4, (C, L, k, err)
    # empty batch: nothing launched, no error even with null buffers
    empty = ops.RayBundle({k: v[:0].cuda() for k, v in batch.items()})
    t0 = torch.zeros(0, S + 1, device='cuda')
    from nerf_lidar_b200.ops import _table_desc
    z4 = NlbRayGrads(None, None, None, None)
    check(load().nlb_encode_input_backward(Ct.byref(empty.desc(t0, None, 0.35)), Ct.byref(_table_desc(enc, enc.embeddings.detach())),
                                           None, Ct.byref(z4), stream()))
    # a missing gradient buffer is an argument error, not a crash
    bad = load().nlb_encode_input_backward(Ct.byref(rays.desc(t.cuda(), degc, 0.35)),
                                           Ct.byref(_table_desc(enc, enc.embeddings.detach())), ptr(g.cuda()), Ct.byref(z4), stream())
    assert bad != 0


@pytest.mark.parametrize('heads', ['torch32', 'fused'])
def test_model_ray_gradients_vs_oracle(heads):
    """The whole step: all losses differentiated w.r.t. origins, directions, viewdirs, base_x, base_y against fp32
    autograd through the oracle.  With fp32 torch heads the geometry gradients sit at the fine levels' ulp
    sensitivity (measured 0.7-1.8e-2 in the L2 sense; these are sign-mixed sums over 160 intervals x 7 samples x
    10 levels of slopes of a random table) and the view direction at 1e-6..6e-4; the tcgen05 NerfMLP with its
    bf16 operands adds the feature gradients' bf16 rounding (measured 4-6e-2, view direction 1-2.5e-2)."""
    from nerf_lidar_b200 import configs, models, train
    from tests.helpers import use_torch_heads
    B = 512
    sd = synthetic.init_state_dict(seed=23, table_std=0.2)
    batch = synthetic.to_torch(synthetic.make_train_batch(B, seed=23))
    n = batch['origins'].shape[0]
    rin = [{k: torch.from_numpy(v) for k, v in r.items()} for r in synthetic.make_rand_inputs(n, seed=23)]
    step, num_patch = 600, 0
    keys = GEOM + ('viewdirs',)
    ref = to.RefTrainer(sd)
    train_frac = float(np.clip((step - 1) / (25000 - 1), 0, 1))
    leaf = {k: batch[k].clone().requires_grad_(True) for k in keys}
    rend, hist = zo.model_forward(ref.p, {**batch, **leaf}, rin, train_frac, True, training=False)
    ls_ref = to.losses(batch, rend, hist, step, num_patch)
    sum(ls_ref.values()).backward()

    cfg = configs.nuscenes_single()
    model = models.Model(cfg, training=True).cuda()
    model.load_state_dict(sd, strict=False)
    if heads == 'torch32':
        use_torch_heads(model, torch.float32)
    tr = train.Trainer(model, cfg)
    cb = {k: v.cuda() for k, v in batch.items()}
    cu = {k: cb[k].clone().requires_grad_(True) for k in keys}
    crin = [{k: v.cuda() for k, v in r.items()} for r in rin]
    r2, h2 = model(True, {**cb, **cu}, train_frac, True, rand_inputs=crin)
    ls = train.compute_losses(cb, r2, h2, cfg, step, num_patch)
    sum(ls.values()).backward()
    for k in keys:
        got, want = cu[k].grad.cpu(), leaf[k].grad
        assert torch.isfinite(got).all(), k
        err = _rel_l2(got, want)
        bar = (2e-3 if k == 'viewdirs' else 4e-2) if heads == 'torch32' else (5e-2 if k == 'viewdirs' else 1e-1)
        assert err < bar, (heads, k, err)


def test_posenet_matches_reference_formula():
    """LearnPose / refine_rays against the closed forms of posenet_v2.py:44-63 and Z/train.py:208-221."""
    from nerf_lidar_b200 import posenet
    net = posenet.LearnPose(2, num_lidars=1, t_ratio=0.25).cuda()
    with torch.no_grad():
        net.r.copy_(torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, np.pi / 2], [0.1, -0.2, 0.3]]))
        net.t.copy_(torch.tensor([[0.0, 0.0, 0.0], [4.0, 0.0, 0.0], [1.0, 2.0, 3.0]]))
    pose = net(torch.tensor([0, 1, 2]).cuda()).detach().cpu()
    assert torch.allclose(pose[0], torch.eye(4), atol=1e-7)
    assert torch.allclose(pose[1, :3, :3], torch.tensor([[0., -1, 0], [1, 0, 0], [0, 0, 1]]), atol=1e-6)
    assert torch.allclose(pose[1, :3, 3], torch.tensor([1.0, 0.0, 0.0]))
    R = pose[2, :3, :3]
    assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-6) and abs(float(torch.det(R.detach())) - 1) < 1e-6
    # axis is invariant, the rotation angle is |r|
    r = torch.tensor([0.1, -0.2, 0.3])
    assert torch.allclose(R @ r, r, atol=1e-6)
    assert abs(float((torch.trace(R) - 1) / 2) - float(torch.cos(r.norm()))) < 1e-6
    batch = {k: torch.randn(3, 3).cuda() for k in ('origins',) + posenet.RAY_ROTATED}
    batch['glo_idx'] = torch.tensor([0, 1, 2]).cuda()
    out = posenet.refine_rays(batch, net)
    assert torch.allclose(out['origins'].cpu(), batch['origins'].cpu() + pose[:, :3, 3])
    for k in posenet.RAY_ROTATED:
        want = torch.einsum('bij,bj->bi', pose[:, :3, :3], batch[k].cpu())
        assert torch.allclose(out[k].cpu(), want, atol=1e-6), k


@pytest.mark.parametrize('graphed', [False, True])
def test_trainer_refines_poses_inside_the_window(graphed):
    """Inside the window the corrections move (Adam on the ray-geometry gradients), the graph-replayed step
    tracks the eager one, and after the window they are applied but frozen."""
    from nerf_lidar_b200 import configs, models, posenet, train
    cfg = configs.nuscenes_single()
    assert (cfg.start_step, cfg.end_step, cfg.learn_R, cfg.learn_t) == (0, 5000, True, False)
    sd = synthetic.init_state_dict(seed=31, table_std=0.2)
    batch = synthetic.make_train_batch(512, seed=31)
    batch['glo_idx'] = synthetic.sensor_index(batch, num_cams=1)
    cb = {k: v.cuda() for k, v in synthetic.to_torch(batch).items()}
    model = models.Model(cfg, training=True).cuda()
    model.load_state_dict(sd, strict=False)
    tr = train.Trainer(model, cfg)
    net, opt, lr_fn = posenet.create_posenet(1, cfg, num_lidars=1, device='cuda')
    tr.attach_posenet(net, opt, lr_fn)
    step_fn = tr.train_step_graphed if graphed else tr.train_step
    torch.manual_seed(0)
    r_hist = []
    for step in range(1, 6):
        out = step_fn(cb, step, 0)
        assert torch.isfinite(out['loss']).all()
        r_hist.append(net.r.detach().cpu().clone())
    # gin: learn_R only; Adam's first steps move every touched coordinate by ~lr
    assert net.t.grad is None and float(net.t.abs().max()) == 0.0
    assert float(r_hist[0].abs().max()) > 0.0
    lr = lr_fn(1)
    assert float(r_hist[0].abs().max()) <= 1.01 * lr
    assert float((r_hist[-1] - r_hist[0]).abs().max()) > 0.0
    # both sensors are in the batch: both rows move
    assert (r_hist[-1].abs().amax(-1) > 0).all()
    # after the window: applied without gradients, frozen
    before = net.r.detach().clone()
    out = step_fn(cb, cfg.end_step + 10, 0)
    assert torch.isfinite(out['loss']).all()
    assert torch.equal(net.r.detach(), before)
