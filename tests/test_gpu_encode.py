"""Fused cast_rays + contract + hash-grid + erf-mean kernels (NeRF level and
proposal levels) against the oracle chain.

At the finest level (resolution 8192) one float32 ulp of a coordinate is 5e-4 of
a cell, so features of a RANDOM table move by ~1e-4 relative when the sample
point is computed with a different (equally valid) fp32 operation order.  The
tests therefore split the chain: (1) the generated points against the oracle's
cast_rays + contract to a few ulp, (2) the encode + erf-mean against the oracle
evaluated ON THE KERNEL'S POINTS to 1e-5, (3) the whole chain loosely."""
import numpy as np
import pytest
import torch

from oracle import grid_oracle as go
from oracle import zipnerf_oracle as zo
from nerf_lidar_b200 import synthetic
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def _setup(seed, S, rand):
    batch = synthetic.to_torch(synthetic.make_train_batch(256, seed=seed))
    N = batch['origins'].shape[0]
    g = torch.Generator().manual_seed(seed)
    s = torch.sort(torch.rand(N, S + 1, generator=g), -1).values
    t = zo.s_to_t(s, batch['near'], batch['far'])
    deg = torch.rand(N, S, 7, generator=g) if rand else None
    return batch, t, deg


def _model(sd):
    from nerf_lidar_b200 import configs, models
    model = models.Model(configs.nuscenes_single()).cuda()
    model.load_state_dict(sd, strict=False)
    return model


def _oracle_features_from_points(pts, emb, offsets, grid_sizes, C):
    """encode + erf re-weighting + mean (models.py:974-977) on given grid-space points."""
    N, S, n, _ = pts.shape
    L = offsets.shape[0] - 1
    out, _ = go.grid_encode_forward(pts[..., :3].reshape(-1, 3), emb, offsets, 1.0, 16)
    feat = out.permute(1, 0, 2).reshape(N, S, n, L, C)
    sd = pts[..., 3]
    w = torch.erf(1 / torch.clamp(torch.sqrt(8 * sd[..., None] ** 2 * grid_sizes ** 2), min=1e-10))
    return (feat * w[..., None]).mean(-3).flatten(-2, -1)


@pytest.mark.parametrize('rand', [False, True])
def test_sample_points_vs_oracle(rand):
    from nerf_lidar_b200 import ops
    batch, t, deg = _setup(5, 32, rand)
    means, stds = zo.cast_rays(t, batch['origins'], batch['directions'], batch['radii'], batch['base_x'],
                               batch['base_y'], deg)
    z, sd = zo.contract_mean_std(means.reshape(-1, 3), stds.reshape(-1))
    want_x = ((z / 2 + 1) / 2).reshape(*means.shape)
    want_s = (sd / 2).reshape(*stds.shape)
    rays = ops.RayBundle({k: v.cuda() for k, v in batch.items()})
    pts = ops.sample_points(t.cuda(), None if deg is None else deg.cuda(), rays).cpu()
    # coordinates live in [0,1]: 4 ulp of 1.0
    assert float((pts[..., :3] - want_x).abs().max()) <= 4 * 1.2e-7
    assert_close(pts[..., 3], want_s, 1e-5, 'std')


@pytest.mark.parametrize('rand', [False, True])
def test_nerf_encode_forward_backward(rand, full_state_dict_visible):
    from nerf_lidar_b200 import ops
    sd = full_state_dict_visible
    batch, t, deg = _setup(5, 32, rand)
    model = _model(sd)
    enc = model.nerf_mlp.encoder
    rays = ops.RayBundle({k: v.cuda() for k, v in batch.items()})
    degc = None if deg is None else deg.cuda()
    feat = ops.nerf_encode(t.cuda(), degc, enc, rays, 0.35)
    N, S = t.shape[0], 32
    pts = ops.sample_points(t.cuda(), degc, rays).cpu()
    emb = sd['nerf_mlp.encoder.embeddings']
    want = _oracle_features_from_points(pts, emb, sd['nerf_mlp.encoder.offsets'], sd['nerf_mlp.encoder.grid_sizes'], 4)
    assert_close(feat.reshape(N, S, 40), want, 1e-5, 'nerf features on the kernel points')
    # whole chain against the oracle's own points (ulp-sensitive at level 9, see module doc)
    means, stds = zo.cast_rays(t, batch['origins'], batch['directions'], batch['radii'], batch['base_x'],
                               batch['base_y'], deg)
    chain = zo.encode_features(means, stds, emb, sd['nerf_mlp.encoder.offsets'], sd['nerf_mlp.encoder.grid_sizes'], 4)
    assert_close(feat.reshape(N, S, 40), chain, 1e-3, 'nerf features, whole chain')
    # coarse levels (0-5, resolution <= 512) are insensitive: tight on the whole chain
    assert_close(feat.reshape(N, S, 10, 4)[:, :, :6], chain.reshape(N, S, 10, 4)[:, :, :6], 3e-5, 'coarse levels')
    # backward: the op is linear in the table -> <g, F(e)> = <F^T g, e>, and F^T g must
    # equal the oracle's scatter on the same points
    g = torch.randn(N * S, 40, generator=torch.Generator().manual_seed(1))
    feat.backward(g.cuda())
    ge = enc.embeddings.grad.cpu()
    lhs = float((g.double() * feat.detach().cpu().double()).sum())
    rhs = float((ge.double() * emb.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * float((g.double() * feat.detach().cpu().double()).abs().sum())
    # explicit scatter oracle: d feat / d table = erf_w/7 * trilinear weights
    g_pts = g.reshape(N, S, 1, 10, 4).expand(N, S, 7, 10, 4)
    sdv = pts[..., 3]
    wj = torch.erf(1 / torch.clamp(torch.sqrt(8 * sdv[..., None] ** 2 * sd['nerf_mlp.encoder.grid_sizes'] ** 2), min=1e-10))
    g_lbc = (g_pts * wj[..., None] / 7).reshape(-1, 10, 4).permute(1, 0, 2).contiguous()
    want_ge, _ = go.grid_encode_backward(g_lbc, pts[..., :3].reshape(-1, 3), emb, sd['nerf_mlp.encoder.offsets'], 1.0, 16)
    assert_close(ge, want_ge, 2e-5, 'table gradient')


@pytest.mark.parametrize('lvl,L', [(0, 6), (1, 8)])
def test_prop_level_forward_backward(lvl, L, full_state_dict_visible):
    from nerf_lidar_b200 import _lib, ops
    import ctypes as C
    sd = full_state_dict_visible
    pre = f'prop_mlp_{lvl}.'
    batch, t, deg = _setup(6 + lvl, 64, True)
    N, S = t.shape[0], 64
    model = _model(sd)
    mlp = model.get_submodule(f'prop_mlp_{lvl}')
    rays = ops.RayBundle({k: v.cuda() for k, v in batch.items()})
    tc, degc = t.cuda(), deg.cuda()
    dens = ops.prop_level(tc, degc, mlp, rays, 0.35)
    pts = ops.sample_points(tc, degc, rays).cpu()
    emb = sd[pre + 'encoder.embeddings']
    feat = _oracle_features_from_points(pts, emb, sd[pre + 'encoder.offsets'], sd[pre + 'encoder.grid_sizes'], 1)
    feat = feat.clone().requires_grad_(True)
    p = {k: (v.clone().requires_grad_(True) if 'density_layer' in k and k.startswith(pre) else v) for k, v in sd.items()}
    want = zo.prop_mlp(p, pre, feat)
    assert_close(dens, want, 1e-5, 'prop density on the kernel points')
    # whole chain
    means, stds = zo.cast_rays(t, batch['origins'], batch['directions'], batch['radii'], batch['base_x'],
                               batch['base_y'], deg)
    chain = zo.prop_mlp(sd, pre, zo.encode_features(means, stds, emb, sd[pre + 'encoder.offsets'],
                                                    sd[pre + 'encoder.grid_sizes'], 1))
    assert_close(dens, chain, 2e-4, 'prop density, whole chain')
    gd = torch.randn(N, S, generator=torch.Generator().manual_seed(2))
    want.backward(gd)
    dens.backward(gd.cuda())
    l0, l2 = mlp.density_layer[0], mlp.density_layer[2]
    # ReLU units within rounding of zero may flip between the two evaluations; the
    # gradients are compared relative to the sum of absolute contributions
    assert_close(l0.weight.grad, p[pre + 'density_layer.0.weight'].grad, 2e-4, 'gW0')
    assert_close(l0.bias.grad, p[pre + 'density_layer.0.bias'].grad, 2e-4, 'gb0')
    assert_close(l2.weight.grad, p[pre + 'density_layer.2.weight'].grad, 2e-4, 'gW1')
    assert_close(l2.bias.grad, p[pre + 'density_layer.2.bias'].grad, 2e-4, 'gb1')
    # table gradient = scatter of d loss / d features on the kernel points
    gf = feat.grad.reshape(N, S, 1, L, 1).expand(N, S, 7, L, 1)
    sdv = pts[..., 3]
    wj = torch.erf(1 / torch.clamp(torch.sqrt(8 * sdv[..., None] ** 2 * sd[pre + 'encoder.grid_sizes'] ** 2), min=1e-10))
    g_lbc = (gf * wj[..., None] / 7).reshape(-1, L, 1).permute(1, 0, 2).contiguous()
    want_ge, _ = go.grid_encode_backward(g_lbc, pts[..., :3].reshape(-1, 3), emb, sd[pre + 'encoder.offsets'], 1.0, 16)
    assert_close(mlp.encoder.embeddings.grad, want_ge, 2e-4, 'table gradient')


def _slice_batch(batch, n):
    return {k: v[:n].contiguous() for k, v in batch.items()}


@pytest.mark.parametrize('n_rays,S', [(37, 5), (1, 3), (129, 33)])
def test_encode_ragged_sizes(n_rays, S, full_state_dict_visible):
    """Row counts that are not multiples of the warp / tile size (the scatter kernels are
    persistent, warp-aggregated and use whole-warp shuffles): forward on the kernel's points
    and the table gradient against the oracle scatter, NeRF table (C=4) and a proposal
    table (C=1)."""
    from nerf_lidar_b200 import ops
    sd = full_state_dict_visible
    batch, _, _ = _setup(11, S, True)
    batch = _slice_batch(batch, n_rays)
    g0 = torch.Generator().manual_seed(n_rays)
    s = torch.sort(torch.rand(n_rays, S + 1, generator=g0), -1).values
    t = zo.s_to_t(s, batch['near'], batch['far'])
    deg = torch.rand(n_rays, S, 7, generator=g0)
    model = _model(sd)
    rays = ops.RayBundle({k: v.cuda() for k, v in batch.items()})
    pts = ops.sample_points(t.cuda(), deg.cuda(), rays).cpu()
    for pre, enc, C in (('nerf_mlp.', model.nerf_mlp.encoder, 4), ('prop_mlp_1.', model.prop_mlp_1.encoder, 1)):
        L = enc.num_levels
        emb, offs, gs = sd[pre + 'encoder.embeddings'], sd[pre + 'encoder.offsets'], sd[pre + 'encoder.grid_sizes']
        enc.embeddings.grad = None
        feat = ops.nerf_encode(t.cuda(), deg.cuda(), enc, rays, 0.35)
        assert feat.shape == (n_rays * S, L * C)
        want = _oracle_features_from_points(pts, emb, offs, gs, C)
        assert_close(feat.reshape(n_rays, S, L * C), want, 1e-5, pre + 'features')
        g = torch.randn(n_rays * S, L * C, generator=g0)
        feat.backward(g.cuda())
        g_pts = g.reshape(n_rays, S, 1, L, C).expand(n_rays, S, 7, L, C)
        wj = torch.erf(1 / torch.clamp(torch.sqrt(8 * pts[..., 3][..., None] ** 2 * gs ** 2), min=1e-10))
        g_lbc = (g_pts * wj[..., None] / 7).reshape(-1, L, C).permute(1, 0, 2).contiguous()
        want_ge, _ = go.grid_encode_backward(g_lbc, pts[..., :3].reshape(-1, 3), emb, offs, 1.0, 16)
        assert_close(enc.embeddings.grad, want_ge, 2e-5, pre + 'table gradient')


@pytest.mark.parametrize('C,L,log2', [(2, 7, 15), (8, 4, 14), (1, 5, 12)])
def test_encode_other_level_dims(C, L, log2):
    """The fused encode for tables the static zipnerf path does not use: level_dim 2 (ObjMLP's L7 x C2 grid,
    models.py:878), level_dim 8, an odd level count; small hash maps so that most levels are hashed (the L7
    table takes the pair-lane scatter launch, the others the mixed dense + hashed one).  Forward on the
    kernel's points and the table gradient against the oracle, ragged row count."""
    from nerf_lidar_b200 import ops
    from nerf_lidar_b200.gridencoder import GridEncoder
    n_rays, S = 53, 9
    batch, _, _ = _setup(17, S, True)
    batch = _slice_batch(batch, n_rays)
    g0 = torch.Generator().manual_seed(100 * C + L)
    s = torch.sort(torch.rand(n_rays, S + 1, generator=g0), -1).values
    t = zo.s_to_t(s, batch['near'], batch['far'])
    deg = torch.rand(n_rays, S, 7, generator=g0)
    enc = GridEncoder(input_dim=3, num_levels=L, level_dim=C, base_resolution=16, log2_hashmap_size=log2).cuda()
    with torch.no_grad():
        enc.embeddings.copy_(torch.rand(enc.embeddings.shape, generator=g0) * 2 - 1)
    emb, offs, gs = enc.embeddings.detach().cpu(), enc.offsets.cpu(), enc.grid_sizes.cpu()
    rays = ops.RayBundle({k: v.cuda() for k, v in batch.items()})
    pts = ops.sample_points(t.cuda(), deg.cuda(), rays).cpu()
    feat = ops.nerf_encode(t.cuda(), deg.cuda(), enc, rays, 0.35)
    assert feat.shape == (n_rays * S, L * C)
    want = _oracle_features_from_points(pts, emb, offs, gs, C)
    assert_close(feat.reshape(n_rays, S, L * C), want, 1e-5, 'features')
    g = torch.randn(n_rays * S, L * C, generator=g0)
    feat.backward(g.cuda())
    g_pts = g.reshape(n_rays, S, 1, L, C).expand(n_rays, S, 7, L, C)
    wj = torch.erf(1 / torch.clamp(torch.sqrt(8 * pts[..., 3][..., None] ** 2 * gs ** 2), min=1e-10))
    g_lbc = (g_pts * wj[..., None] / 7).reshape(-1, L, C).permute(1, 0, 2).contiguous()
    want_ge, _ = go.grid_encode_backward(g_lbc, pts[..., :3].reshape(-1, 3), emb, offs, 1.0, 16)
    assert_close(enc.embeddings.grad, want_ge, 2e-5, 'table gradient')


def test_encode_adjoint_at_bench_size(full_state_dict_visible):
    """Size-independent property at BASELINE's full size (10 240 rays): the fused encode is
    linear in the table, so <g, F(e)> = <F^T g, e> must hold for the forward / scatter pair."""
    from nerf_lidar_b200 import ops
    sd = full_state_dict_visible
    model = _model(sd)
    batch = synthetic.to_torch(synthetic.make_train_batch(8192, seed=3))
    N = batch['origins'].shape[0]
    rays = ops.RayBundle({k: v.cuda() for k, v in batch.items()})
    for enc, S in ((model.nerf_mlp.encoder, 32), (model.prop_mlp_0.encoder, 64), (model.prop_mlp_1.encoder, 64)):
        g0 = torch.Generator(device='cuda').manual_seed(S)
        s = torch.sort(torch.rand(N, S + 1, device='cuda', generator=g0), -1).values
        t = zo.s_to_t(s.cpu(), batch['near'], batch['far']).cuda()
        deg = torch.rand(N, S, 7, device='cuda', generator=g0)
        enc.embeddings.grad = None
        feat = ops.nerf_encode(t, deg, enc, rays, 0.35)
        g = torch.randn(feat.shape, device='cuda', generator=g0)
        feat.backward(g)
        lhs = (g.double() * feat.detach().double()).sum()
        rhs = (enc.embeddings.grad.double() * enc.embeddings.detach().double()).sum()
        scale = (g.double() * feat.detach().double()).abs().sum()
        assert float((lhs - rhs).abs()) <= 1e-5 * float(scale), (S, float(lhs), float(rhs))


def test_empty_ray_batch():
    """Zero rays: every op returns empty outputs without launching."""
    from nerf_lidar_b200 import configs, models
    model = models.Model(configs.nuscenes_single()).cuda()
    model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=1).items()}, strict=False)
    batch = {k: v[:0].cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(64, seed=1)).items()}
    with torch.no_grad():
        rend, hist = model(False, batch, 1.0, True)
    assert rend[-1]['rgb'].shape == (0, 3) and rend[-1]['depth'].shape == (0,)
    assert hist[-1]['weights'].shape == (0, 32)
