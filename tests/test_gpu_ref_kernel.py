"""The B200 grid-encoder kernels and the CPU restatement against the REFERENCE'S OWN CUDA kernel
(Z/gridencoder/src/gridencoder.cu:87-245 forward, :248-340 backward, :343-369 input backward,
:473-548 total variation), compiled unmodified into `oracle/_ref/_gridencoder_ref.so` by
oracle/build_ref.py and executed here on the same tensors.

This is what pins a9 / a10 of SURVEY 8: forward values and `dy_dx` bit for bit, corner rows exactly,
gradients to atomic-order noise, and `oracle/grid_oracle.py` (the checker of every other test) against the
kernel it restates."""
import numpy as np
import pytest
import torch

from oracle import grid_oracle as go
from oracle import ref_grid
from tests.helpers import assert_close

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_grid.available(), reason='oracle/_ref/_gridencoder_ref.so not built')]

TABLES = {'prop0': (6, 1, 512), 'prop1': (8, 1, 2048), 'nerf': (10, 4, 8192)}


def _points(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, generator=g)
    x[0] = torch.tensor([0., 0., 0.])
    x[1] = torch.tensor([1., 1., 1.])
    x[2] = torch.tensor([0.5, 0.25, 0.75])
    x[3] = torch.tensor([1.0000001, 0.5, 0.5])          # out of range -> zeros
    x[4] = torch.tensor([-1e-7, 0.5, 0.5])
    x[5] = torch.tensor([(7 - 0.5) / 15, (3 - 0.5) / 31, (100 - 0.5) / 8191])   # exact cell boundaries
    x[6] = torch.tensor([np.nextafter(np.float32(1), np.float32(0)), 0.999999, 1e-8])
    x[7] = torch.tensor([0.5 / 15, 1.5 / 15, 14.5 / 15])
    return x


def _encoder(name, seed=3):
    from nerf_lidar_b200.gridencoder import GridEncoder
    L, C, desired = TABLES[name]
    enc = GridEncoder(3, L, C, base_resolution=16, desired_resolution=desired, log2_hashmap_size=21).cuda()
    g = torch.Generator().manual_seed(seed)
    enc.embeddings.data.copy_(torch.rand(enc.embeddings.shape, generator=g) * 2 - 1)
    return enc


def _ours_forward(x, enc, calc_dy_dx):
    from nerf_lidar_b200 import _gridencoder as be
    B, L, C = x.shape[0], enc.num_levels, enc.level_dim
    out = torch.empty(L, B, C, device='cuda')
    dy_dx = torch.empty(B, L * 3 * C, device='cuda') if calc_dy_dx else None
    be.grid_encode_forward(x, enc.embeddings.data, enc.offsets, out, B, 3, C, L, float(np.log2(enc.per_level_scale)),
                           enc.base_resolution, dy_dx, 0, False, 0)
    return out, dy_dx


@pytest.mark.parametrize('name', list(TABLES))
def test_forward_and_dy_dx_bit_identical(name):
    enc = _encoder(name)
    x = _points(16384, 1).cuda()
    ours, ours_d = _ours_forward(x, enc, True)
    ref, ref_d = ref_grid.encode_forward(x, enc.embeddings.data, enc.offsets, enc.per_level_scale,
                                         enc.base_resolution, calc_dy_dx=True, permute=False)
    assert torch.equal(ours, ref), f'{name}: max |d| {float((ours - ref).abs().max()):.3e}'
    assert torch.equal(ours_d, ref_d), f'{name} dy_dx: max |d| {float((ours_d - ref_d).abs().max()):.3e}'
    assert float(ours[:, 3:5].abs().max()) == 0.0   # out-of-range points


@pytest.mark.parametrize('name', list(TABLES))
def test_backward_vs_reference_kernel(name):
    from nerf_lidar_b200 import _gridencoder as be
    enc = _encoder(name)
    L, C = enc.num_levels, enc.level_dim
    x = _points(8192, 2).cuda()
    S = float(np.log2(enc.per_level_scale))
    _, dy_dx = ref_grid.encode_forward(x, enc.embeddings.data, enc.offsets, enc.per_level_scale, enc.base_resolution,
                                       calc_dy_dx=True)
    g = torch.randn(x.shape[0], L * C, generator=torch.Generator().manual_seed(4)).cuda()
    ref_ge, ref_gi = ref_grid.encode_backward(g, x, enc.embeddings.data, enc.offsets, enc.per_level_scale,
                                              enc.base_resolution, dy_dx)
    g_lbc = g.view(-1, L, C).permute(1, 0, 2).contiguous()
    ge, gi = torch.zeros_like(enc.embeddings.data), torch.zeros_like(x)
    be.grid_encode_backward(g_lbc, x, enc.embeddings.data, enc.offsets, ge, x.shape[0], 3, C, L, S,
                            enc.base_resolution, dy_dx, gi, 0, False, 0)
    # same products, atomics in a different order: fp32 summation noise only
    assert_close(ge, ref_ge, 1e-6, f'{name} grad_embeddings')
    assert torch.equal(ge != 0, ref_ge != 0), 'the two kernels touch different rows'
    assert_close(gi, ref_gi, 1e-5, f'{name} grad_inputs')


@pytest.mark.parametrize('name', list(TABLES))
def test_corner_rows_exact(name):
    """The rows the reference kernel writes for ONE point (backward with unit gradient) are exactly the
    level-local corner indices of nlb_grid_corner_indices -- dense and hashed levels, real table sizes."""
    from nerf_lidar_b200 import _lib
    enc = _encoder(name)
    L, C = enc.num_levels, enc.level_dim
    offs = enc.offsets.cpu().numpy().astype(np.int64)
    pts = _points(40, 7)
    pts = pts[[0, 2, 6] + list(range(8, 40))].contiguous().cuda()     # in-range points only
    idx = torch.empty(L, pts.shape[0], 8, dtype=torch.int32, device='cuda')
    _lib.check(_lib.load().nlb_grid_corner_indices(pts.data_ptr(), enc.offsets.data_ptr(), idx.data_ptr(),
                                                   pts.shape[0], 3, L, float(np.log2(enc.per_level_scale)), 16, 0, 0,
                                                   _lib.stream()))
    idx = idx.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    g = torch.ones(1, L * C, device='cuda')
    for p in range(pts.shape[0]):
        ge, _ = ref_grid.encode_backward(g, pts[p:p + 1].contiguous(), enc.embeddings.data, enc.offsets,
                                         enc.per_level_scale, enc.base_resolution)
        rows = torch.nonzero(ge[:, 0]).reshape(-1).cpu().numpy()
        for l in range(L):
            ref_rows = set((rows[(rows >= offs[l]) & (rows < offs[l + 1])] - offs[l]).tolist())
            ours = set(idx[l, p].tolist())
            # a corner with an exactly-zero interpolation weight leaves no trace in the reference's gradient
            assert ref_rows <= ours, (name, p, l, ref_rows, ours)
            if p >= 3:
                assert ref_rows == ours, (name, p, l, ref_rows, ours)


@pytest.mark.parametrize('name', list(TABLES))
def test_cpu_restatement_vs_reference_kernel(name):
    """oracle/grid_oracle.py -- the checker behind every other grid / encode test and the stand-in for the
    CUDA kernel in the golden fixtures -- against the kernel it restates."""
    enc = _encoder(name)
    L, C = enc.num_levels, enc.level_dim
    x = _points(4096, 5)
    emb, offs = enc.embeddings.data.cpu(), enc.offsets.cpu()
    want, want_d = go.grid_encode_forward(x, emb, offs, float(np.log2(enc.per_level_scale)), 16, calc_dy_dx=True)
    ref, ref_d = ref_grid.encode_forward(x.cuda(), enc.embeddings.data, enc.offsets, enc.per_level_scale, 16,
                                         calc_dy_dx=True, permute=False)
    assert_close(want, ref.cpu(), 5e-7, f'{name} forward')
    assert_close(want_d, ref_d.cpu(), 2e-6, f'{name} dy_dx')
    g = torch.randn(x.shape[0], L * C, generator=torch.Generator().manual_seed(6))
    g_lbc = g.view(-1, L, C).permute(1, 0, 2).contiguous()
    want_ge, want_gi = go.grid_encode_backward(g_lbc, x, emb, offs, float(np.log2(enc.per_level_scale)), 16, want_d)
    ref_ge, ref_gi = ref_grid.encode_backward(g.cuda(), x.cuda(), enc.embeddings.data, enc.offsets,
                                              enc.per_level_scale, 16, ref_d)
    assert_close(want_ge, ref_ge.cpu(), 1e-6, f'{name} grad_embeddings')
    assert_close(want_gi, ref_gi.cpu(), 1e-5, f'{name} grad_inputs')


def test_total_variation_vs_reference_kernel():
    from nerf_lidar_b200 import _gridencoder as be
    from nerf_lidar_b200.gridencoder import GridEncoder
    enc = GridEncoder(3, 6, 2, base_resolution=16, desired_resolution=512, log2_hashmap_size=17).cuda()
    enc.embeddings.data.uniform_(-1, 1)
    x = torch.rand(20000, 3, generator=torch.Generator().manual_seed(8)).cuda()
    S = float(np.log2(enc.per_level_scale))
    ours, ref = torch.zeros_like(enc.embeddings.data), torch.zeros_like(enc.embeddings.data)
    be.grad_total_variation(x, enc.embeddings.data, ours, enc.offsets, 1e-3, x.shape[0], 3, 2, 6, S, 16, 0, False)
    ref_grid.backend().grad_total_variation(x, enc.embeddings.data, ref, enc.offsets, 1e-3, x.shape[0], 3, 2, 6, S, 16,
                                            0, False)
    assert float(ref.abs().sum()) > 0
    assert_close(ours, ref, 2e-6, 'total-variation gradient')


@pytest.mark.parametrize('rand', [False, True])
def test_fused_nerf_encode_vs_reference_chain(rand, full_state_dict_visible):
    """The fused NeRF-level encode (cast_rays + contract + gather + erf-mean in one kernel, scatter in
    another) against the chain it replaces, run on the reference kernel: kernel_grid `[L,B,C]` + permute
    copy (Z/gridencoder/grid.py:54-57) + erf re-weighting and mean (Z/internal/models.py:974-977), and
    kernel_grid_backward on the autograd of that chain -- on the fused kernel's own sample points."""
    from nerf_lidar_b200 import configs, models, ops, synthetic
    from oracle import zipnerf_oracle as zo
    sd = full_state_dict_visible
    batch = synthetic.to_torch(synthetic.make_train_batch(256, seed=5))
    N, S = batch['origins'].shape[0], 32
    gen = torch.Generator().manual_seed(5)
    s = torch.sort(torch.rand(N, S + 1, generator=gen), -1).values
    t = zo.s_to_t(s, batch['near'], batch['far']).cuda()
    deg = torch.rand(N, S, 7, generator=gen).cuda() if rand else None
    model = models.Model(configs.nuscenes_single()).cuda()
    model.load_state_dict(sd, strict=False)
    enc = model.nerf_mlp.encoder
    rays = ops.RayBundle({k: v.cuda() for k, v in batch.items()})
    feat = ops.nerf_encode(t, deg, enc, rays, 0.35)
    pts = ops.sample_points(t, deg, rays)                      # [N,S,7,4]
    x = pts[..., :3].reshape(-1, 3).contiguous()
    emb = enc.embeddings.detach().clone().requires_grad_(True)

    class RefEncode(torch.autograd.Function):                   # Z/gridencoder/grid.py:24-89 on the reference kernel
        @staticmethod
        def forward(ctx, e):
            out, _ = ref_grid.encode_forward(x, e, enc.offsets, enc.per_level_scale, enc.base_resolution)
            ctx.save_for_backward(e)
            return out

        @staticmethod
        def backward(ctx, g):
            (e,) = ctx.saved_tensors
            return ref_grid.encode_backward(g.contiguous(), x, e, enc.offsets, enc.per_level_scale,
                                            enc.base_resolution)[0]

    f7 = RefEncode.apply(emb).reshape(N * S, 7, 40)
    want = ref_grid.erf_mean(f7, pts[..., 3].reshape(N * S, 7), enc.grid_sizes, 10)
    assert_close(feat, want, 1e-5, 'fused features vs reference chain')
    g = torch.randn(N * S, 40, generator=torch.Generator().manual_seed(1)).cuda()
    feat.backward(g)
    want.backward(g)
    assert_close(enc.embeddings.grad, emb.grad, 2e-5, 'fused table gradient vs reference chain')
