"""Model.forward on the B200 kernels against the outputs of the reference's own
Python (tests/golden/*.npz).

Two comparisons.  TEACHER-FORCED: every stage of every level is fed the
reference's own inputs for that stage (golden sdist/tdist/density), so stage
errors do not compound and the per-stage bars apply (compositing 1e-5, densities
limited by the fp32 sensitivity of level-9 features, resampling limited by the
fp32 conditioning of CDF inversion).  FREE-RUNNING: the whole forward pass,
looser because differences compound over three levels."""
import numpy as np
import pytest
import torch

from oracle import zipnerf_oracle as zo
from nerf_lidar_b200 import synthetic
from tests.helpers import CASES, load_case, assert_close, use_torch_heads

pytestmark = pytest.mark.gpu


def _model(sd, dtype):
    from nerf_lidar_b200 import configs, models
    model = models.Model(configs.nuscenes_single()).cuda()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    model.eval()
    model.training = False
    if dtype == torch.float32:     # the reference's fp32 head arithmetic around the other kernels
        use_torch_heads(model, torch.float32)
    return model


@pytest.mark.parametrize('name', list(CASES))
def test_teacher_forced_stages(name):
    from nerf_lidar_b200 import ops
    case, golden, sd, batch, rin = load_case(name, 'cuda')
    model = _model(sd, torch.float32)
    G = {k: torch.from_numpy(v).cuda() for k, v in golden.items()}
    rays = ops.RayBundle(batch)
    samples = (64, 64, 32)
    anneal = (10 * case['train_frac']) / (9 * case['train_frac'] + 1)
    prod = 1
    for lvl, S in enumerate(samples):
        deg = rin[lvl]['deg'] if rin is not None else None
        jit = rin[lvl]['jitter'] if rin is not None else None
        # --- resampling from the reference's previous-level step function
        sd_in = None if lvl == 0 else G[f'hist{lvl - 1}_sdist']
        w_in = None if lvl == 0 else G[f'hist{lvl - 1}_weights']
        dilation = 0.0025 + 0.5 / prod
        s, t = ops.resample_level(sd_in, w_in, batch['near'], batch['far'], S, lvl > 0, dilation, anneal, jit,
                                  case['rand'])
        prod *= S
        assert_close(s, G[f'hist{lvl}_sdist'], 2e-4, f'level {lvl} sdist')
        assert_close(t, G[f'hist{lvl}_tdist'], 2e-3, f'level {lvl} tdist')
        # median error is at rounding level; the tail is the ill-conditioned bins
        med = float((s - G[f'hist{lvl}_sdist']).abs().median())
        assert med <= 2e-7, f'level {lvl} median sdist error {med}'
        # --- field query on the reference's intervals
        t_ref = G[f'hist{lvl}_tdist'].contiguous()
        with torch.no_grad():
            if lvl < 2:
                dens = ops.prop_level(t_ref, deg, model.get_submodule(f'prop_mlp_{lvl}'), rays, 0.35)
                assert_close(dens, G[f'hist{lvl}_density'], 2e-4, f'level {lvl} density')
            else:
                feat = ops.nerf_encode(t_ref, deg, model.nerf_mlp.encoder, rays, 0.35)
                res = model.nerf_mlp.heads(feat, batch['viewdirs'], S)
                assert_close(res['density'], G['hist2_density'], 1e-3, 'nerf density')
                assert_close(res['rgb'], G['hist2_rgb'], 1e-3, 'nerf rgb')
                assert_close(res['semantic'], G['hist2_semantic'], 1e-3, 'nerf semantic')
                assert_close(res['intensity'], G['hist2_intensity'], 1e-3, 'nerf intensity')
        # --- compositing of the reference's per-sample values: fp32 1e-5
        kw = {}
        if lvl == 2:
            kw = dict(rgb=G['hist2_rgb'], semantic=G['hist2_semantic'], intensity=G['hist2_intensity'])
        comp = ops.composite(G[f'hist{lvl}_density'], t_ref, rays.directions, batch['far'], **kw)
        assert_close(comp['weights'], G[f'hist{lvl}_weights'], 1e-5, f'level {lvl} weights')
        assert_close(comp['rgb'], G[f'rend{lvl}_rgb'], 1e-5, f'level {lvl} rgb', atol=1e-6)
        for k in ('depth', 'acc', 'distance_mean'):
            assert_close(comp[k], G[f'rend{lvl}_{k}'], 1e-5, f'level {lvl} {k}')
        pct = comp['distance_percentiles']
        assert_close(pct[:, 0], G[f'rend{lvl}_distance_percentile_5'], 1e-5, 'p5')
        assert_close(pct[:, 1], G[f'rend{lvl}_distance_median'], 1e-5, 'median')
        assert_close(pct[:, 2], G[f'rend{lvl}_distance_percentile_95'], 1e-5, 'p95')
        if lvl == 2:
            assert_close(comp['semantic'], G['rend2_semantic'], 1e-5, 'semantic')
            assert_close(comp['intensity'], G['rend2_intensity'], 1e-5, 'intensity')


# Bars of the free-running forward (max |d| / max |ref| per key), set from tools/parity_report.py on the B200
# (gpurun_out/parity_report.log, round 2) with ~3x margin.  north_star: fp32 compositing 1e-5, bf16 MLP outputs and
# rendered depth / intensity 1e-3.
#   fp32 head: every rendered quantity is at 1e-5 or better (measured <= 1.3e-5); the exception is intensity on the
#   "visible" random table (1.1e-4): a 1-ulp shift of a sample point is 5e-4 of a level-9 cell, i.e. ~1e-4 of a
#   random feature, and the random-init intensity head's whole output range is 0.034.
#   bf16 tensor-core head: depth 7e-5, rgb 5.6e-4, semantic 1.6e-4, distances 1.1e-4 -- inside 1e-3.  Intensity is
#   3.5e-3 of its 0.044 range = 1.5e-4 ABSOLUTE: the bf16 rounding of the (shared) weights is a systematic offset of
#   2^-9 of the head's TERMS, and at random init those terms cancel to an output 20x smaller than they are
#   (tests/test_gpu_mlp.py holds the intensity head to 1e-3 of its term scale).
_REND_BARS = {
    torch.float32: dict(rgb=3e-5, depth=3e-5, semantic=3e-5, intensity=3e-4, acc=1e-5, distance_mean=3e-5,
                        distance_median=4e-5, distance_percentile_5=3e-5, distance_percentile_95=3e-5),
    torch.bfloat16: dict(rgb=1e-3, depth=2e-4, semantic=5e-4, intensity=6e-3, acc=1e-5, distance_mean=3e-4,
                         distance_median=3e-4, distance_percentile_5=4e-4, distance_percentile_95=3e-4),
}
_INTENSITY_ABS = {torch.float32: 1.5e-5, torch.bfloat16: 3e-4}
# per-sample history: sample positions to 6e-5 (CDF inversion conditioning); per-sample values shift with them,
# so they are held on the MEDIAN (and loosely on the maximum)
_HIST_POS = 6e-5
_HIST_MED = {torch.float32: 1e-5, torch.bfloat16: 5e-4}
_HIST_MED_INTENSITY = {torch.float32: 3e-5, torch.bfloat16: 6e-3}


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('name', list(CASES))
def test_free_running_forward(name, dtype):
    """The whole Model.forward (three levels, errors compound) against the reference's own outputs.
    float32 = the reference's head arithmetic (plain torch) around the kernels; bfloat16 = the shipped path."""
    case, golden, sd, batch, rin = load_case(name, 'cuda')
    model = _model(sd, dtype)
    with torch.no_grad():
        rend, hist = model(case['rand'], batch, case['train_frac'], True, rand_inputs=rin)
    for key, ref in golden.items():
        kind, k = key.split('_', 1)
        i = int(kind[-1])
        src = hist[i] if kind.startswith('hist') else rend[i]
        got = src[k].float().cpu().numpy().reshape(ref.shape)
        scale = np.abs(ref).max() + 1e-30
        d = np.abs(got - ref)
        if kind.startswith('rend'):
            if k.startswith('ray_'):
                continue
            if k == 'rgb' and i < 2:
                assert d.max() <= 1e-6, key            # proposal levels render black
                continue
            assert d.max() <= _REND_BARS[dtype][k] * scale, f'{key}: {d.max() / scale:.3e}'
            if k == 'intensity':
                assert d.max() <= _INTENSITY_ABS[dtype], f'{key}: abs {d.max():.3e}'
        elif k in ('sdist', 'tdist'):
            assert d.max() <= _HIST_POS * scale, f'{key}: {d.max() / scale:.3e}'
        elif i < 2 and k == 'rgb':
            assert d.max() == 0.0, key
        else:
            level_dtype = dtype if i == 2 else torch.float32      # the proposal levels have no bf16 stage
            med = (_HIST_MED_INTENSITY if k == 'intensity' else _HIST_MED)[level_dtype]
            assert np.median(d) <= med * scale, f'{key}: median {np.median(d) / scale:.3e}'
            assert d.max() <= 1e-2 * scale, f'{key}: max {d.max() / scale:.3e}'


def test_lidar_sweep_subset_vs_oracle():
    """BASELINE configs[2]: 512 rays of a full 32 x 1084 LiDAR sweep rendered by the shipped path (bf16 tensor-core
    MLP, graphs, chunking) against the oracle evaluated on the same rays: depth / intensity / rgb, and the semantic
    argmax wherever the oracle's top-2 margin is clear of bf16 noise."""
    from nerf_lidar_b200 import configs, models
    cfg = configs.nuscenes_single()
    sd = synthetic.init_state_dict(seed=41, table_std=0.3)
    model = models.Model(cfg).cuda()
    model.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=False)
    sweep = synthetic.to_torch(synthetic.make_lidar_sweep(seed=41))
    n = sweep['origins'].shape[0]
    out = models.render_image(model, None, {k: v.cuda() for k, v in sweep.items()}, False, cfg, image=False, verbose=False)
    pick = torch.linspace(0, n - 1, 512).long()
    sub = {k: v[pick].contiguous() for k, v in sweep.items()}
    want, _ = zo.model_forward(sd, sub, None, 1.0)
    for k, tol in (('depth', 1e-3), ('rgb', 2e-3), ('acc', 1e-5)):
        assert_close(out[k][pick.cuda()].reshape(512, -1), want[-1][k].reshape(512, -1), tol, 'sweep ' + k)
    a, b = out['intensity'][pick.cuda()].reshape(-1).cpu(), want[-1]['intensity'].reshape(-1)
    assert float((a - b).abs().max()) <= 3e-4, 'sweep intensity'
    ws = want[-1]['semantic'].reshape(512, -1)
    top2 = ws.topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 2e-3 * float(ws.max())
    got_label = out['semantic'][pick.cuda()].reshape(512, -1).argmax(-1).cpu()
    assert int(clear.sum()) > 256
    assert torch.equal(got_label[clear], ws.argmax(-1)[clear]), 'semantic argmax'


def test_state_dict_keys_match_reference():
    from nerf_lidar_b200 import configs, models
    model = models.Model(configs.nuscenes_single())
    keys = set(model.state_dict())
    for k in ('nerf_mlp.encoder.embeddings', 'nerf_mlp.encoder.offsets', 'nerf_mlp.encoder.idx',
              'nerf_mlp.encoder.grid_sizes', 'nerf_mlp.density_layer.0.weight', 'nerf_mlp.lin_second_stage_1.bias',
              'nerf_mlp.rgb_layer.weight', 'nerf_mlp.sem_layer.2.weight', 'nerf_mlp.intensity_layer.0.bias',
              'prop_mlp_0.encoder.embeddings', 'prop_mlp_1.density_layer.2.bias'):
        assert k in keys
    assert sum(p.numel() for p in model.parameters()) == 77656777


def test_render_image_chunked_equals_single_pass():
    """models.render_image (chunks of config.render_chunk_size, replayed as CUDA graphs) on a full 32 x 1084
    LiDAR sweep (BASELINE configs[2]) against ONE eager Model.forward over all rays, and the [H, W] image
    layout; a second call with another train_frac reuses the graphs (anneal is a dynamic scalar).  Tolerance:
    north_star's 1e-3 for outputs behind the bf16 MLP -- the two schedules composite with different kernels
    (warp per ray below 32768 rays, thread per ray above), whose ulp-level differences in the proposal weights
    move sample points and are amplified by the bf16 rounding of the MLP operands."""
    from nerf_lidar_b200 import configs, models
    cfg = configs.nuscenes_single()
    model = models.Model(cfg).cuda()
    model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=31, table_std=0.3).items()}, strict=False)
    sweep = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_lidar_sweep(seed=31)).items()}
    n = sweep['origins'].shape[0]
    assert n == 32 * 1084
    for train_frac in (1.0, 0.3):
        out = models.render_image(model, None, sweep, False, cfg, train_frac=train_frac, image=False, verbose=False)
        model.eval()
        model.training = False
        with torch.no_grad():
            rend, _ = model(False, sweep, train_frac, True)
        for k in ('rgb', 'depth', 'semantic', 'intensity', 'acc', 'distance_median'):
            a, b = out[k].reshape(n, -1), rend[-1][k].reshape(n, -1)
            assert float((a - b).abs().max()) <= 1e-3 * float(b.abs().max() + 1e-30), (train_frac, k)
    assert len(model._render_graphs) == 2          # chunk shapes: 16384 rays (twice) and 1920 rays
    # new weights in the same model (load_state_dict / training between renders): the replayed graphs must see
    # them, incl. the NerfMLP's packed bf16 operand images that are refreshed outside the graph
    with torch.no_grad():
        for name, p in model.nerf_mlp.named_parameters():
            if p.ndim == 2 and 'embeddings' not in name:
                p.mul_(0.8)
    out = models.render_image(model, None, sweep, False, cfg, image=False, verbose=False)
    with torch.no_grad():
        rend, _ = model(False, sweep, 1.0, True)
    assert len(model._render_graphs) == 2
    for k in ('rgb', 'semantic', 'intensity'):
        a, b = out[k].reshape(n, -1), rend[-1][k].reshape(n, -1)
        assert float((a - b).abs().max()) <= 1e-3 * float(b.abs().max() + 1e-30), ('new weights', k)
    img = {k: v[:60 * 40].reshape(60, 40, *v.shape[1:]) for k, v in sweep.items()}
    out = models.render_image(model, None, img, False, cfg, image=True, verbose=False)
    assert out['rgb'].shape == (60, 40, 3) and out['depth'].shape[:2] == (60, 40)
