"""Host-side operators of the zipnerf hot path: thin torch.autograd wrappers that
allocate outputs and call the C ABI in libnlb200.so (include/nlb200.h).

Reference functions replaced (Z/ = NeRF_LiDAR/zipnerf/):
  resample_level   Z/internal/models.py:320-372 (stepfun.max_dilate_weights,
                   stepfun.sample_intervals, math.sorted_interp, s_to_t)
  sorted_interp    Z/internal/math.py:89-108
  prop_level       render.cast_rays + MLP.predict_density + softplus for PropMLP
  nerf_encode      render.cast_rays + contract + GridEncoder + erf-weighted mean
  composite        render.compute_alpha_weights + render.volumetric_rendering
There is no CPU path: every function requires CUDA tensors."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
from torch.autograd import Function

from . import _lib
from ._lib import timed, NlbBf16SumJob, NlbSumTerm, NlbScaleJob, NlbRayGrads, NlbObjGrads, NlbObjMlp, NlbLossesIn, NlbNerfMlpWeights, NlbNerfMlpSaved, NlbNerfMlpGradIn, NlbNerfMlpGradOut, NlbCompositeGrad, NlbCompositeIn, NlbCompositeOut, NlbRays, NlbTable, check, f32, load, ptr, stream

EPS = float(torch.finfo(torch.float32).eps)
_u_cache: Dict[Tuple, torch.Tensor] = {}


def _u_base(num_samples: int, rand: bool, device) -> Tuple[torch.Tensor, float]:
    """The linspace of stepfun.sample (Z/internal/stepfun.py:199-216), computed by
    torch on the host exactly as the reference does, cached on the device."""
    if num_samples <= 1:
        raise RuntimeError(f'num_samples must be > 1, is {num_samples}.')
    key = (num_samples, bool(rand), str(device))
    if key not in _u_cache:
        if not rand:
            pad = 1 / (2 * num_samples)
            u = torch.linspace(pad, 1. - pad - EPS, num_samples)
        else:
            u_max = EPS + (1 - EPS) / num_samples
            u = torch.linspace(0, 1 - u_max, num_samples)
        _u_cache[key] = u.to(device)
    u_max = EPS + (1 - EPS) / num_samples
    max_jitter = (1 - u_max) / (num_samples - 1) - EPS
    return _u_cache[key], max_jitter


@torch.no_grad()
def resample_level(sdist: Optional[torch.Tensor], weights: Optional[torch.Tensor], near: torch.Tensor,
                   far: torch.Tensor, num_samples: int, dilate: bool, dilation: float, anneal: float,
                   jitter: Optional[torch.Tensor], rand: bool, lam: float = -1.5, resample_padding: float = 0.0,
                   return_index: bool = False):
    """One level of interval resampling.  sdist=None means the initial [0,1]
    interval with weight 1 (models.py:296-300).  Returns (sdist, tdist[, idx])."""
    near, far = f32(near).reshape(-1), f32(far).reshape(-1)
    N = near.shape[0]
    dev = near.device
    if sdist is None:
        n_in, sd, w = 1, None, None
    else:
        sd, w = f32(sdist), f32(weights)
        n_in = w.shape[-1]
    u_base, max_jitter = _u_base(num_samples, rand, dev)
    jit = f32(jitter).reshape(-1) if (rand and jitter is not None) else None
    if rand and jit is None:
        raise RuntimeError('resample_level: rand=True needs the per-ray jitter tensor')
    s_out = torch.empty(N, num_samples + 1, device=dev, dtype=torch.float32)
    t_out = torch.empty_like(s_out)
    idx = torch.empty(N, num_samples, device=dev, dtype=torch.int32) if return_index else None
    with torch.cuda.device(dev):
        with timed('resample'):
            check(load().nlb_resample(ptr(sd), ptr(w), n_in, int(dilate), float(dilation), float(anneal),
                                      float(resample_padding), ptr(u_base), ptr(jit), float(max_jitter), ptr(near),
                                      ptr(far), float(lam), num_samples, N, ptr(s_out), ptr(t_out), ptr(idx), stream()))
    return (s_out, t_out, idx) if return_index else (s_out, t_out)


@torch.no_grad()
def sorted_interp(x: torch.Tensor, xp: torch.Tensor, fp: torch.Tensor, return_index: bool = False):
    """math.sorted_interp for 2-D [N, n] inputs."""
    x, xp, fp = f32(x), f32(xp), f32(fp)
    N, nx = x.shape
    out = torch.empty_like(x)
    idx = torch.empty(N, nx, device=x.device, dtype=torch.int32) if return_index else None
    with torch.cuda.device(x.device):
        check(load().nlb_sorted_interp(ptr(x), ptr(xp), ptr(fp), N, nx, xp.shape[1], ptr(out), ptr(idx), stream()))
    return (out, idx) if return_index else out


class RayBundle:
    """Contiguous fp32 device views of the ray fields the kernels read."""

    def __init__(self, batch: Dict[str, torch.Tensor]):
        self.origins = f32(batch['origins'])
        self.directions = f32(batch['directions'])
        self.radii = f32(batch['radii']).reshape(-1)
        self.base_x = f32(batch['base_x'])
        self.base_y = f32(batch['base_y'])
        self.N = self.origins.shape[0]
        self.device = self.origins.device

    def desc(self, tdist: torch.Tensor, deg_noise: Optional[torch.Tensor], std_scale: float,
             points_cache: Optional[torch.Tensor] = None, points_mode: int = 0) -> NlbRays:
        """points_cache [7, N*S, 4] fp32 + mode 1 (write, training forward) / 2 (read, backward): the
        backward kernels reuse the forward's sample points instead of regenerating them."""
        S = tdist.shape[1] - 1
        return NlbRays(ptr(tdist), ptr(self.origins), ptr(self.directions), ptr(self.radii), ptr(self.base_x),
                       ptr(self.base_y), ptr(deg_noise), self.N, S, float(std_scale), ptr(points_cache),
                       int(points_mode) if points_cache is not None else 0)

    def new_points_cache(self, S: int) -> torch.Tensor:
        return torch.empty(7, self.N * S, 4, device=self.device, dtype=torch.float32)


@torch.no_grad()
def sample_points(tdist, deg_noise, rays: RayBundle, std_scale: float = 0.35) -> torch.Tensor:
    """Parity probe: [N,S,7,4] grid-space points (x,y,z in [0,1], std) of the fused kernels."""
    tdist = f32(tdist)
    pts = torch.empty(rays.N, tdist.shape[1] - 1, 7, 4, device=rays.device, dtype=torch.float32)
    with torch.cuda.device(rays.device):
        check(load().nlb_sample_points(C.byref(rays.desc(tdist, deg_noise, std_scale)), ptr(pts), stream()))
    return pts


def host_offsets(encoder):
    """Host copy (ctypes int32 array) of the encoder's level offsets, cached on the module."""
    cached = getattr(encoder, '_nlb_offsets_host', None)
    if cached is None:
        offs = encoder.offsets.tolist()
        cached = (C.c_int32 * len(offs))(*offs)
        encoder._nlb_offsets_host = cached
    return cached


def _table_desc(encoder, embeddings: torch.Tensor) -> NlbTable:
    return NlbTable(ptr(embeddings), ptr(encoder.offsets), ptr(encoder.grid_sizes), encoder.num_levels,
                    encoder.level_dim, int(encoder.base_resolution), float(math.log2(encoder.per_level_scale)),
                    host_offsets(encoder))


def _grad_buffer(param: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    """Gradient accumulation target for a table.  When the trainer has attached a
    persistent, pre-zeroed `_nlb_grad` buffer to the parameter the kernels add into
    it directly (no 240 MB memset + copy per step); otherwise a fresh zero tensor is
    returned through autograd."""
    buf = getattr(param, '_nlb_grad', None)
    if buf is not None:
        return buf, True
    return torch.zeros_like(param), False


def _ray_grad_buffers(rays: RayBundle):
    """Zeroed [4,N,3] accumulation target + its descriptor: gradients w.r.t. origins / directions / base_x /
    base_y of one level (pose-refinement window, Z/train.py:200-221)."""
    g = torch.zeros(4, rays.N, 3, device=rays.device, dtype=torch.float32)
    return g, NlbRayGrads(ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]))


class _PropLevel(Function):
    @staticmethod
    def forward(ctx, tdist, deg_noise, embeddings, W0, b0, W1, b1, rays: RayBundle, encoder, std_scale, emb_param,
                origins, directions, base_x, base_y):
        # the last four are rays.origins ... rays.base_y again, as differentiable inputs: they need gradients only
        # while poses are refined
        ctx.ray_grad = any(ctx.needs_input_grad[11:15])
        N, S = rays.N, tdist.shape[1] - 1
        density = torch.empty(N, S, device=rays.device, dtype=torch.float32)
        need_grad = any(ctx.needs_input_grad)
        feats = torch.empty(N * S, encoder.num_levels, device=rays.device, dtype=torch.float32) if need_grad else None
        pts = rays.new_points_cache(S) if need_grad else None
        W0c, b0c, W1c, b1c = f32(W0), f32(b0), f32(W1).reshape(-1), f32(b1)
        with torch.cuda.device(rays.device):
            with timed(f'prop{encoder.num_levels}_fwd'):
                check(load().nlb_prop_forward(C.byref(rays.desc(tdist, deg_noise, std_scale, pts, 1)),
                                              C.byref(_table_desc(encoder, embeddings)), ptr(W0c), ptr(b0c), ptr(W1c),
                                              ptr(b1c), ptr(density), ptr(feats), stream()))
        ctx.save_for_backward(tdist, deg_noise, embeddings, W0c, b0c, W1c, b1c, feats, pts)
        ctx.rays, ctx.encoder, ctx.std_scale, ctx.emb_param = rays, encoder, std_scale, emb_param
        ctx.wparams = (W0, b0, W1, b1)     # the Parameter objects (saved tensors lose their attributes)
        return density

    @staticmethod
    def backward(ctx, g_density):
        tdist, deg_noise, embeddings, W0, b0, W1, b1, feats, pts = ctx.saved_tensors
        rays, encoder = ctx.rays, ctx.encoder
        g_emb, in_place = _grad_buffer(ctx.emb_param if ctx.emb_param is not None else embeddings)
        # the trainer's persistent gradient buffers (views of its flat gradient) are added into directly: no zero
        # fills, no AccumulateGrad adds; without a trainer the gradients are returned through autograd
        bufs = [getattr(prm, '_nlb_grad', None) for prm in ctx.wparams]
        w_in_place = all(b is not None and b.is_contiguous() for b in bufs)
        if w_in_place:
            gW0, gb0, gW1, gb1 = bufs
        else:
            gW0, gb0 = torch.zeros_like(W0), torch.zeros_like(b0)
            gW1, gb1 = torch.zeros_like(W1), torch.zeros_like(b1)
        g_density = f32(g_density)
        tab = _table_desc(encoder, embeddings)
        ws_bytes = load().nlb_prop_backward_workspace_bytes(rays.N, tdist.shape[1] - 1, C.byref(tab))
        ws = torch.empty(ws_bytes // 4, device=rays.device, dtype=torch.float32)
        with torch.cuda.device(rays.device):
            with timed(f'prop{encoder.num_levels}_bwd'):
                check(load().nlb_prop_backward(C.byref(rays.desc(tdist, deg_noise, ctx.std_scale, pts, 2)),
                                               C.byref(tab), ptr(W0), ptr(b0), ptr(W1),
                                               ptr(b1), ptr(feats), ptr(g_density), ptr(g_emb), ptr(gW0), ptr(gb0),
                                               ptr(gW1), ptr(gb1), ptr(ws), stream()))
            g_rays = (None,) * 4
            if ctx.ray_grad:
                g, desc = _ray_grad_buffers(rays)
                with timed(f'prop{encoder.num_levels}_input_bwd'):
                    check(load().nlb_prop_input_backward(C.byref(rays.desc(tdist, deg_noise, ctx.std_scale)),
                                                         C.byref(tab), ptr(ws), C.byref(desc), stream()))
                g_rays = tuple(g.unbind(0))
        if w_in_place:
            return (None, None, None if in_place else g_emb, None, None, None, None, None, None, None, None, *g_rays)
        return (None, None, None if in_place else g_emb, gW0, gb0, gW1.reshape(1, -1), gb1, None, None, None, None,
                *g_rays)


def prop_level(tdist, deg_noise, mlp, rays: RayBundle, std_scale: float) -> torch.Tensor:
    """Proposal density[N,S] for one level (cast_rays + encode + PropMLP + softplus)."""
    enc = mlp.encoder
    l0, l2 = mlp.density_layer[0], mlp.density_layer[2]
    return _PropLevel.apply(tdist, deg_noise, enc.embeddings, l0.weight, l0.bias, l2.weight, l2.bias, rays, enc,
                            std_scale, enc.embeddings, rays.origins, rays.directions, rays.base_x, rays.base_y)


class _NerfEncode(Function):
    @staticmethod
    def forward(ctx, tdist, deg_noise, embeddings, rays: RayBundle, encoder, std_scale, emb_param,
                origins, directions, base_x, base_y):
        ctx.ray_grad = any(ctx.needs_input_grad[7:11])
        N, S = rays.N, tdist.shape[1] - 1
        feats = torch.empty(N * S, encoder.output_dim, device=rays.device, dtype=torch.float32)
        pts = rays.new_points_cache(S) if any(ctx.needs_input_grad) else None
        with torch.cuda.device(rays.device):
            with timed('nerf_encode_fwd'):
                check(load().nlb_encode_forward(C.byref(rays.desc(tdist, deg_noise, std_scale, pts, 1)),
                                                C.byref(_table_desc(encoder, embeddings)), ptr(feats), stream()))
        ctx.save_for_backward(tdist, deg_noise, embeddings, pts)
        ctx.rays, ctx.encoder, ctx.std_scale, ctx.emb_param = rays, encoder, std_scale, emb_param
        return feats

    @staticmethod
    def backward(ctx, g_feats):
        tdist, deg_noise, embeddings, pts = ctx.saved_tensors
        rays, encoder = ctx.rays, ctx.encoder
        g_emb, in_place = _grad_buffer(ctx.emb_param if ctx.emb_param is not None else embeddings)
        g_feats = f32(g_feats)
        tab = _table_desc(encoder, embeddings)
        ws_bytes = load().nlb_encode_backward_workspace_bytes(C.byref(tab))
        ws = torch.empty(ws_bytes // 4, device=rays.device, dtype=torch.float32) if ws_bytes else None
        with torch.cuda.device(rays.device):
            with timed('nerf_encode_bwd'):
                check(load().nlb_encode_backward(C.byref(rays.desc(tdist, deg_noise, ctx.std_scale, pts, 2)),
                                                 C.byref(tab), ptr(g_feats), ptr(g_emb), ptr(ws), stream()))
            g_rays = (None,) * 4
            if ctx.ray_grad:
                g, desc = _ray_grad_buffers(rays)
                with timed('nerf_encode_input_bwd'):
                    check(load().nlb_encode_input_backward(C.byref(rays.desc(tdist, deg_noise, ctx.std_scale)),
                                                           C.byref(tab), ptr(g_feats), C.byref(desc), stream()))
                g_rays = tuple(g.unbind(0))
        return (None, None, None if in_place else g_emb, None, None, None, None, *g_rays)


def nerf_encode(tdist, deg_noise, encoder, rays: RayBundle, std_scale: float) -> torch.Tensor:
    """features[N*S, L*C] of the NeRF level."""
    return _NerfEncode.apply(tdist, deg_noise, encoder.embeddings, rays, encoder, std_scale, encoder.embeddings,
                             rays.origins, rays.directions, rays.base_x, rays.base_y)


class _Composite(Function):
    @staticmethod
    def forward(ctx, density, rgb, semantic, intensity, tdist, directions, far, bg, opaque, extras):
        ctx.set_materialize_grads(False)  # outputs nobody differentiates arrive as None, not as zero fills
        N, S = density.shape
        dev = density.device
        density = f32(density)
        rgb_c = f32(rgb) if rgb is not None else None
        sem_c = f32(semantic) if semantic is not None else None
        int_c = f32(intensity).reshape(N, S) if intensity is not None else None
        K = sem_c.shape[-1] if sem_c is not None else 0
        far_c = f32(far).reshape(-1)
        new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        weights, o_rgb, depth, acc = new(N, S), new(N, 3), new(N), new(N)
        o_sem = new(N, K) if sem_c is not None else None
        o_int = new(N) if int_c is not None else None
        dmean = new(N) if extras else None
        dpct = new(N, 3) if extras else None
        cin = NlbCompositeIn(ptr(density), ptr(tdist), ptr(directions), ptr(rgb_c), ptr(sem_c), ptr(int_c),
                             ptr(far_c), N, S, K, float(bg), int(opaque), int(extras))
        cout = NlbCompositeOut(ptr(weights), ptr(o_rgb), ptr(depth), ptr(acc), ptr(o_sem), ptr(o_int), ptr(dmean),
                               ptr(dpct))
        with torch.cuda.device(dev):
            with timed(f'composite{S}_fwd'):
                check(load().nlb_composite_forward(C.byref(cin), C.byref(cout), stream()))
        ctx.save_for_backward(density, rgb_c, sem_c, int_c, tdist, directions, far_c, weights)
        ctx.cfg = (N, S, K, float(bg), int(opaque))
        ctx.int_shape = None if intensity is None else intensity.shape
        outs = (weights, o_rgb, depth, acc, o_sem, o_int, dmean, dpct)
        ctx.mark_non_differentiable(*[t for t in (dmean, dpct) if t is not None])
        return outs

    @staticmethod
    def backward(ctx, g_w, g_rgb, g_depth, g_acc, g_sem, g_int, _gm, _gp):
        density, rgb, sem, inten, tdist, directions, far, weights = ctx.saved_tensors
        N, S, K, bg, opaque = ctx.cfg
        dev = density.device
        c = lambda t: None if t is None else f32(t)
        g_w, g_rgb, g_depth, g_acc, g_sem, g_int = c(g_w), c(g_rgb), c(g_depth), c(g_acc), c(g_sem), c(g_int)
        cin = NlbCompositeIn(ptr(density), ptr(tdist), ptr(directions), ptr(rgb), ptr(sem), ptr(inten), ptr(far),
                             N, S, K, bg, opaque, 0)
        cg = NlbCompositeGrad(ptr(g_w), ptr(g_rgb), ptr(g_depth), ptr(g_acc), ptr(g_sem), ptr(g_int))
        gd = torch.empty(N, S, device=dev, dtype=torch.float32)
        need = ctx.needs_input_grad
        g_rgb_s = torch.empty(N, S, 3, device=dev, dtype=torch.float32) if (rgb is not None and need[1] and g_rgb is not None) else None
        g_sem_s = torch.empty(N, S, K, device=dev, dtype=torch.float32) if (sem is not None and need[2] and g_sem is not None) else None
        g_int_s = torch.empty(N, S, device=dev, dtype=torch.float32) if (inten is not None and need[3] and g_int is not None) else None
        with torch.cuda.device(dev):
            with timed(f'composite{S}_bwd'):
                check(load().nlb_composite_backward(C.byref(cin), ptr(weights), C.byref(cg), ptr(gd), ptr(g_rgb_s),
                                                    ptr(g_sem_s), ptr(g_int_s), stream()))
        if g_int_s is not None and ctx.int_shape is not None:
            g_int_s = g_int_s.reshape(ctx.int_shape)
        g_dir = None
        if need[5]:
            # alpha = 1 - exp(-density * delta * |d|) (render.py:170-189): the norm scales every density of the ray,
            # so dL/d|d| = sum_s gd_s density_s / |d| and d|d|/dd = d / |d|   (pose-refinement window only)
            g_dir = directions * ((gd * density).sum(-1) / (directions * directions).sum(-1))[:, None]
        return gd, g_rgb_s, g_sem_s, g_int_s, None, g_dir, None, None, None, None


def composite(density, tdist, directions, far, rgb=None, semantic=None, intensity=None, bg: float = 1.0,
              opaque_background: bool = True, compute_extras: bool = True) -> Dict[str, Optional[torch.Tensor]]:
    """compute_alpha_weights + volumetric_rendering for one level.  Returns a dict
    with weights[N,S], rgb[N,3], depth[N], acc[N], semantic[N,K], intensity[N],
    distance_mean[N], distance_percentiles[N,3]."""
    w, o_rgb, depth, acc, o_sem, o_int, dmean, dpct = _Composite.apply(
        density, rgb, semantic, intensity, f32(tdist), f32(directions), far, bg, opaque_background, compute_extras)
    return dict(weights=w, rgb=o_rgb, depth=depth, acc=acc, semantic=o_sem, intensity=o_int, distance_mean=dmean,
                distance_percentiles=dpct)


# ----------------------------------------------------------------------------- NeRF MLP (tcgen05)
_NERF_WEIGHT_FIELDS = (('W_d0', 'density_layer.0.weight'), ('b_d0', 'density_layer.0.bias'),
                       ('W_d2', 'density_layer.2.weight'), ('b_d2', 'density_layer.2.bias'),
                       ('W_s0', 'sem_layer.0.weight'), ('b_s0', 'sem_layer.0.bias'),
                       ('W_s2', 'sem_layer.2.weight'), ('b_s2', 'sem_layer.2.bias'),
                       ('W_i0', 'intensity_layer.0.weight'), ('b_i0', 'intensity_layer.0.bias'),
                       ('W_i2', 'intensity_layer.2.weight'), ('b_i2', 'intensity_layer.2.bias'),
                       ('W_v0', 'lin_second_stage_0.weight'), ('b_v0', 'lin_second_stage_0.bias'),
                       ('W_v1', 'lin_second_stage_1.weight'), ('b_v1', 'lin_second_stage_1.bias'),
                       ('W_rgb', 'rgb_layer.weight'), ('b_rgb', 'rgb_layer.bias'))


def _mlp_tensors(mlp):
    """Dense-layer parameters in nlb_nerf_mlp_weights_t order; None for the intensity head of a network built
    without it (Config.use_intensity=False: inference only, see nlb_nerf_mlp_pack)."""
    params = dict(mlp.named_parameters())
    return [params.get(name) if name.startswith('intensity_layer') else params[name] for _, name in _NERF_WEIGHT_FIELDS]


def _mlp_tensors_train(mlp):
    """As above for the training kernels, which are compiled with the intensity head: a network built without it
    (Config.use_intensity=False, e.g. with the dynamic-object branch) gets a frozen all-zero head -- plain
    tensors, not parameters, not in the state dict -- whose output is discarded and whose gradients go to scratch."""
    tensors = _mlp_tensors(mlp)
    if all(t is not None for t in tensors):
        return tensors
    dummy = mlp.__dict__.get('_nlb_zero_intensity_head')
    dev = tensors[0].device
    if dummy is None or dummy[0].device != dev:
        z = lambda *s: torch.zeros(*s, device=dev)
        dummy = [z(64, mlp.bottleneck_width), z(64), z(1, 64), z(1)]
        mlp.__dict__['_nlb_zero_intensity_head'] = dummy
    it = iter(dummy)
    return [t if t is not None else next(it) for t in tensors]


@torch.no_grad()
def nerf_mlp_pack(mlp, transposed: bool = False) -> torch.Tensor:
    """Packs the NerfMLP dense layers into the bf16 operand-block blob the fused
    kernels stream (nlb_nerf_mlp_pack / _pack_transposed).  Cached on the module and
    refreshed when any parameter has been modified in place (optimizer step,
    load_state_dict)."""
    tensors = _mlp_tensors_train(mlp) if transposed else _mlp_tensors(mlp)
    version = tuple((t.data_ptr(), t._version) for t in tensors if t is not None)
    key = '_nlb_packed_t' if transposed else '_nlb_packed'
    cache = getattr(mlp, key, None)
    dirty = getattr(mlp, '_nlb_dirty', None)  # set by Trainer.optimizer_step (raw-pointer updates bypass _version)
    if dirty is None:
        mlp._nlb_dirty = dirty = set()
    if dirty is True:
        mlp._nlb_dirty = dirty = {'_nlb_packed', '_nlb_packed_t'}
    if cache is not None and cache[0] == version and key not in dirty:
        return cache[1]
    dirty.discard(key)
    dev = tensors[0].device
    lib = load()
    nbytes = lib.nlb_nerf_mlp_packed_transposed_bytes() if transposed else lib.nlb_nerf_mlp_packed_bytes()
    blob = cache[1] if cache is not None else torch.empty(nbytes, dtype=torch.uint8, device=dev)
    keep = [None if t is None else f32(t.detach()) for t in tensors]
    w = NlbNerfMlpWeights(*[ptr(t) for t in keep])
    with torch.cuda.device(dev):
        fn = lib.nlb_nerf_mlp_pack_transposed if transposed else lib.nlb_nerf_mlp_pack
        check(fn(C.byref(w), ptr(blob), stream()))
    setattr(mlp, key, (version, blob))
    return blob


def _mlp_forward_raw(mlp, features, viewdirs, S, save: bool):
    M, N = features.shape[0], viewdirs.shape[0]
    dev = features.device
    blob = nerf_mlp_pack(mlp)
    new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    density, rgb, sem, inten = new(N, S), new(N, S, 3), new(N, S, 19), new(N, S, 1)
    if getattr(mlp, 'intensity_layer', None) is None:
        inten = None            # (the packed head is all zeros; nothing is written)
    saved = None
    sv = None
    if save:
        bf = lambda c: torch.empty(M, c, device=dev, dtype=torch.bfloat16)
        saved = dict(h0=bf(64), x=bf(256), g=bf(128), h1=bf(256), h2=bf(256), f0=bf(64))
        sv = NlbNerfMlpSaved(*[ptr(saved[k]) for k in ('h0', 'x', 'g', 'h1', 'h2', 'f0')])
    with torch.cuda.device(dev):
        with timed('nerf_mlp_fwd'):
            check(load().nlb_nerf_mlp_forward(ptr(features), ptr(viewdirs), M, S, ptr(blob), ptr(density), ptr(rgb),
                                              ptr(sem), ptr(inten), C.byref(sv) if sv is not None else None, stream()))
    return density, rgb, sem, inten, saved


@torch.no_grad()
def nerf_mlp_forward(mlp, features: torch.Tensor, viewdirs: torch.Tensor, S: int) -> Dict[str, torch.Tensor]:
    """Fused tcgen05 NerfMLP forward (inference): features[N*S,40] -> density[N,S],
    rgb[N,S,3], semantic[N,S,19], intensity[N,S,1]."""
    density, rgb, sem, inten, _ = _mlp_forward_raw(mlp, f32(features), f32(viewdirs), S, False)
    return dict(density=density, rgb=rgb, semantic=sem, intensity=inten)


@torch.no_grad()
def _bf16_rows_ptr(x: torch.Tensor):
    """Pointer of a bf16 matrix whose rows are dense (column slices of a wider buffer are fine)."""
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 2 or x.stride(1) != 1:
        raise RuntimeError('nerf_lidar_b200: expected a CUDA bf16 matrix with unit column stride')
    return C.c_void_p(x.data_ptr())


def colsum_bf16(x: torch.Tensor) -> torch.Tensor:
    """Column sums of a bf16 matrix [M, cols] (dense rows, any row stride) -> fp32 [cols]."""
    M, cols = x.shape
    out = torch.empty(cols, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        with timed('mlp_bias_grad'):
            check(load().nlb_colsum_bf16(_bf16_rows_ptr(x), M, cols, x.stride(0), ptr(out), stream()))
    return out


@torch.no_grad()
def group_sum_bf16(x: torch.Tensor, S: int) -> torch.Tensor:
    """Sums over groups of S consecutive rows of a bf16 [G*S, cols] (dense rows) -> fp32 [G, cols]."""
    M, cols = x.shape
    out = torch.empty(M // S, cols, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        with timed('mlp_ray_sum'):
            check(load().nlb_group_sum_bf16(_bf16_rows_ptr(x), M // S, S, cols, x.stride(0), ptr(out), stream()))
    return out


@torch.no_grad()
def bf16_sums(jobs) -> None:
    """[(x bf16 [M, cols] with dense rows, group, out fp32)]: group = 0 -> out[cols] = column sums, group = S ->
    out[M / S, cols] = sums over S consecutive rows; all jobs in ONE launch (nlb_bf16_sums)."""
    arr = (NlbBf16SumJob * len(jobs))()
    for q, (x, group, out) in zip(arr, jobs):
        M, cols = x.shape
        q.x, q.rows, q.cols, q.ld, q.group, q.out = _bf16_rows_ptr(x).value, M, cols, x.stride(0), int(group), ptr(out)
    with torch.cuda.device(jobs[0][0].device):
        with timed('mlp_grad_sums'):
            check(load().nlb_bf16_sums(arr, len(jobs), stream()))


class _NerfMLP(Function):
    """Training path of the NerfMLP, all on tcgen05: fused forward that saves bf16 activations, fused
    data-gradient chain, and the weight gradients dW = dZ^T A as MN-major UMMA products (csrc/nerf_wgrad.cu)
    added straight into the parameters' gradient buffers."""

    @staticmethod
    def forward(ctx, features, viewdirs, mlp, S, *weights):
        ctx.set_materialize_grads(False)
        features, viewdirs = f32(features), f32(viewdirs)
        density, rgb, sem, inten, saved = _mlp_forward_raw(mlp, features, viewdirs, S, True)
        ctx.mlp, ctx.S = mlp, S
        ctx.save_for_backward(features, viewdirs, density, rgb, sem, *[saved[k] for k in ('h0', 'x', 'g', 'h1', 'h2', 'f0')])
        return density, rgb, sem, inten

    @staticmethod
    def backward(ctx, g_density, g_rgb, g_sem, g_int):
        features, viewdirs, density, rgb, sem, h0, x, g, h1, h2, f0 = ctx.saved_tensors
        mlp, S = ctx.mlp, ctx.S
        M, N = features.shape[0], viewdirs.shape[0]
        dev = features.device
        c = lambda t: None if t is None else f32(t)
        g_density, g_rgb, g_sem, g_int = c(g_density), c(g_rgb), c(g_sem), c(g_int)
        blob_t = nerf_mlp_pack(mlp, transposed=True)
        bf = lambda cols: torch.empty(M, cols, device=dev, dtype=torch.bfloat16)
        d_rgb, d_hs1, d_x, d_h0 = bf(16), bf(32), bf(256), bf(64)
        dcat = bf(640)   # d_g | d_v0 | d_v1: column slices of one buffer (the three share the operand x)
        d_g, d_v0, d_v1 = dcat[:, :128], dcat[:, 128:384], dcat[:, 384:]
        g_feat = torch.empty(M, 40, device=dev, dtype=torch.float32)
        gin = NlbNerfMlpGradIn(ptr(g_density), ptr(g_rgb), ptr(g_sem), ptr(g_int), ptr(density), ptr(rgb), ptr(sem))
        sv = NlbNerfMlpSaved(ptr(h0), ptr(x), ptr(g), ptr(h1), ptr(h2), ptr(f0))
        gout = NlbNerfMlpGradOut(ptr(d_rgb), d_v1.data_ptr(), d_v0.data_ptr(), ptr(d_hs1), d_g.data_ptr(), ptr(d_x),
                                 ptr(d_h0), 640, 640, 640)
        # gradient targets: the trainer's persistent buffers (views of its flat gradient, `param._nlb_grad`) are
        # added into directly; without a trainer the gradients are returned through autograd
        params = _mlp_tensors(mlp)                       # None for the intensity head of a network built without it
        shapes = [t.shape for t in _mlp_tensors_train(mlp)]
        in_place = all(getattr(p, '_nlb_grad', None) is not None for p in params if p is not None)
        targets = []
        for p, shp in zip(params, shapes):
            if p is None:
                targets.append(torch.zeros(shp, device=dev, dtype=torch.float32))      # scratch: discarded
            elif in_place:
                targets.append(p._nlb_grad)
            else:
                targets.append(torch.zeros(shp, device=dev, dtype=torch.float32))
        wg = NlbNerfMlpWeights(*[ptr(t) for t in targets])
        with torch.cuda.device(dev):
            with timed('nerf_mlp_bwd'):
                check(load().nlb_nerf_mlp_backward(C.byref(gin), C.byref(sv), M, ptr(blob_t), ptr(g_feat),
                                                   C.byref(gout), stream()))
            with timed('nerf_mlp_wgrad'):
                check(load().nlb_nerf_mlp_wgrad(C.byref(sv), C.byref(gout), M, C.byref(wg), stream()))
        # bias gradients and per-ray sums (the view-direction encoding is a per-ray constant): ONE bandwidth-bound
        # launch over the seven matrices (csrc/reduce.cu), folded into the gradient buffers by one small kernel
        cs_x, cs_g, cs_h0, cs_hs1, cs_rgb = torch.empty(496, device=dev, dtype=torch.float32).split([256, 128, 64, 32, 16])
        rs_v0, rs_v1 = torch.empty(2, M // S, 256, device=dev, dtype=torch.float32).unbind(0)
        bf16_sums([(d_x, 0, cs_x), (d_g, 0, cs_g), (d_h0, 0, cs_h0), (d_hs1, 0, cs_hs1), (d_rgb, 0, cs_rgb),
                   (d_v0, S, rs_v0), (d_v1, S, rs_v1)])
        with torch.cuda.device(dev):
            with timed('nerf_mlp_wgrad_finish'):
                check(load().nlb_nerf_mlp_wgrad_finish(ptr(rs_v0), ptr(rs_v1), ptr(viewdirs), N, ptr(cs_x), ptr(cs_g),
                                                       ptr(cs_h0), ptr(cs_hs1), ptr(cs_rgb), C.byref(wg), stream()))
        g_view = None
        if ctx.needs_input_grad[1]:
            # the encoded view direction enters both view layers (skip_layer_dir = 0, models.py:1010-1041): columns
            # [256,283) of W_v0 and [512,539) of W_v1; chain through coord.pos_enc with torch on the [N,3] tensor
            Wv0, Wv1 = f32(mlp.lin_second_stage_0.weight), f32(mlp.lin_second_stage_1.weight)
            nd = Wv0.shape[1] - mlp.bottleneck_width
            o1 = Wv1.shape[1] - nd
            g_enc = rs_v0 @ Wv0[:, mlp.bottleneck_width:] + rs_v1 @ Wv1[:, o1:]
            with torch.enable_grad():
                v = viewdirs.detach().requires_grad_(True)
                (g_view,) = torch.autograd.grad(mlp.dir_enc(v), v, g_enc)
        return (g_feat, g_view, None, None, *[None if (in_place or p is None) else t for p, t in zip(params, targets)])


def nerf_mlp_train(mlp, features: torch.Tensor, viewdirs: torch.Tensor, S: int) -> Dict[str, torch.Tensor]:
    density, rgb, sem, inten = _NerfMLP.apply(features, viewdirs, mlp, S, *_mlp_tensors(mlp))
    return dict(density=density, rgb=rgb, semantic=sem, intensity=inten)


# ----------------------------------------------------------------------------- per-ray regularisers
class _Distortion(Function):
    @staticmethod
    def forward(ctx, sdist, weights):
        sdist, weights = f32(sdist), f32(weights)
        N, S = weights.shape
        loss = torch.empty(N, device=weights.device, dtype=torch.float32)
        grad = torch.empty_like(weights)
        with torch.cuda.device(weights.device):
            with timed('distortion'):
                check(load().nlb_distortion_loss(ptr(sdist), ptr(weights), N, S, ptr(loss), ptr(grad), stream()))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, grad * g[:, None]


def distortion_per_ray(sdist: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """stepfun.lossfun_distortion(sdist, weights) -> [N]."""
    return _Distortion.apply(sdist, weights)


class _Interlevel(Function):
    @staticmethod
    def forward(ctx, c, w, cp, wp, pulse_width):
        c, w, cp, wp = f32(c), f32(w), f32(cp), f32(wp)
        N, Sc = w.shape
        Sp = wp.shape[1]
        loss = torch.empty(N, device=wp.device, dtype=torch.float32)
        grad = torch.empty_like(wp)
        with torch.cuda.device(wp.device):
            with timed('interlevel'):
                check(load().nlb_interlevel_loss(ptr(c), ptr(w), Sc, ptr(cp), ptr(wp), Sp, float(pulse_width), N,
                                                 ptr(loss), ptr(grad), stream()))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, None, None, grad * g[:, None], None


def interlevel_per_ray(c, w, cp, wp, pulse_width: float) -> torch.Tensor:
    """Sum over the proposal intervals of max(w_s - wp, 0)^2 / (wp + 1e-5) -> [N];
    (c, w) are the detached final-level histogram."""
    return _Interlevel.apply(c.detach(), w.detach(), cp.detach(), wp, pulse_width)


# ----------------------------------------------------------------------------- supervision losses
def _render_losses_launch(rgb, depth, semantic, intensity, batch, cfg):
    """One call of csrc/render_losses.cu: (losses[6], scales[6], g_rgb, g_depth, g_sem, g_int, g_dsmo, g_ssmo) -- the six
    loss values [data, depth, sem, int, d_smo, s_smo] and their gradients up to the per-loss scale factors."""
    rgb, depth = f32(rgb), f32(depth).reshape(-1)
    N = depth.shape[0]
    dev = depth.device
    sem = f32(semantic) if semantic is not None else None
    inten = f32(intensity).reshape(-1) if intensity is not None else None
    K = sem.shape[-1] if sem is not None else 0
    t = {k: f32(batch[k]).reshape(-1) if k != 'rgb' else f32(batch[k][..., :3])
         for k in ('rgb', 'depth', 'semantic', 'intensity', 'patch_mask', 'lidar_mask') if k in batch}
    # Z/train.py:286-289,307: the loss applies where the dataset mask is non-zero; instance_obj clears it
    valid = None
    if not cfg.get('instance_obj', False) and batch.get('mask') is not None:
        valid = f32(batch['mask']).reshape(-1)
    new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    losses, scales = new(6), new(6)
    g_rgb, g_depth = new(N, 3), new(N)
    g_sem = new(N, K) if sem is not None else None
    g_int = new(N) if inten is not None else None
    num_patch = int(cfg['num_patch'])
    g_dsmo = g_ssmo = None
    if num_patch > 0:   # accumulated with atomics: one zero fill for both
        smo = torch.zeros(N * (1 + K), device=dev)
        g_dsmo = smo[:N]
        g_ssmo = smo[N:].view(N, K) if sem is not None else None
    ws = new(load().nlb_render_losses_workspace_bytes() // 4)
    lin = NlbLossesIn(ptr(rgb), ptr(depth), ptr(sem), ptr(inten), ptr(t['rgb']), ptr(t['depth']),
                      ptr(t.get('semantic')), ptr(t.get('intensity')), ptr(t['patch_mask']), ptr(t['lidar_mask']),
                      ptr(valid), N, K, num_patch, int(cfg['patch_size']), int(cfg['lidar_supervision']),
                      int(cfg['only_lidar_supervision']), int(cfg['charb']), float(cfg['charb_padding']),
                      float(cfg['depth_mult']), float(cfg['sem_mult']), float(cfg['int_mult']),
                      float(cfg['smooth_mult']), 0., 0.)
    with torch.cuda.device(dev):
        with timed('render_losses'):
            check(load().nlb_render_losses(C.byref(lin), ptr(losses), ptr(scales), ptr(g_rgb), ptr(g_depth),
                                           ptr(g_sem), ptr(g_int), ptr(g_dsmo), ptr(g_ssmo), ptr(ws), stream()))
    return losses, scales, g_rgb, g_depth, g_sem, g_int, g_dsmo, g_ssmo


class _RenderLosses(Function):
    """data / depth / sem / int / d_smo / s_smo of Z/train.py:283-455 in four launches
    (csrc/render_losses.cu); returns the six loss values as one tensor."""

    @staticmethod
    def forward(ctx, rgb, depth, semantic, intensity, batch, cfg):
        losses, scales, g_rgb, g_depth, g_sem, g_int, g_dsmo, g_ssmo = _render_losses_launch(
            rgb, depth, semantic, intensity, batch, cfg)
        ctx.save_for_backward(scales, g_rgb, g_depth, g_sem, g_int, g_dsmo, g_ssmo)
        ctx.int_shape = None if intensity is None else intensity.shape
        ctx.depth_shape = depth.shape
        # six 0-dim outputs: indexing ONE output tensor costs a select node each -- a zero fill and a copy per
        # term in the backward pass
        return tuple(losses.unbind(0))

    @staticmethod
    def backward(ctx, *gos):
        scales, g_rgb, g_depth, g_sem, g_int, g_dsmo, g_ssmo = ctx.saved_tensors
        zero = scales.new_zeros(()) if any(g is None for g in gos) else None
        go = torch.stack([g if g is not None else zero for g in gos])
        w = go * scales
        o_rgb = g_rgb * w[0]
        o_depth = g_depth * w[1]
        if g_dsmo is not None:
            o_depth = torch.addcmul(o_depth, g_dsmo, w[4])
        o_sem = None
        if g_sem is not None:
            o_sem = g_sem * w[2]
            if g_ssmo is not None:
                o_sem = torch.addcmul(o_sem, g_ssmo, w[5])
        o_int = (g_int * w[3]).reshape(ctx.int_shape) if g_int is not None else None
        return o_rgb, o_depth.reshape(ctx.depth_shape), o_sem, o_int, None, None


def render_losses(rendering: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], cfg: Dict) -> torch.Tensor:
    """[data, depth, sem, int, d_smo, s_smo] for the final rendering (see _RenderLosses)."""
    return _RenderLosses.apply(rendering['rgb'], rendering['depth'], rendering.get('semantic'),
                               rendering.get('intensity'), batch, cfg)


# ----------------------------------------------------------------------------- loss assembly in two launches
def weighted_sums(terms, out: torch.Tensor) -> None:
    """nlb_weighted_sums: terms = [(x, w, coef, out_index)] with x a float32 CUDA tensor (every element is summed),
    w an optional weight tensor of the same size; x = an int names an output this call has already formed."""
    arr = (NlbSumTerm * len(terms))()
    for q, (x, w, coef, oi) in zip(arr, terms):
        if isinstance(x, int):
            q.x, q.w, q.n = None, None, x
        else:
            q.x, q.w, q.n = ptr(x), ptr(w), x.numel()
        q.coef, q.out_index = float(coef), int(oi)
    with torch.cuda.device(out.device):
        with timed('loss_sums'):
            check(load().nlb_weighted_sums(arr, len(terms), ptr(out), out.numel(), stream()))


def scale_tensors(jobs) -> None:
    """nlb_scale_tensors: jobs = [(src, src2, dst, g, s, s2, coef, coef2)]: dst = src * (coef g s) + src2 * (coef2 g s2);
    g / s / s2 one-element CUDA tensors or None (= 1)."""
    arr = (NlbScaleJob * len(jobs))()
    for q, (src, src2, dst, g, s1, s2, coef, coef2) in zip(arr, jobs):
        q.src, q.src2, q.dst, q.n = ptr(src), ptr(src2), ptr(dst), src.numel()
        q.g, q.s, q.s2 = ptr(g), ptr(s1), ptr(s2)
        q.coef, q.coef2 = float(coef), float(coef2)
    with torch.cuda.device(jobs[0][0].device):
        with timed('loss_seed'):
            check(load().nlb_scale_tensors(arr, len(jobs), stream()))


class _InterlevelTotal(Function):
    """mult * sum over the proposal levels of mean(anti-interlevel loss) (train_utils.py:134-172) as ONE scalar:
    a k_interlevel launch per level, one k_weighted_sums; the backward seeds every level's weights in one launch."""

    @staticmethod
    def forward(ctx, c, w, mult, pulse_widths, cps, *wps):
        c, w = f32(c), f32(w)
        N, Sc = w.shape
        dev = w.device
        out = torch.empty(1, device=dev, dtype=torch.float32)
        grads, terms, coefs = [], [], []
        for cp, wp, pw in zip(cps, wps, pulse_widths):
            cp, wp = f32(cp), f32(wp)
            Sp = wp.shape[1]
            loss = torch.empty(N, device=dev, dtype=torch.float32)
            grad = torch.empty_like(wp)
            with torch.cuda.device(dev):
                with timed('interlevel'):
                    check(load().nlb_interlevel_loss(ptr(c), ptr(w), Sc, ptr(cp), ptr(wp), Sp, float(pw), N,
                                                     ptr(loss), ptr(grad), stream()))
            coef = float(mult) / (N * Sp)
            grads.append(grad)
            coefs.append(coef)
            terms.append((loss, None, coef, 0))
        weighted_sums(terms, out)
        ctx.save_for_backward(*grads)
        ctx.coefs = coefs
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        g = f32(g)
        outs = [torch.empty_like(t) for t in ctx.saved_tensors]
        scale_tensors([(src, None, dst, g, None, None, coef, 0.) for src, dst, coef in zip(ctx.saved_tensors, outs, ctx.coefs)])
        return (None, None, None, None, None, *outs)


def interlevel_total(c, w, cps, wps, pulse_widths, mult: float) -> torch.Tensor:
    """config.anti_interlevel_loss_mult * sum_l mean(interlevel_l); (c, w) the detached final-level histogram."""
    return _InterlevelTotal.apply(c.detach(), w.detach(), float(mult), tuple(float(p) for p in pulse_widths),
                                  tuple(t.detach() for t in cps), *wps)


class _MainLoss(Function):
    """Every loss term that reaches the NeRF level -- the six supervision losses of csrc/render_losses.cu and the
    distortion loss -- with their sum, and the step's total, formed by ONE k_weighted_sums launch; the backward
    seeds rgb / depth / semantic / intensity / final-level weights in ONE k_scale_tensors launch.
    Returns (main, values[6] = [data, depth, sem, int, d_smo, s_smo], extras[3] = [distortion, total, hash_decay
    copy]); only `main` carries a gradient."""

    @staticmethod
    def forward(ctx, rgb, depth, semantic, intensity, weights, sdist, batch, cfg, used, dist_mult, prop_value,
                hash_decay, extra_values):
        ctx.set_materialize_grads(False)   # the two value-only outputs would each get a zero fill in the backward pass
        losses, scales, g_rgb, g_depth, g_sem, g_int, g_dsmo, g_ssmo = _render_losses_launch(
            rgb, depth, semantic, intensity, batch, cfg)
        dev = losses.device
        out = torch.empty(4, device=dev, dtype=torch.float32)
        # (zero-weight terms name every output, so all four are written whatever the configuration)
        terms = [(losses[:1], None, 0.0, k) for k in range(4)] + [(losses[k:k + 1], None, 1.0, 0) for k in range(6) if used[k]]
        g_dist = None
        ctx.dist_coef = 0.
        if dist_mult > 0:   # stepfun.lossfun_distortion on the final level, mean over the rays
            sdist, weights_c = f32(sdist), f32(weights)
            N, S = weights_c.shape
            dl = torch.empty(N, device=dev, dtype=torch.float32)
            g_dist = torch.empty_like(weights_c)
            with torch.cuda.device(dev):
                with timed('distortion'):
                    check(load().nlb_distortion_loss(ptr(sdist), ptr(weights_c), N, S, ptr(dl), ptr(g_dist), stream()))
            ctx.dist_coef = float(dist_mult) / N
            terms += [(dl, None, ctx.dist_coef, 1), (1, None, 1.0, 0)]
        for v in extra_values:          # reported values without a gradient (latent_reg)
            terms.append((f32(v).reshape(1), None, 1.0, 0))
        terms.append((0, None, 1.0, 2))            # total = main (+ proposal losses + hash decay)
        if prop_value is not None:
            terms.append((f32(prop_value).reshape(1), None, 1.0, 2))
        if hash_decay is not None:
            hd = f32(hash_decay).reshape(1)
            terms += [(hd, None, 1.0, 3), (3, None, 1.0, 2)]
        weighted_sums(terms, out)
        ctx.save_for_backward(scales, g_rgb, g_depth, g_sem, g_int, g_dsmo, g_ssmo, g_dist)
        ctx.used = tuple(1.0 if u else 0.0 for u in used)
        ctx.shapes = (rgb.shape, depth.shape, None if semantic is None else semantic.shape,
                      None if intensity is None else intensity.shape)
        main, extras = out.split([1, 3])
        ctx.mark_non_differentiable(losses, extras)
        return main.view(()), losses, extras

    @staticmethod
    def backward(ctx, g, _gl, _go):
        scales, g_rgb, g_depth, g_sem, g_int, g_dsmo, g_ssmo, g_dist = ctx.saved_tensors
        if g is None:
            return (None,) * 13
        g = f32(g)
        u = ctx.used
        sc = lambda k: scales[k:k + 1]
        o_rgb, o_depth = torch.empty_like(g_rgb), torch.empty_like(g_depth)
        jobs = [(g_rgb, None, o_rgb, g, sc(0), None, u[0], 0.),
                (g_depth, g_dsmo, o_depth, g, sc(1), sc(4), u[1], u[4])]
        o_sem = o_int = o_w = None
        if g_sem is not None:
            o_sem = torch.empty_like(g_sem)
            jobs.append((g_sem, g_ssmo, o_sem, g, sc(2), sc(5), u[2], u[5]))
        if g_int is not None:
            o_int = torch.empty_like(g_int)
            jobs.append((g_int, None, o_int, g, sc(3), None, u[3], 0.))
        if g_dist is not None:
            o_w = torch.empty_like(g_dist)
            jobs.append((g_dist, None, o_w, g, None, None, ctx.dist_coef, 0.))
        scale_tensors(jobs)
        shp = ctx.shapes
        return (o_rgb.view(shp[0]), o_depth.view(shp[1]), None if o_sem is None else o_sem.view(shp[2]),
                None if o_int is None else o_int.view(shp[3]), o_w, None, None, None, None, None, None, None, None)


def main_loss(rendering, weights, sdist, batch, cfg, used, dist_mult: float, prop_value=None, hash_decay=None,
              extra_values=()):
    """See _MainLoss."""
    return _MainLoss.apply(rendering['rgb'], rendering['depth'], rendering.get('semantic'), rendering.get('intensity'),
                           weights, sdist.detach(), batch, cfg, tuple(bool(x) for x in used), float(dist_mult),
                           None if prop_value is None else prop_value.detach(),
                           None if hash_decay is None else hash_decay.detach(), tuple(extra_values))


# ----------------------------------------------------------------------------- dynamic objects (csrc/obj.cu)
@torch.no_grad()
def obj_pose(timestamp: torch.Tensor, tracks: torch.Tensor) -> torch.Tensor:
    """obj_utils.get_pose (Z/internal/obj_utils.py:431-475): timestamp [N,1], tracks [n_obj,T,9] -> [N,n_obj,9]."""
    t, tr = f32(timestamp).reshape(-1), f32(tracks)
    if tr.dim() != 3 or tr.shape[-1] != 9:
        raise RuntimeError(f'tracks must be [n_obj, T, 9], got {tuple(tr.shape)}')
    N, n_obj, T = t.shape[0], tr.shape[0], tr.shape[1]
    pose = torch.empty(N, n_obj, 9, device=t.device, dtype=torch.float32)
    with torch.cuda.device(t.device):
        check(load().nlb_obj_pose(ptr(t), ptr(tr), N, n_obj, T, ptr(pose), stream()))
    return pose


def _obj_mlp_desc(mlp, latent):
    """Pointers / sizes of an ObjMLP for the kernel; keeps the fp32 tensors alive through the returned list."""
    if mlp.net_depth_viewdirs != 2 or mlp.skip_layer_dir != 0 or mlp.disable_rgb or not mlp.fixed_semantic \
            or mlp.re_weights or mlp.warp_fn is not None:
        raise NotImplementedError('ObjMLP: only the nuscenes_single.gin object network is built (two view layers with a '
                                  'skip after the first, fixed semantic class, no contraction, no re-weighting)')
    d0, d2 = mlp.density_layer[0], mlp.density_layer[2]
    v0, v1 = mlp.lin_second_stage_0, mlp.lin_second_stage_1
    keep = [f32(x.detach()) for x in (d0.weight, d0.bias, d2.weight, d2.bias, v0.weight, v0.bias, v1.weight, v1.bias,
                                      mlp.rgb_layer.weight, mlp.rgb_layer.bias)]
    lat = None
    shape_n = tex_n = 0
    if latent is not None:
        lat = f32(latent.detach())
        if mlp.split_latent:
            shape_n = mlp.latent_size // 2
            tex_n = mlp.latent_size - shape_n
        else:
            shape_n = mlp.latent_size
        keep.append(lat)
    F = mlp.encoder.output_dim
    dir_dim = 3 + 6 * mlp.deg_view
    want = {'density_layer.0': (d0.weight, (64, F + shape_n)), 'lin_second_stage_0': (v0.weight, (mlp.net_width_viewdirs, mlp.bottleneck_width + dir_dim + tex_n)),
            'lin_second_stage_1': (v1.weight, (mlp.net_width_viewdirs, mlp.net_width_viewdirs + mlp.bottleneck_width + dir_dim + tex_n)),
            'rgb_layer': (mlp.rgb_layer.weight, (3, mlp.net_width_viewdirs))}
    for name, (w, shp) in want.items():
        if tuple(w.shape) != shp:
            raise RuntimeError(f'ObjMLP.{name}: weight {tuple(w.shape)}, expected {shp}')
    desc = NlbObjMlp(*[ptr(x) for x in keep[:10]], ptr(lat), d0.weight.shape[0], mlp.bottleneck_width,
                     mlp.net_width_viewdirs, mlp.deg_view, shape_n, tex_n, float(mlp.density_bias),
                     float(mlp.rgb_premultiplier), float(mlp.rgb_bias), float(mlp.rgb_padding), int(mlp.class_type),
                     int(mlp.class_num))
    return desc, keep


def _obj_param_list(mlp, latent):
    d0, d2 = mlp.density_layer[0], mlp.density_layer[2]
    v0, v1 = mlp.lin_second_stage_0, mlp.lin_second_stage_1
    return [d0.weight, d0.bias, d2.weight, d2.bias, v0.weight, v0.bias, v1.weight, v1.bias, mlp.rgb_layer.weight,
            mlp.rgb_layer.bias, latent, mlp.encoder.embeddings]


def _obj_forward_all(model, tdist, rays, viewdirs, pose, density, rgb, sem, owner):
    """All tracks of one level, in track order, overwriting density / rgb / semantic in place; returns the byte mask."""
    N, S = rays.N, tdist.shape[1] - 1
    mask = torch.zeros(N, S, device=rays.device, dtype=torch.uint8)
    n_obj = pose.shape[1]
    lib = load()
    with torch.cuda.device(rays.device):
        for track_id in range(n_obj):
            mlp, latent = model._obj_network(track_id)
            desc, keep = _obj_mlp_desc(mlp, latent)
            tab = _table_desc(mlp.encoder, mlp.encoder.embeddings.detach())
            with timed('obj_forward'):
                check(lib.nlb_obj_forward(ptr(tdist), ptr(rays.origins), ptr(rays.directions), ptr(viewdirs), ptr(pose),
                                          n_obj, track_id, N, S, C.byref(tab), C.byref(desc), ptr(density), ptr(rgb),
                                          ptr(sem), ptr(mask), ptr(owner), stream()))
    return mask


class _ObjBranch(Function):
    """Final-level object branch in training: (density_o, rgb_o) hold the ObjMLP outputs on the samples inside a
    box (anything elsewhere); gradients go to the ObjMLP weights, the latent codes and the object tables of the
    owning track (csrc/obj.cu k_obj_backward: forward recomputed, nothing saved but the owner map)."""

    @staticmethod
    def forward(ctx, model, tdist, rays, viewdirs, pose, n_params, *params):
        N, S = rays.N, tdist.shape[1] - 1
        dev = rays.device
        density = torch.zeros(N, S, device=dev)
        rgb = torch.zeros(N, S, 3, device=dev)
        sem = torch.zeros(N, S, model.nerf_mlp.class_num, device=dev)
        owner = torch.full((N, S), -1, device=dev, dtype=torch.int32)
        mask = _obj_forward_all(model, tdist, rays, viewdirs, pose, density, rgb, sem, owner)
        ctx.model, ctx.rays, ctx.param_order = model, rays, params
        ctx.save_for_backward(tdist, viewdirs, pose, owner)
        ctx.mark_non_differentiable(sem, mask)
        return density, rgb, sem, mask

    @staticmethod
    def backward(ctx, g_density, g_rgb, _gs, _gm):
        tdist, viewdirs, pose, owner = ctx.saved_tensors
        model, rays = ctx.model, ctx.rays
        N, S = rays.N, tdist.shape[1] - 1
        n_obj = pose.shape[1]
        g_density = None if g_density is None else f32(g_density)
        g_rgb = None if g_rgb is None else f32(g_rgb)
        grads = {}          # id(param) -> (param, gradient target, returned through autograd?)
        lib = load()
        # track refinement (Z/train.py:244-257): the interpolated poses carry a graph back to Track_opt
        g_pose = torch.zeros_like(pose) if ctx.needs_input_grad[4] else None
        with torch.cuda.device(rays.device):
            for track_id in range(n_obj):
                mlp, latent = model._obj_network(track_id)
                plist = _obj_param_list(mlp, latent)
                tgt = []
                for prm in plist:
                    if prm is None:
                        tgt.append(None)
                        continue
                    if id(prm) not in grads:
                        buf = getattr(prm, '_nlb_grad', None)     # the trainer's persistent buffers: added into directly
                        grads[id(prm)] = (prm, buf if buf is not None else torch.zeros_like(prm), buf is None)
                    tgt.append(grads[id(prm)][1])
                desc, keep = _obj_mlp_desc(mlp, latent)
                tab = _table_desc(mlp.encoder, mlp.encoder.embeddings.detach())
                gd = NlbObjGrads(*[ptr(t) for t in tgt], ptr(g_pose))
                with timed('obj_backward'):
                    check(lib.nlb_obj_backward(ptr(tdist), ptr(rays.origins), ptr(rays.directions), ptr(viewdirs), ptr(pose),
                                               n_obj, track_id, N, S, C.byref(tab), C.byref(desc), ptr(owner), ptr(g_density),
                                               ptr(g_rgb), C.byref(gd), stream()))
        out = []
        for prm in ctx.param_order:
            p_, t_, ret = grads.get(id(prm), (None, None, False))
            out.append(t_ if ret else None)
        return (None, None, None, None, g_pose, None, *out)


def obj_pose_torch(time: torch.Tensor, tracks: torch.Tensor) -> torch.Tensor:
    """obj_utils.get_pose (Z/internal/obj_utils.py:431-475) in torch, for a track table that carries a graph (track
    refinement): the two entries closest in time (stable order on ties, like the kernel), blended by
    clamp(|t - t2| / (|t1 - t2| + 1e-9), 0, 1).  time [N,1], tracks [n_obj, T, 9] -> [N, n_obj, 9]."""
    ts = tracks[:, :, -2]                                                  # [n_obj, T]
    diff = (time.reshape(-1, 1, 1) - ts[None]).abs()                        # [N, n_obj, T]
    idx = torch.sort(diff.detach(), dim=-1, stable=True).indices[..., :2]
    n_obj, T, D = tracks.shape
    flat = tracks.reshape(n_obj * T, D)
    base = torch.arange(n_obj, device=tracks.device)[None, :, None] * T
    info = flat.index_select(0, (idx + base).reshape(-1)).reshape(*idx.shape, D)   # [N, n_obj, 2, 9]
    t1, t2 = info[..., 0, -2], info[..., 1, -2]
    w1 = ((time.reshape(-1, 1) - t2).abs() / ((t1 - t2).abs() + 1e-9)).clamp(0, 1)[..., None]
    return w1 * info[..., 0, :] + (1 - w1) * info[..., 1, :]


def obj_apply_train(model, res, tdist, rays, viewdirs, pose, is_prop: bool) -> torch.Tensor:
    """Training-mode object branch (Z/internal/models.py:401-477).  Proposal levels: the object densities are
    detached (`:449-451`), so the merge is a masked overwrite that only blocks the PropMLP's gradient on the object
    samples.  Final level: `torch.where(mask, object outputs, scene outputs)` with the object outputs carrying
    gradients to the object networks."""
    N, S = rays.N, tdist.shape[1] - 1
    if res.get('intensity') is not None:
        raise NotImplementedError('object branch with an intensity head (see Model._init_objects)')
    if is_prop:
        with torch.no_grad():
            d_obj = res['density'].detach().clone()
            mask = _obj_forward_all(model, tdist, rays, viewdirs, pose, d_obj, None, None, None).bool()
        res['density'] = torch.where(mask, d_obj, res['density'])
        return mask
    params, seen = [], set()
    for track_id in range(pose.shape[1]):
        mlp, latent = model._obj_network(track_id)
        for prm in _obj_param_list(mlp, latent):
            if prm is not None and id(prm) not in seen:
                seen.add(id(prm))
                params.append(prm)
    d_o, rgb_o, sem_o, mask = _ObjBranch.apply(model, tdist, rays, viewdirs, pose, len(params), *params)
    mask = mask.bool()
    res['density'] = torch.where(mask, d_o, res['density'])
    if res.get('rgb') is not None:
        res['rgb'] = torch.where(mask[..., None], rgb_o, res['rgb'])
    if res.get('semantic') is not None:
        res['semantic'] = torch.where(mask[..., None], sem_o, res['semantic'])
    return mask


@torch.no_grad()
def obj_apply(model, res: Dict[str, torch.Tensor], tdist: torch.Tensor, rays: RayBundle, viewdirs: torch.Tensor,
              pose: torch.Tensor, is_prop: bool) -> torch.Tensor:
    """The per-track loop of Z/internal/models.py:415-477 for one sampling level: box test, ObjMLP on the hits,
    masked overwrite of density (and rgb / semantic at the final level), in track order, without leaving the
    device.  Returns obj_mask [N,S] (bool)."""
    N, S = rays.N, tdist.shape[1] - 1
    if res.get('intensity') is not None:
        raise NotImplementedError('object branch with an intensity head (see Model._init_objects)')
    density = res['density']
    rgb = res['rgb'] if (not is_prop and res.get('rgb') is not None) else None
    sem = res['semantic'] if (not is_prop and res.get('semantic') is not None) else None
    return _obj_forward_all(model, tdist, rays, viewdirs, pose, density, rgb, sem, None).bool()
