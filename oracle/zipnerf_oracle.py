"""ORACLE (test infrastructure, not product code) -- CPU restatement in torch
fp32 of the reference's zipnerf volume-rendering hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this module; nerf_lidar_b200/ never does.

Citations are into /root/reference/NeRF_LiDAR/zipnerf/ (`Z/`).  The reference
has no tests or golden vectors for this path; this restatement is
pinned against the reference's OWN Python executed in the build container:
tests/golden/make_golden.py runs Z/internal/models.py `Model.forward` through
oracle/ref_shims.py and stores its outputs; tests/test_oracle_golden.py checks
this file against them (and tests/test_reference_import.py re-runs the live
comparison whenever /root/reference is present).

Random draws are INPUTS here (`rand_inputs`, one dict per level with 'jitter'
[N,1] and 'deg' [N,S,7]) so CPU and GPU see the same jitter.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from oracle import grid_oracle

EPS = float(torch.finfo(torch.float32).eps)


# ---------------------------------------------------------------- ray warps
def power_fwd(x, lam):
    """Z/internal/coord.py:103-108 power_transformation."""
    a = abs(lam - 1)
    return a / lam * ((x / a + 1) ** lam - 1)


def power_inv(y, lam):
    """Z/internal/coord.py:111-118 inv_power_transformation (eps inside the pow)."""
    a = abs(lam - 1)
    return ((y * lam / a + 1 + EPS) ** (1 / lam) - 1) * a


def s_to_t(s, near, far, lam=-1.5):
    """Z/internal/coord.py:143-162 with fn='power_transformation'."""
    s_near, s_far = power_fwd(near * 2, lam), power_fwd(far * 2, lam)
    return power_inv(s * s_far + (1 - s) * s_near, lam) / 2


# ---------------------------------------------------------------- step functions
def max_dilate_weights(t, w, dilation, lo, hi):
    """Z/internal/stepfun.py:64-105 (weight_to_pdf, max_dilate, pdf_to_weight,
    renormalize=True) for t[N,n+1], w[N,n]."""
    p = w / (t[:, 1:] - t[:, :-1]).clamp_min(EPS)
    a = t[:, :-1] - dilation
    b = t[:, 1:] + dilation
    td = torch.sort(torch.cat([t, a, b], -1), -1).values.clamp(lo, hi)
    inside = (a[:, None, :] <= td[:, :, None]) & (b[:, None, :] > td[:, :, None])
    pd = torch.where(inside, p[:, None, :], torch.zeros((), dtype=p.dtype)).max(-1).values[:, :-1]
    wd = pd * (td[:, 1:] - td[:, :-1])
    wd = wd / wd.sum(-1, keepdim=True).clamp_min(EPS)
    return td, wd


def interp_sorted(x, xp, fp):
    """Z/internal/math.py:89-108 sorted_interp for sorted xp AND non-decreasing
    fp, restated with a binary search: lower knot = last j with xp_j <= x
    (default 0), upper knot = first j with xp_j > x (default last).  Returns the
    interpolated values and the lower-knot index (the 'sample index')."""
    n = xp.shape[-1]
    cnt = torch.searchsorted(xp.contiguous(), x.contiguous(), right=True)  # #{j: xp_j <= x}
    i0 = (cnt - 1).clamp(0, n - 1)
    i1 = cnt.clamp(0, n - 1)
    x0, x1 = xp.gather(-1, i0), xp.gather(-1, i1)
    f0, f1 = fp.gather(-1, i0), fp.gather(-1, i1)
    off = torch.nan_to_num((x - x0) / (x1 - x0), 0).clamp(0, 1)
    return f0 + off * (f1 - f0), i0


def cdf_from_logits(logits):
    """softmax + integrate_weights, Z/internal/stepfun.py:108-128,154-158."""
    w = torch.softmax(logits, -1)
    cw = torch.cumsum(w[:, :-1], -1).clamp_max(1)
    z = torch.zeros(w.shape[0], 1, dtype=w.dtype)
    return torch.cat([z, cw, z + 1], -1)


def sample_u(num_samples: int, jitter: Optional[torch.Tensor], n_rays: int):
    """Sample positions in the CDF, Z/internal/stepfun.py:199-216 with
    deterministic_center=True and single_jitter=True."""
    if jitter is None:
        pad = 1 / (2 * num_samples)
        u = torch.linspace(pad, 1. - pad - EPS, num_samples)
        return u.expand(n_rays, num_samples)
    u_max = EPS + (1 - EPS) / num_samples
    max_jitter = (1 - u_max) / (num_samples - 1) - EPS
    return torch.linspace(0, 1 - u_max, num_samples) + jitter * max_jitter


def sample_intervals(t, logits, num_samples, jitter, lo, hi):
    """Z/internal/stepfun.py:251-294.  Returns (fenceposts [N,S+1], knot index
    of every center [N,S])."""
    u = sample_u(num_samples, jitter, t.shape[0])
    centers, idx = interp_sorted(u, cdf_from_logits(logits), t)
    mid = (centers[:, 1:] + centers[:, :-1]) / 2
    first = (2 * centers[:, :1] - mid[:, :1]).clamp_min(lo)
    last = (2 * centers[:, -1:] - mid[:, -1:]).clamp_max(hi)
    return torch.cat([first, mid, last], -1), idx


def resample_level(sdist, weights, i_level, num_samples, prod_num_samples, train_frac, jitter,
                   dilation_bias=0.0025, dilation_multiplier=0.5, anneal_slope=10.0,
                   resample_padding=0.0):
    """The per-level chain of Z/internal/models.py:320-369."""
    dilation = dilation_bias + dilation_multiplier * 1.0 / prod_num_samples
    if i_level > 0:
        sdist, weights = max_dilate_weights(sdist, weights, dilation, 0.0, 1.0)
        sdist, weights = sdist[:, 1:-1], weights[:, 1:-1]
    anneal = (anneal_slope * train_frac) / ((anneal_slope - 1) * train_frac + 1)
    logits = torch.where(sdist[:, 1:] > sdist[:, :-1], anneal * torch.log(weights + resample_padding),
                         torch.full_like(weights, -math.inf))
    new_s, idx = sample_intervals(sdist, logits, num_samples, jitter, 0.0, 1.0)
    return new_s, idx, sdist, logits


# ---------------------------------------------------------------- sample points
def cast_rays(tdist, origins, directions, radii, base_x, base_y, deg_noise, n=7, m=3, std_scale=0.35):
    """Z/internal/render.py:129-168: 7-point hexagonal multisample per interval."""
    t0, t1 = tdist[:, :-1, None], tdist[:, 1:, None]
    j = torch.arange(n)
    t = t0 + (t1 - t0) * (j + 0.5) / n
    deg = (2 * math.pi * m * j / n).expand(t.shape)
    if deg_noise is not None:
        deg = deg + deg_noise * math.pi * 2
    r = radii[:, :, None]
    local = torch.stack([r * t * torch.cos(deg) / 2, r * t * torch.sin(deg) / 2, t], -1)
    basis = torch.stack([base_x, base_y, directions], -1)  # [N, xyz, k]
    means = torch.matmul(local, basis[:, None].transpose(-1, -2)) + origins[:, None, None, :]
    stds = std_scale * r * t
    return means, stds


def contract_mean_std(x, std):
    """Z/internal/coord.py:51-63 (+ the /2 of models.py:970-973 is applied by the caller)."""
    m2 = (x ** 2).sum(-1, keepdim=True).clamp_min(EPS)
    m = torch.sqrt(m2)
    inside = m2 <= 1
    z = torch.where(inside, x, ((2 * torch.sqrt(m2) - 1) / m2) * x)
    det = ((1 / m2) * ((2 / m - 1 / m2) ** 2))[..., 0]
    std = torch.where(inside[..., 0], std, (det ** (1 / 3)) * std)
    return z, std


def encode_features(means, stds, table, offsets, grid_sizes, C, base_resolution=16):
    """predict_density up to the feature vector, Z/internal/models.py:965-977:
    contract, /2, (x+1)/2, hash-grid lookup, erf re-weighting, mean over the 7
    multisamples.  Returns features[N,S,L*C]."""
    N, S, n, _ = means.shape
    z, sd = contract_mean_std(means.reshape(-1, 3), stds.reshape(-1))
    z, sd = z / 2, sd / 2
    x01 = (z + 1) / 2
    L = offsets.shape[0] - 1
    # autograd-capable wrapper around grid_encode_forward: gradient w.r.t. the table, and w.r.t. the points through
    # dy_dx when they carry a graph (gridencoder/grid.py:169: calc_grad_inputs = inputs.requires_grad -- the
    # pose-refinement window)
    flat = grid_oracle._GridEncodeFn.apply(x01, table, offsets, 1.0, base_resolution, bool(x01.requires_grad), 0,
                                           False, 0)
    feat = flat.reshape(N, S, n, L, C)
    sd = sd.reshape(N, S, n)
    w = torch.erf(1 / torch.clamp(torch.sqrt(8 * sd[..., None] ** 2 * grid_sizes ** 2), min=1e-10))
    return (feat * w[..., None]).mean(-3).flatten(-2, -1)


def pos_enc_dirs(v, deg=4):
    """Z/internal/coord.py:199-210 pos_enc(min_deg=0, max_deg=4, append_identity)."""
    scales = 2 ** torch.arange(0, deg)
    sx = (v[..., None, :] * scales[:, None]).reshape(*v.shape[:-1], -1)
    return torch.cat([v, torch.sin(torch.cat([sx, sx + 0.5 * math.pi], -1))], -1)


# ---------------------------------------------------------------- MLPs
def _lin(p, name, x, q=None):
    """nn.Linear; `q` (optional) rounds the OPERANDS (activations and weights), e.g. to
    bfloat16, while the accumulation and the bias add stay fp32 -- the arithmetic of
    the tensor-core kernel."""
    if q is None:
        return F.linear(x, p[name + '.weight'], p[name + '.bias'])
    return F.linear(q(x), q(p[name + '.weight'])) + p[name + '.bias']


def prop_mlp(p, prefix, feat):
    """PropMLP head, Z/internal/models.py:887-889,996-997,1116."""
    h = torch.relu(_lin(p, prefix + 'density_layer.0', feat))
    raw = _lin(p, prefix + 'density_layer.2', h)[..., 0]
    return F.softplus(raw - 1.0)


def nerf_mlp(p, feat, viewdirs, use_intensity=True, cast=None):
    """NerfMLP, Z/internal/models.py:996-997,1116-1251.  `cast` (e.g. bfloat16
    round-trip) lets tests model the bf16 operand rounding of the tensor-core
    kernel; None = the reference's fp32."""
    q = cast
    pre = 'nerf_mlp.'
    x = _lin(p, pre + 'density_layer.2', torch.relu(_lin(p, pre + 'density_layer.0', feat, q)), q)
    density = F.softplus(x[..., 0] - 1.0)
    sem = torch.softmax(_lin(p, pre + 'sem_layer.2', torch.relu(_lin(p, pre + 'sem_layer.0', x, q)), q), -1)
    inten = None
    if use_intensity:
        inten = _lin(p, pre + 'intensity_layer.2', torch.relu(_lin(p, pre + 'intensity_layer.0', x, q)), q)
    de = pos_enc_dirs(viewdirs)
    de = de[:, None, :].expand(x.shape[0], x.shape[1], de.shape[-1])
    h_in = torch.cat([x, de], -1)
    h = torch.relu(_lin(p, pre + 'lin_second_stage_0', h_in, q))
    h = torch.cat([h, h_in], -1)
    h = torch.relu(_lin(p, pre + 'lin_second_stage_1', h, q))
    rgb = torch.sigmoid(_lin(p, pre + 'rgb_layer', h, q)) * (1 + 2 * 0.001) - 0.001
    return dict(density=density, rgb=rgb, semantic=sem, intensity=inten, bottleneck=x)


# ---------------------------------------------------------------- compositing
def alpha_weights(density, tdist, dirs, opaque_background=True):
    """Z/internal/render.py:170-189."""
    delta = (tdist[:, 1:] - tdist[:, :-1]) * torch.norm(dirs[:, None, :], dim=-1)
    dd = density * delta
    if opaque_background:
        dd = torch.cat([dd[:, :-1], torch.full_like(dd[:, -1:], math.inf)], -1)
    alpha = 1 - torch.exp(-dd)
    trans = torch.exp(-torch.cat([torch.zeros_like(dd[:, :1]), torch.cumsum(dd[:, :-1], -1)], -1))
    return alpha * trans, alpha, trans


def weighted_percentile(t, w, ps=(5, 50, 95)):
    """Z/internal/stepfun.py:329-339."""
    cw = torch.cumsum(w[:, :-1], -1).clamp_max(1)
    z = torch.zeros(w.shape[0], 1, dtype=w.dtype)
    cw = torch.cat([z, cw, z + 1], -1)
    q = (torch.tensor(ps, dtype=torch.float32) / 100).expand(t.shape[0], len(ps))
    return interp_sorted(q, cw, t)[0]


def composite(rgbs, weights, tdist, far, bg=1.0, semantic=None, intensity=None, compute_extras=True):
    """Z/internal/render.py:192-284 volumetric_rendering."""
    out = {}
    acc = weights.sum(-1)
    bg_w = (1 - acc[:, None]).clamp_min(0.)
    out['rgb'] = (weights[..., None] * rgbs).sum(-2) + bg_w * bg
    t_mid = 0.5 * (tdist[:, :-1] + tdist[:, 1:])
    out['depth'] = (weights * t_mid).sum(-1) / acc.clamp_min(EPS)
    if semantic is not None:
        out['semantic'] = (weights.detach()[..., None] * semantic).sum(-2)
    if intensity is not None:
        out['intensity'] = (weights.detach() * intensity.reshape(weights.shape)).sum(-1)
    if compute_extras:
        out['acc'] = acc
        e = (weights * torch.log(t_mid)).sum(-1) / acc.clamp_min(EPS)
        dm = torch.nan_to_num(torch.exp(e), math.inf)
        out['distance_mean'] = torch.minimum(torch.maximum(dm, tdist[:, 0]), tdist[:, -1])
        pct = weighted_percentile(torch.cat([tdist, far], -1), torch.cat([weights, bg_w], -1))
        out['distance_percentile_5'] = pct[:, 0]
        out['distance_median'] = pct[:, 1]
        out['distance_percentile_95'] = pct[:, 2]
    return out


# ---------------------------------------------------------------- the path
def model_forward(p: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor],
                  rand_inputs: Optional[Sequence[Dict[str, torch.Tensor]]], train_frac: float,
                  compute_extras: bool = True, samples=(64, 64, 32), use_intensity: bool = True,
                  mlp_cast=None, training: bool = False, hash_decay_mults: float = 0.1):
    """Z/internal/models.py:239-576 `Model.forward` for the static-scene path of
    nuscenes_single.gin (+Config.use_intensity=True, instance_obj=False)."""
    N = batch['origins'].shape[0]
    near, far = batch['near'], batch['far']
    sdist = torch.cat([torch.zeros_like(near), torch.ones_like(far)], -1)
    weights = torch.ones_like(near)
    prod = 1
    renderings, history = [], []
    for lvl, S in enumerate(samples):
        is_prop = lvl < len(samples) - 1
        r_in = None if rand_inputs is None else rand_inputs[lvl]
        sdist, idx, knots, logits = resample_level(sdist, weights, lvl, S, prod, train_frac,
                                                   None if r_in is None else r_in['jitter'])
        prod *= S
        sdist = sdist.detach()
        tdist = s_to_t(sdist, near, far)
        means, stds = cast_rays(tdist, batch['origins'], batch['directions'], batch['radii'],
                                batch['base_x'], batch['base_y'], None if r_in is None else r_in['deg'])
        pre = f'prop_mlp_{lvl}.' if is_prop else 'nerf_mlp.'
        C = p[pre + 'encoder.embeddings'].shape[1]
        feat = encode_features(means, stds, p[pre + 'encoder.embeddings'], p[pre + 'encoder.offsets'],
                               p[pre + 'encoder.grid_sizes'], C)
        if is_prop:
            res = dict(density=prop_mlp(p, pre, feat), rgb=torch.zeros(N, S, 3), semantic=None, intensity=None)
        else:
            res = nerf_mlp(p, feat, batch['viewdirs'], use_intensity, cast=mlp_cast)
        res['features'] = feat
        weights = alpha_weights(res['density'], tdist, batch['directions'])[0]
        rend = composite(res['rgb'], weights, tdist, far, 1.0,
                         None if is_prop else res['semantic'],
                         None if is_prop or not use_intensity else res['intensity'], compute_extras)
        renderings.append(rend)
        res.update(sdist=sdist.clone(), weights=weights.clone(), tdist=tdist.clone(),
                   sample_idx=idx, resample_knots=knots, resample_logits=logits)
        history.append(res)
    if training and hash_decay_mults > 0:
        renderings[-1]['hash_decay'] = hash_decay_mults * hash_decay(p)
    return renderings, history


def hash_decay(p):
    """Z/internal/models.py:203-223: sum over tables of mean over levels of the
    per-level mean of squared entries (segment_coo 'mean' + .mean())."""
    total = 0.
    for pre in sorted(k[:-len('encoder.embeddings')] for k in p if k.endswith('encoder.embeddings')):
        emb, offs = p[pre + 'encoder.embeddings'], p[pre + 'encoder.offsets']
        L = offs.shape[0] - 1
        per = torch.stack([(emb[int(offs[l]):int(offs[l + 1])] ** 2).mean(0) for l in range(L)])
        total = total + per.mean()
    return total
