"""ORACLE (test infrastructure, not product code) -- CPU restatement of the
reference's multi-resolution hash-grid encoder.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this module.  The shipped path (nerf_lidar_b200/) never does.

The reference implementation is CUDA-only
(NeRF_LiDAR/zipnerf/gridencoder/src/gridencoder.cu); there is no CPU code to
compile, so this file restates the kernel's arithmetic with vectorised torch
CPU ops.  Pinning: the reference ships no test vectors, so this
restatement is pinned by hand-computed known-answer cells in
tests/test_oracle_grid.py and, end to end, by the golden outputs of the reference's own
Python (tests/test_oracle_golden.py) in which this file stands in for the CUDA kernel.
The reference .cu itself is never executed (it needs the torch-extension build and a GPU
with /root/reference present): for the encoder, parity is pinned by restatement + known
answers, not by running the reference kernel.

Arithmetic notes that make the integer part bit-exact with the CUDA kernel:
  * gridencoder.cu:148 `pos = x*scale + 0.5f` is contracted by nvcc into one
    fma.rn.f32; we compute the product and sum in float64 (exact product of two
    float32) and round once to float32.
  * gridencoder.cu:50-63 hash = xor_d(pos[d] * prime[d]) in uint32 wrap-around.
  * gridencoder.cu:66-84 dense index while stride <= hashmap_size, else hash;
    always `% hashmap_size`.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737)
_U32 = 0xFFFFFFFF


def level_geometry(level: int, S: float, H: int, offsets: np.ndarray):
    """scale / resolution / hashmap_size of one level (gridencoder.cu:137-139).

    `scale = exp2f(level * S) * H - 1.0f` evaluated in float32 like the kernel.
    """
    ls = np.float32(level) * np.float32(S)
    scale = np.float32(np.float32(np.exp2(ls)) * np.float32(H) - np.float32(1.0))
    resolution = int(math.ceil(float(scale))) + 1
    hashmap_size = int(offsets[level + 1]) - int(offsets[level])
    return scale, resolution, hashmap_size


def _fma32(x: torch.Tensor, a: float, b: float) -> torch.Tensor:
    """float32 fma(x, a, b): exact product in float64, one final rounding."""
    return (x.double() * float(a) + float(b)).float()


def grid_index(pos_grid: torch.Tensor, hashmap_size: int, resolution: int,
               gridtype: int = 0, align_corners: bool = False) -> torch.Tensor:
    """get_grid_index of gridencoder.cu:66-84 for int64 `pos_grid[..., D]`
    holding uint32 values.  Returns the row index (before `*C + ch`)."""
    D = pos_grid.shape[-1]
    stride = 1
    index = torch.zeros(pos_grid.shape[:-1], dtype=torch.int64)
    d = 0
    while d < D and stride <= hashmap_size:
        index = (index + pos_grid[..., d] * stride) & _U32
        stride = (stride * (resolution if align_corners else resolution + 1)) & _U32
        d += 1
    if gridtype == 0 and stride > hashmap_size:
        h = torch.zeros_like(index)
        for dd in range(D):
            h = h ^ ((pos_grid[..., dd] * PRIMES[dd]) & _U32)
        index = h
    return index % hashmap_size


def corner_setup(x01: torch.Tensor, level: int, S: float, H: int, offsets: np.ndarray,
                 gridtype: int = 0, align_corners: bool = False, interp: int = 0):
    """For points x01[B,D] in [0,1]: corner row indices [B,2^D] (level-local),
    corner weights [B,2^D] float32, validity mask [B], fractional pos [B,D],
    and scale.  Follows gridencoder.cu:110-191."""
    B, D = x01.shape
    scale, resolution, hashmap_size = level_geometry(level, S, H, offsets)
    valid = ~((x01 < 0) | (x01 > 1)).any(-1)
    pos = _fma32(x01, float(scale), 0.0 if align_corners else 0.5)
    pg_f = torch.floor(pos)
    # (uint32) conversion of a non-negative float
    pos_grid = pg_f.to(torch.int64).clamp_min(0) & _U32
    frac = pos - pg_f
    if interp == 1:
        deriv = 6 * frac * (1.0 - frac)
        frac = frac * frac * (3.0 - 2.0 * frac)
    else:
        deriv = torch.ones_like(frac)
    idxs, ws = [], []
    for c in range(1 << D):
        w = torch.ones(B, dtype=torch.float32)
        pgl = pos_grid.clone()
        for d in range(D):
            if (c >> d) & 1:
                w = w * frac[:, d]
                pgl[:, d] = (pgl[:, d] + 1) & _U32
            else:
                w = w * (1 - frac[:, d])
        idxs.append(grid_index(pgl, hashmap_size, resolution, gridtype, align_corners))
        ws.append(w)
    return (torch.stack(idxs, -1), torch.stack(ws, -1), valid, frac, deriv, pos_grid,
            scale, resolution, hashmap_size)


def grid_encode_forward(inputs: torch.Tensor, embeddings: torch.Tensor, offsets: torch.Tensor,
                        S: float, H: int, calc_dy_dx: bool = False, gridtype: int = 0,
                        align_corners: bool = False, interp: int = 0
                        ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """outputs[L,B,C] (+ dy_dx[B,L*D*C]) as kernel_grid, gridencoder.cu:87-245."""
    inputs = inputs.detach().float().contiguous()
    emb = embeddings.detach().float()
    B, D = inputs.shape
    C = emb.shape[1]
    offs = offsets.cpu().numpy().astype(np.int64)
    L = offs.shape[0] - 1
    out = torch.zeros(L, B, C, dtype=torch.float32)
    dy_dx = torch.zeros(B, L, D, C, dtype=torch.float32) if calc_dy_dx else None
    for l in range(L):
        idx, w, valid, frac, deriv, pos_grid, scale, resolution, hs = corner_setup(
            inputs, l, S, H, offs, gridtype, align_corners, interp)
        table = emb[offs[l]:offs[l + 1]]
        acc = torch.zeros(B, C, dtype=torch.float32)
        for c in range(1 << D):  # same accumulation order as the kernel
            acc = acc + w[:, c:c + 1] * table[idx[:, c]]
        out[l] = torch.where(valid[:, None], acc, torch.zeros_like(acc))
        if calc_dy_dx:
            for gd in range(D):
                g = torch.zeros(B, C, dtype=torch.float32)
                others = [d for d in range(D) if d != gd]
                for c in range(1 << (D - 1)):
                    wq = torch.full((B,), float(scale), dtype=torch.float32)
                    pgl = pos_grid.clone()
                    for nd, d in enumerate(others):
                        if (c >> nd) & 1:
                            wq = wq * frac[:, d]
                            pgl[:, d] = (pgl[:, d] + 1) & _U32
                        else:
                            wq = wq * (1 - frac[:, d])
                    left = grid_index(pgl, hs, resolution, gridtype, align_corners)
                    pgl[:, gd] = (pgl[:, gd] + 1) & _U32
                    right = grid_index(pgl, hs, resolution, gridtype, align_corners)
                    g = g + wq[:, None] * (table[right] - table[left]) * deriv[:, gd:gd + 1]
                dy_dx[:, l, gd, :] = torch.where(valid[:, None], g, torch.zeros_like(g))
    return out, (dy_dx.reshape(B, L * D * C) if calc_dy_dx else None)


def grid_encode_backward(grad: torch.Tensor, inputs: torch.Tensor, embeddings: torch.Tensor,
                         offsets: torch.Tensor, S: float, H: int,
                         dy_dx: Optional[torch.Tensor] = None, gridtype: int = 0,
                         align_corners: bool = False, interp: int = 0):
    """grad[L,B,C] -> grad_embeddings[rows,C] (+ grad_inputs[B,D]).
    kernel_grid_backward gridencoder.cu:248-340 (atomicAdd -> index_add_, so the
    summation order differs from the GPU's non-deterministic one) and
    kernel_input_backward :343-369."""
    inputs = inputs.detach().float().contiguous()
    grad = grad.detach().float()
    B, D = inputs.shape
    C = embeddings.shape[1]
    offs = offsets.cpu().numpy().astype(np.int64)
    L = offs.shape[0] - 1
    ge = torch.zeros(embeddings.shape, dtype=torch.float64)
    for l in range(L):
        idx, w, valid, *_ = corner_setup(inputs, l, S, H, offs, gridtype, align_corners, interp)
        g = torch.where(valid[:, None], grad[l], torch.zeros_like(grad[l]))
        for c in range(1 << D):
            ge.index_add_(0, idx[:, c] + int(offs[l]), (w[:, c:c + 1] * g).double())
    gi = None
    if dy_dx is not None:
        dd = dy_dx.float().reshape(B, L, D, C)
        gi = torch.einsum('lbc,bldc->bd', grad, dd)
    return ge.float(), gi


def make_offsets(input_dim, num_levels, per_level_scale, base_resolution,
                 log2_hashmap_size, align_corners=False):
    """Level sizing of GridEncoder.__init__ (gridencoder/grid.py:120-141)."""
    offsets, resolutions, offset = [], [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        res = res if align_corners else res + 1
        n = min(max_params, res ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        resolutions.append(res)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), np.array(resolutions, dtype=np.int32)


class _GridEncodeFn(torch.autograd.Function):
    """CPU stand-in for gridencoder/grid.py:24-89 (_grid_encode)."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, S, H, calc_grad_inputs, gridtype, align_corners, interp):
        out, dy_dx = grid_encode_forward(inputs, embeddings, offsets, S, H, calc_grad_inputs,
                                         gridtype, align_corners, interp)
        L, B, C = out.shape
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.cfg = (S, H, gridtype, align_corners, interp, L, B, C)
        return out.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        S, H, gridtype, align_corners, interp, L, B, C = ctx.cfg
        g = grad.reshape(B, L, C).permute(1, 0, 2).contiguous()
        ge, gi = grid_encode_backward(g, inputs, embeddings, offsets, S, H, dy_dx,
                                      gridtype, align_corners, interp)
        return gi, ge, None, None, None, None, None, None, None


class GridEncoder(torch.nn.Module):
    """CPU GridEncoder with the reference module's ctor / buffers / state-dict
    (gridencoder/grid.py:96-174), used only to let the reference's own
    internal/models.py run on CPU when generating golden vectors."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2,
                 base_resolution=16, log2_hashmap_size=19, desired_resolution=None,
                 gridtype='hash', align_corners=False, interpolation='linear', init_std=1e-4):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution = per_level_scale, base_resolution
        self.log2_hashmap_size = log2_hashmap_size
        self.output_dim = num_levels * level_dim
        self.gridtype_id = {'hash': 0, 'tiled': 1}[gridtype]
        self.interp_id = {'linear': 0, 'smoothstep': 1}[interpolation]
        self.align_corners, self.init_std = align_corners, init_std
        offs, res = make_offsets(input_dim, num_levels, per_level_scale, base_resolution,
                                 log2_hashmap_size, align_corners)
        self.register_buffer('offsets', torch.from_numpy(offs))
        idx = torch.empty(int(offs[-1]), dtype=torch.long)
        for i in range(num_levels):
            idx[offs[i]:offs[i + 1]] = i
        self.register_buffer('idx', idx)
        self.register_buffer('grid_sizes', torch.from_numpy(res))
        self.embeddings = torch.nn.Parameter(torch.empty(int(offs[-1]), level_dim).uniform_(-init_std, init_std))

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        prefix = list(inputs.shape[:-1])
        flat = inputs.reshape(-1, self.input_dim)
        out = _GridEncodeFn.apply(flat, self.embeddings, self.offsets, float(np.log2(self.per_level_scale)),
                                  self.base_resolution, flat.requires_grad, self.gridtype_id,
                                  self.align_corners, self.interp_id)
        return out.view(prefix + [self.output_dim])
