"""TEST INFRASTRUCTURE -- the reference's own CUDA grid encoder, loaded from the prebuilt
`oracle/_ref/_gridencoder_ref.so` (built by oracle/build_ref.py from the reference sources where they
lie; nothing of the reference is copied into this repository).

`backend()` returns the compiled module: its three functions are the reference's
`grid_encode_forward / grid_encode_backward / grad_total_variation`
(Z/gridencoder/src/bindings.cpp:5-9, Z/gridencoder/src/gridencoder.cu:371-470).  The helpers below drive
it exactly the way the reference's Python does (Z/gridencoder/grid.py:24-89: outputs `[L,B,C]` + the
permute copy; Z/internal/models.py:974-977: erf re-weighting and the mean over the 7 multisamples), so the
fused kernels can be checked and timed against the chain they replace.

Only tests/, __graft_entry__.smoke() and bench.py's reference-kernel leg may import this."""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, '_ref', '_gridencoder_ref.so')
_mod = None


def available() -> bool:
    return os.path.exists(SO)


def backend():
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError(f'{SO} is missing: run `python oracle/build_ref.py` where /root/reference exists')
        spec = importlib.util.spec_from_file_location('_gridencoder_ref', SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def encode_forward(x01: torch.Tensor, embeddings: torch.Tensor, offsets: torch.Tensor, per_level_scale: float,
                   base_resolution: int, calc_dy_dx: bool = False, permute: bool = True):
    """_grid_encode.forward (Z/gridencoder/grid.py:27-63) on the reference kernel.  x01 in [0,1], [B,3].
    Returns (outputs [B, L*C] (permute copy) or [L,B,C], dy_dx or None)."""
    B, D = x01.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    S = float(np.log2(per_level_scale))
    out = torch.empty(L, B, C, device=x01.device, dtype=embeddings.dtype)
    dy_dx = torch.empty(B, L * D * C, device=x01.device, dtype=embeddings.dtype) if calc_dy_dx else None
    backend().grid_encode_forward(x01, embeddings, offsets, out, B, D, C, L, S, base_resolution, dy_dx, 0, False, 0)
    if permute:
        out = out.permute(1, 0, 2).reshape(B, L * C)
    return out, dy_dx


def encode_backward(grad_blc: torch.Tensor, x01, embeddings, offsets, per_level_scale, base_resolution, dy_dx=None):
    """_grid_encode.backward (Z/gridencoder/grid.py:65-89): grad [B, L*C] -> (grad_embeddings, grad_inputs)."""
    B, D = x01.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    S = float(np.log2(per_level_scale))
    grad = grad_blc.view(B, L, C).permute(1, 0, 2).contiguous()
    g_emb = torch.zeros_like(embeddings)
    g_in = torch.zeros_like(x01) if dy_dx is not None else None
    backend().grid_encode_backward(grad, x01, embeddings, offsets, g_emb, B, D, C, L, S, base_resolution, dy_dx,
                                   g_in, 0, False, 0)
    return g_emb, g_in


def erf_mean(features: torch.Tensor, stds: torch.Tensor, grid_sizes: torch.Tensor, L: int):
    """Z/internal/models.py:974-977: features [M,7,L*C], stds [M,7] -> [M, L*C]."""
    f = features.unflatten(-1, (L, -1))
    w = torch.erf(1 / torch.clamp(torch.sqrt(8 * stds[..., None] ** 2 * grid_sizes ** 2), min=1e-10))
    return (f * w[..., None]).mean(dim=-3).flatten(-2, -1)
