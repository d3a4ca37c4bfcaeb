"""Drop-in for the reference's compiled extension module `_gridencoder`
(Z/gridencoder/src/bindings.cpp:5-9; Z/gridencoder/grid.py:9-10 imports this name
first).  Same three functions, same argument order; tensors are unwrapped to
device pointers and handed to the C ABI in libnlb200.so (include/nlb200.h).

Put this directory on PYTHONPATH and the reference's own gridencoder/grid.py runs
on the B200 kernels with zero edits."""
from __future__ import annotations

import os
import sys

import torch

try:  # imported as nerf_lidar_b200._gridencoder
    from . import _lib
except ImportError:  # imported as the top-level module `_gridencoder`
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from nerf_lidar_b200 import _lib


def _chk(name, t, dtype=None):
    """CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_* of gridencoder.cu:15-18."""
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor')
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be a contiguous tensor')
    if dtype == 'int':
        if t.dtype != torch.int32:
            raise RuntimeError(f'{name} must be an int tensor')
    elif dtype == 'float':
        if t.dtype not in (torch.float32, torch.float16, torch.float64):
            raise RuntimeError(f'{name} must be a floating tensor')
        if t.dtype != torch.float32:
            raise NotImplementedError(f'{name}: libnlb200 is built for float32 tables only (got {t.dtype})')


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners,
                        interp):
    _chk('inputs', inputs, 'float'); _chk('embeddings', embeddings, 'float')
    _chk('offsets', offsets, 'int'); _chk('outputs', outputs, 'float')
    if dy_dx is not None:
        _chk('dy_dx', dy_dx, 'float')
    with torch.cuda.device(inputs.device):
        _lib.check(_lib.load().nlb_grid_encode_forward(
            _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets), _lib.ptr(outputs), B, D, C, L, float(S), H,
            _lib.ptr(dy_dx), gridtype, int(bool(align_corners)), interp, _lib.stream()))


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs,
                         gridtype, align_corners, interp):
    _chk('grad', grad, 'float'); _chk('inputs', inputs, 'float'); _chk('embeddings', embeddings, 'float')
    _chk('offsets', offsets, 'int'); _chk('grad_embeddings', grad_embeddings, 'float')
    with torch.cuda.device(inputs.device):
        _lib.check(_lib.load().nlb_grid_encode_backward(
            _lib.ptr(grad), _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets), _lib.ptr(grad_embeddings),
            B, D, C, L, float(S), H, _lib.ptr(dy_dx), _lib.ptr(grad_inputs), gridtype, int(bool(align_corners)),
            interp, _lib.stream()))


def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
    _chk('inputs', inputs, 'float'); _chk('embeddings', embeddings, 'float'); _chk('grad', grad, 'float')
    _chk('offsets', offsets, 'int')
    with torch.cuda.device(inputs.device):
        _lib.check(_lib.load().nlb_grad_total_variation(
            _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(grad), _lib.ptr(offsets), float(weight), B, D, C, L,
            float(S), H, gridtype, int(bool(align_corners)), _lib.stream()))
