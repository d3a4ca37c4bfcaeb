"""`gridencoder.GridEncoder` with the reference module's constructor, attributes,
persistent buffers and state-dict keys (Z/gridencoder/grid.py:96-198), running on
libnlb200.so.  `_grid_encode` mirrors Z/gridencoder/grid.py:24-89: the caller
allocates outputs [L,B,C] / dy_dx, the backend writes in place, the result is
permuted back to [B, L*C]."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _gridencoder as _backend

_gridtype_to_id = {'hash': 0, 'tiled': 1}
_interp_to_id = {'linear': 0, 'smoothstep': 1}


class _grid_encode(Function):
    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False, interpolation=0):
        inputs = inputs.contiguous().float()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        emb = embeddings.contiguous()
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=emb.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=emb.dtype) if calc_grad_inputs else None
        _backend.grid_encode_forward(inputs, emb, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype,
                                     align_corners, interpolation)
        ctx.save_for_backward(inputs, emb, offsets, dy_dx)
        ctx.dims = (B, D, C, L, S, H, gridtype, interpolation, align_corners)
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, emb, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, align_corners = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(emb)
        grad_inputs = torch.zeros_like(inputs, dtype=emb.dtype) if dy_dx is not None else None
        _backend.grid_encode_backward(grad, inputs, emb, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx,
                                      grad_inputs, gridtype, align_corners, interpolation)
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
                 interpolation='linear', init_std=1e-4):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners
        self.init_std = init_std

        sizes, offsets, offset = [], [], 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            res = int(np.ceil(base_resolution * per_level_scale ** i))
            res = res if align_corners else res + 1
            rows = int(np.ceil(min(self.max_params, res ** input_dim) / 8) * 8)
            sizes.append(res)
            offsets.append(offset)
            offset += rows
        offsets.append(offset)
        self.register_buffer('offsets', torch.from_numpy(np.array(offsets, dtype=np.int32)))
        # level id of every row; the reference keeps this 8-byte-per-row buffer for its
        # segment_coo hash-decay loss (models.py:203-223), so it is part of the state dict.
        self.register_buffer('idx', torch.repeat_interleave(
            torch.arange(num_levels, dtype=torch.long),
            torch.from_numpy(np.diff(np.array(offsets, dtype=np.int64)))))
        self.register_buffer('grid_sizes', torch.from_numpy(np.array(sizes, dtype=np.int32)))
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-self.init_std, self.init_std)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> "
                f"{int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} "
                f"gridtype={self.gridtype} align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id)
        return outputs.view(prefix_shape + [self.output_dim])

    @torch.no_grad()
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        D, C, L = self.input_dim, self.embeddings.shape[1], self.offsets.shape[0] - 1
        S, H = float(np.log2(self.per_level_scale)), self.base_resolution
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        _backend.grad_total_variation(inputs.contiguous(), self.embeddings, self.embeddings.grad, self.offsets,
                                      weight, B, D, C, L, S, H, self.gridtype_id, self.align_corners)
