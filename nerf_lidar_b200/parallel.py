"""Data parallelism over rays: one process per GPU, parameters
replicated, rays sharded.  Training = the gradient sum DDP does for the reference at
Z/train.py:459, split into its two halves around a sharded optimizer pass: reduce-scatter of
the gradients, fused hash-decay + Adam on this rank's 1/world slice, all-gather of the updated
parameters (same bytes on the wire as the all-reduce, 1/world of the optimizer work and of
the Adam moments per GPU); rendering = contiguous ray
ranges per rank and one final gather of the packed per-ray outputs (instead of
the reference's per-chunk, per-leaf all-gathers, Z/internal/models.py:1426-1476).
Works with the gloo backend on CPU tensors for the host-logic tests."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n items for `rank`; sizes differ by at most 1."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, torch.Tensor], world: int, rank: int) -> Dict[str, torch.Tensor]:
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(n, world, rank)
    return {k: v[lo:hi] for k, v in batch.items()}


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def allreduce_grads(buffers: List[torch.Tensor], average: bool = False):
    """Sum (optionally mean) the flat gradient buffers over all ranks.  Issued
    largest-first so the NeRF table (final first in backward order) overlaps with
    the remaining backward work when called from a side stream."""
    if not is_dist() or dist.get_world_size() == 1:
        return
    handles = [dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=True) for b in buffers]
    for h in handles:
        h.wait()
    if average:
        w = float(dist.get_world_size())
        for b in buffers:
            b.div_(w)


def allreduce_grads_async(buffers: List[torch.Tensor]):
    """Starts the sum of the buffers over all ranks and returns the work handles (empty
    when not distributed); the collectives run on the backend's own stream, so kernels
    launched afterwards on the current stream overlap with them."""
    if not is_dist() or dist.get_world_size() == 1:
        return []
    return [dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=True) for b in buffers]


def shard_chunk(n: int, world: int, align: int = 4) -> int:
    """Elements per rank when a buffer of n elements is cut into `world` equal, `align`-element-aligned
    chunks (the buffer is padded to chunk * world)."""
    per = (n + world - 1) // world
    return (per + align - 1) // align * align


def reduce_scatter_async(arena: torch.Tensor, chunk: int, rank: int):
    """Sum over the ranks of the padded 1-D buffer `arena` (chunk * world elements), leaving this rank's chunk
    arena[rank * chunk : (rank + 1) * chunk] reduced IN PLACE (the other chunks hold partial garbage afterwards
    and are cleared by the caller).  Returns the work handles.  gloo has no reduce-scatter: the host-logic
    tests run the equivalent all-reduce."""
    if not is_dist() or dist.get_world_size() == 1:
        return []
    mine = arena[rank * chunk:(rank + 1) * chunk]
    if dist.get_backend() == 'gloo':
        return [dist.all_reduce(arena, op=dist.ReduceOp.SUM, async_op=True)]
    return [dist.reduce_scatter_tensor(mine, arena, op=dist.ReduceOp.SUM, async_op=True)]


def all_gather_async(arena: torch.Tensor, chunk: int, rank: int):
    """Every rank's chunk of the padded 1-D buffer -> all ranks, in place."""
    if not is_dist() or dist.get_world_size() == 1:
        return []
    mine = arena[rank * chunk:(rank + 1) * chunk]
    if dist.get_backend() == 'gloo':
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        h = dist.all_gather(parts, mine.clone(), async_op=True)
        h.wait()
        arena.copy_(torch.cat(parts))
        return []
    return [dist.all_gather_into_tensor(arena, mine, async_op=True)]


def wait_all(handles):
    for h in handles:
        h.wait()


def gather_rendering(local: Dict, num_rays: int, world: int, rank: int) -> Dict:
    """All-gathers per-ray leaves [n_local, ...] into [num_rays, ...] with ONE
    collective: leaves are flattened to [n_local, width], concatenated along the
    width, padded to the largest shard and exchanged as a single tensor."""
    keys = [k for k, v in local.items() if torch.is_tensor(v) and not k.startswith('ray_') and 'hash' not in k]
    list_keys = [k for k, v in local.items() if isinstance(v, list)]
    sizes = [shard_range(num_rays, world, r) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    if keys:
        shapes = {k: tuple(local[k].shape[1:]) for k in keys}
        flat = torch.cat([local[k].reshape(local[k].shape[0], -1).float() for k in keys], dim=1)
    else:
        shapes, flat = {}, None
    out = dict(local)
    if flat is not None:
        pad = torch.zeros(n_max, flat.shape[1], device=flat.device, dtype=flat.dtype)
        pad[:flat.shape[0]] = flat
        recv = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(recv, pad)
        full = torch.cat([recv[r][:sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
        col = 0
        for k in keys:
            width = 1
            for s in shapes[k]:
                width *= s
            out[k] = full[:, col:col + width].reshape((num_rays,) + shapes[k])
            col += width
    for k in list_keys:  # small visualisation bundles: keep rank-local
        out[k] = local[k]
    return out
