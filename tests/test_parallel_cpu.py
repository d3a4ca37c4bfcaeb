"""Host-side logic of the data-parallel path on CPU (gloo, world_size 2):
ray sharding, the single packed gather of rendering leaves, gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_lidar_b200 import parallel


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 34688, 5760000):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = dict(rgb=torch.rand(n, 3, generator=g), depth=torch.rand(n, generator=g),
                    semantic=torch.rand(n, 19, generator=g))
        lo, hi = parallel.shard_range(n, world, rank)
        local = {k: v[lo:hi].clone() for k, v in full.items()}
        local['ray_sdist'] = [torch.zeros(4, 65)]
        out = parallel.gather_rendering(local, n, world, rank)
        for k in full:
            assert torch.equal(out[k], full[k]), k
        assert isinstance(out['ray_sdist'], list)
        # gradient all-reduce: sum then mean
        a = torch.full((10,), float(rank + 1))
        b = torch.full((3, 4), float(10 * (rank + 1)))
        parallel.allreduce_grads([a, b])
        assert torch.all(a == 3.0) and torch.all(b == 30.0)
        parallel.allreduce_grads([a], average=True)
        assert torch.all(a == 3.0)
        # split reduction used by the training step: early (NeRF table) and late (the rest) handles
        c, d = torch.full((5,), float(rank)), torch.full((2,), 1.0)
        early = parallel.allreduce_grads_async([c])
        late = parallel.allreduce_grads_async([d])
        parallel.wait_all(early + late)
        assert torch.all(c == 1.0) and torch.all(d == 2.0)
        # sharded optimizer: reduce-scatter of a padded gradient buffer, "update" of this rank's chunk, all-gather
        nel = 1003
        chunk = parallel.shard_chunk(nel, world)
        assert chunk % 4 == 0 and chunk * world >= nel
        grad = torch.zeros(chunk * world)
        grad[:nel] = torch.arange(nel).float() * (rank + 1)
        parallel.wait_all(parallel.reduce_scatter_async(grad, chunk, rank))
        mine = slice(rank * chunk, (rank + 1) * chunk)
        want = torch.zeros(chunk * world)
        want[:nel] = torch.arange(nel).float() * 3
        assert torch.equal(grad[mine], want[mine])
        param = torch.zeros(chunk * world)
        param[mine] = -grad[mine]
        parallel.wait_all(parallel.all_gather_async(param, chunk, rank))
        assert torch.equal(param, -want)
        sh = parallel.shard_batch({'origins': torch.arange(n * 3).reshape(n, 3).float()}, world, rank)
        assert sh['origins'].shape[0] == hi - lo
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [101, 34688])
def test_gather_and_allreduce_world2(n):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)
