"""Host logic of models.render_image (Z/internal/models.py:1379-1507) on CPU with a stub model: chunking by
config.render_chunk_size, per-leaf result buffers, the ray_* visualisation bundles, [H, W] layout, the restored
training flag, and -- over gloo, world_size 2 -- contiguous ray shards with the single packed gather.  (The
CUDA-graph replay of deterministic chunks needs a device and is covered by the -m gpu suite.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_lidar_b200 import configs, models


class _StubModel:
    """Deterministic per-ray outputs derived from the origins, three levels like Model.forward."""
    anneal_slope = 10.

    def __init__(self, cfg):
        self.config, self.training, self.calls = cfg, True, []

    def eval(self):
        self.training = False

    def train(self):
        self.training = True

    def __call__(self, rand, chunk, train_frac=1., compute_extras=True, zero_glo=True):
        assert not self.training and compute_extras and zero_glo
        o = chunk['origins']
        self.calls.append(o.shape[0])
        n = self.config.vis_num_rays
        rend = [dict(rgb=o * (lvl + 1), depth=o.sum(-1) * train_frac, semantic=o.repeat(1, 2), acc=o[:, 0],
                     ray_sdist=o[:n, :1].repeat(1, 5), ray_weights=o[:n, :1].repeat(1, 4),
                     ray_rgbs=o[:n, None, :].repeat(1, 4, 1)) for lvl in range(3)]
        hist = [dict(weights=o[:, :1].repeat(1, 4) + lvl, sdist=o[:, :1].repeat(1, 5)) for lvl in range(3)]
        return rend, hist


def _batch(n):
    return dict(origins=torch.arange(n * 3).float().reshape(n, 3), directions=torch.ones(n, 3),
                near=torch.zeros(n, 1), absent=None)


def _cfg(chunk):
    cfg = configs.nuscenes_single()
    cfg.render_chunk_size = chunk
    return cfg


@pytest.mark.parametrize('n,chunk', [(257, 100), (100, 100), (5, 16384)])
def test_chunks_assemble_in_order(n, chunk):
    cfg = _cfg(chunk)
    model = _StubModel(cfg)
    b = _batch(n)
    out = models.render_image(model, None, b, False, cfg, train_frac=0.5, image=False, verbose=False,
                              return_weights=True)
    assert model.calls == [min(chunk, n - i) for i in range(0, n, chunk)]
    assert model.training                                  # restored
    assert torch.equal(out['rgb'], b['origins'] * 3)       # the final level's rendering
    assert torch.equal(out['depth'], b['origins'].sum(-1, keepdim=True) * 0.5)
    assert torch.equal(out['semantic'], b['origins'].repeat(1, 2))
    assert torch.equal(out['weights'], b['origins'][:, :1].repeat(1, 4) + 2)   # ray_history[-1]
    keep = min(cfg.vis_num_rays, sum(min(cfg.vis_num_rays, c) for c in model.calls))
    for k, width in (('ray_sdist', (5,)), ('ray_weights', (4,)), ('ray_rgbs', (4, 3))):
        assert len(out[k]) == 3 and all(z.shape == (keep,) + width for z in out[k])


def test_image_layout():
    cfg = _cfg(64)
    h, w = 9, 13
    b = {k: (v.reshape(h, w, -1) if v is not None else None) for k, v in _batch(h * w).items()}
    out = models.render_image(_StubModel(cfg), None, b, False, cfg, image=True, verbose=False)
    assert out['rgb'].shape == (h, w, 3) and out['acc'].shape == (h, w)
    assert torch.equal(out['rgb'], b['origins'] * 3)
    assert 'weights' not in out


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Acc:
    def __init__(self, rank, world):
        self.process_index, self.num_processes, self.is_main_process = rank, world, rank == 0


def _worker(rank, world, port, n, chunk):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        cfg = _cfg(chunk)
        model = _StubModel(cfg)
        b = _batch(n)
        out = models.render_image(model, _Acc(rank, world), b, False, cfg, image=False, verbose=False)
        lo, hi = n * rank // world, n * (rank + 1) // world
        assert sum(model.calls) in (hi - lo, hi - lo + 1, hi - lo - 1)     # this rank rendered only its shard
        assert sum(model.calls) < n
        assert torch.equal(out['rgb'], b['origins'] * 3)                   # every rank holds the full image
        assert torch.equal(out['semantic'], b['origins'].repeat(1, 2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n,chunk', [(257, 50), (34688, 16384)])
def test_sharded_render_world2(n, chunk):
    mp.spawn(_worker, args=(2, _free_port(), n, chunk), nprocs=2, join=True)


REF = '/root/reference/NeRF_LiDAR/zipnerf'


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_same_result_as_the_reference_render_image():
    """The reference's OWN render_image (Z/internal/models.py:1379-1507, imported through oracle/ref_shims.py)
    and this build's, driven with the same stub model, rays and RNG seed: identical leaves, ray_* bundles
    included, for the flat LiDAR layout and the [H, W] image layout."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f'''
import sys, contextlib, warnings
warnings.filterwarnings('ignore')
sys.path.insert(0, {root!r})
import torch
from oracle import ref_shims
ref_models = ref_shims.import_reference()
from nerf_lidar_b200 import models
from tests.test_render_image_cpu import _StubModel, _batch, _cfg


class Acc:
    process_index, num_processes, is_main_process = 0, 1, True
    autocast = staticmethod(contextlib.nullcontext)
    gather = staticmethod(lambda v: v)


for n, chunk, image in ((257, 100, False), (9 * 13, 64, True)):
    cfg = _cfg(chunk)
    b = _batch(n)
    if image:
        b = {{k: (v.reshape(9, 13, -1) if v is not None else None) for k, v in b.items()}}
    torch.manual_seed(3)
    want = ref_models.render_image(_StubModel(cfg), Acc(), dict(b), False, cfg, train_frac=0.5, verbose=False, image=image)
    torch.manual_seed(3)
    got = models.render_image(_StubModel(cfg), None, dict(b), False, cfg, train_frac=0.5, verbose=False, image=image)
    assert sorted(got) == sorted(want), (sorted(got), sorted(want))
    for k, v in want.items():
        if isinstance(v, list):
            assert len(got[k]) == len(v) and all(torch.equal(x, y) for x, y in zip(got[k], v)), k
        else:
            assert got[k].shape == v.shape and torch.equal(got[k], v), (k, got[k].shape, v.shape)
print('ok')
'''
    env = dict(os.environ, TORCHDYNAMO_DISABLE='1')
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stdout[-500:] + r.stderr[-3000:]
