"""Generates tests/golden/rays_ref.npz: outputs of the reference's OWN ray-generation functions
(Z/internal/camera_utils.py:454-564 pixels_to_rays, Z/internal/lidar_utils.py:8-33 cast_lidar_ray_batch,
:559-568 get_directions) on seeded pixels / poses, so that nerf_lidar_b200/synthetic.py -- which restates them
to build the bench and test batches -- is pinned against the reference on the GPU box as well, where
/root/reference does not exist.
  TORCHDYNAMO_DISABLE=1 python tests/golden/make_ray_golden.py            # writes the fixture
  TORCHDYNAMO_DISABLE=1 python tests/golden/make_ray_golden.py --check    # re-runs the reference, compares"""
import importlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
from nerf_lidar_b200 import synthetic as sy  # noqa: E402
from oracle import ref_shims  # noqa: E402

SEED, N_PIX, LIDAR_W = 5, 512, 1084
OUT = os.path.join(HERE, 'rays_ref.npz')


def inputs():
    rng = np.random.default_rng(SEED)
    o, R = sy._pose(rng, N_PIX)
    px = rng.integers(0, sy.IMG_W, N_PIX)
    py = rng.integers(0, sy.IMG_H, N_PIX)
    return o, R, px, py


def run_reference():
    ref_shims.import_reference()
    cu = importlib.import_module('internal.camera_utils')
    lu = importlib.import_module('internal.lidar_utils')
    o, R, px, py = inputs()
    K = np.array([[sy.FOCAL, 0, sy.IMG_W / 2], [0, sy.FOCAL, sy.IMG_H / 2], [0, 0, 1.]])
    c2w = np.concatenate([R, o[:, :, None]], -1)
    ro, rd, rv, rr, _, bx, by = cu.pixels_to_rays(px, py, np.linalg.inv(K)[None], c2w)
    out = dict(cam_origins=ro, cam_directions=rd, cam_viewdirs=rv, cam_radii=rr, cam_base_x=bx, cam_base_y=by)
    d = lu.get_directions(sy.LIDAR_ELEVATIONS_DEG, np.linspace(270, -90, LIDAR_W) / 180 * np.pi)
    rb = lu.cast_lidar_ray_batch(np.zeros_like(d), d, {})
    out.update({'lidar_' + k: np.asarray(rb[k]) for k in ('directions', 'viewdirs', 'radii')})
    # base_x = base_y = directions in the reference (lidar_utils.py:17-18): stored as two flags, not two copies
    out['lidar_base_is_directions'] = np.array([np.array_equal(rb['base_x'], d), np.array_equal(rb['base_y'], d)])
    return {k: np.ascontiguousarray(v) for k, v in out.items()}


if __name__ == '__main__':
    got = run_reference()
    if '--check' in sys.argv:
        gold = np.load(OUT)
        assert sorted(gold.files) == sorted(got), (gold.files, sorted(got))
        for k in got:
            assert got[k].dtype == gold[k].dtype and np.array_equal(got[k], gold[k]), k
        print('ok')
    else:
        np.savez_compressed(OUT, **got)
        print('wrote', OUT, {k: (v.shape, str(v.dtype)) for k, v in got.items()})
