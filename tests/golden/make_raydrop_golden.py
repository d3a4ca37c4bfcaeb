"""Generates tests/golden/raydrop_ref.npz: outputs of the reference's OWN stage-3 helpers -- LaserScan.set_points /
do_range_projection (NeRF_Lidar_code/src/lidar_utils.py:57-275) and depth_filter (src/depth_filter.py:4-31) --
on a seeded synthetic 32-beam sweep, plus the drop selection of src/drop_simulation_rays.py:104-140
(save_near=False), which is inline in a function that needs a trained U-Net: its dozen numpy lines are executed
here on the reference's LaserScan with seeded logits.
  python tests/golden/make_raydrop_golden.py [--check]"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from nerf_lidar_b200 import synthetic as sy  # noqa: E402

REF = '/root/reference/NeRF_LiDAR/NeRF_Lidar_code/src'
OUT = os.path.join(HERE, 'raydrop_ref.npz')
H, W, SEED = 32, 1024, 9


def inputs():
    """A rendered sweep as render_lidar.py writes it (points_XXXX.npy, points_semantic_XXXX.npy): beam-major
    32 x W points in the LiDAR frame, piecewise-smooth ranges with depth edges, labels in patches."""
    rng = np.random.default_rng(SEED)
    d = sy.lidar_directions(W).astype(np.float32)
    base = 12 + 8 * np.sin(np.linspace(0, 6 * np.pi, W))[None, :] + 2 * np.arange(H)[:, None] / H
    steps = (rng.random((H, W)) < 0.02).cumsum(1) % 3 * 9.0               # depth discontinuities
    rngs = (base + steps + rng.normal(0, 0.05, (H, W))).astype(np.float32).reshape(-1, 1)
    pts = (d * rngs).astype(np.float32)
    pts[5] = pts[6]                                                        # a duplicated return
    sem = np.repeat(rng.integers(0, 16, (H, W // 16)), 16, axis=1).astype(np.float32).reshape(-1)
    rgb = rng.random((H * W, 3)).astype(np.float32)
    logits = rng.normal(0, 2, (2, H, W)).astype(np.float32)
    return pts, sem, rgb, logits


def softmax(x, axis):
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def run_reference():
    sys.path.insert(0, REF)
    lu = importlib.import_module('lidar_utils')
    df = importlib.import_module('depth_filter')
    pts, sem, rgb, logits = inputs()
    out = {}
    scan = lu.LaserScan(H=H, W=W, fov_up=10.67, fov_down=-30.67)
    scan.set_points(pts, remissions=None, semantic=sem, rgb=rgb)
    scan.do_range_projection()
    for k in ('proj_range', 'proj_xyz', 'proj_semantic', 'proj_rgb', 'proj_idx', 'proj_mask', 'proj_x', 'proj_y', 'unproj_range'):
        out[k] = np.asarray(getattr(scan, k))
    out['filter_sem'] = df.depth_filter(pts, sem, return_mask=True, width=1, threshold=1)      # args.semantic_align
    out['filter_plain'] = df.depth_filter(pts, return_mask=True, width=5)
    # drop_simulation_rays.py:104-140 with save_near=False, mask_thre=0.5, place_car=False
    pred_mask = softmax(logits, axis=0)[1]
    out['prob'] = pred_mask.astype(np.float32)
    pred_mask = pred_mask > 0.5
    mask = (pred_mask == 1) & (scan.proj_mask == 1)
    mask = (mask[scan.proj_y, scan.proj_x] == 1) & (out['filter_sem'] == 1)
    remain_points, remain_labels = pts[mask], sem[mask]
    sky = remain_labels == 10
    remain_points, remain_labels = remain_points[~sky], remain_labels[~sky]
    road = (remain_labels == 0) & (remain_points[:, 2] < -3)
    out['remain_points'], out['remain_labels'] = remain_points[~road], remain_labels[~road]
    return {k: np.ascontiguousarray(v) for k, v in out.items()}


if __name__ == '__main__':
    got = run_reference()
    if '--check' in sys.argv:
        gold = np.load(OUT)
        for k in gold.files:
            assert np.array_equal(gold[k], got[k]), k
        print('ok')
    else:
        np.savez_compressed(OUT, **got)
        print('wrote', OUT, os.path.getsize(OUT) // 1024, 'KiB', {k: (v.shape, str(v.dtype)) for k, v in got.items()})
        print('occupied pixels', int(got['proj_mask'].sum()), 'kept by filter', int(got['filter_sem'].sum()), int(got['filter_plain'].sum()),
              'remaining', got['remain_points'].shape[0])
