"""Stage-3 ray-drop helpers (csrc/raydrop.cu, nerf_lidar_b200/raydrop.py) against the reference's OWN LaserScan /
depth_filter and the drop selection of drop_simulation_rays.py:104-140 -- tests/golden/raydrop_ref.npz, generated
by tests/golden/make_raydrop_golden.py from the imported reference on a seeded 32 x 1024 sweep.
Bars: integer / index work exact; a point whose projection lands within float32 rounding of a pixel border (atan2f /
asinf differ from numpy's libm in the last ulp) may move to the neighbouring pixel -- those are counted and bounded."""
import os
import sys

import numpy as np
import pytest
import torch

from nerf_lidar_b200 import raydrop

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))


def _inputs():
    import make_raydrop_golden as mg
    pts, sem, rgb, logits = mg.inputs()
    t = lambda a: torch.from_numpy(a).cuda()
    return mg, np.load(mg.OUT), t(pts), t(sem), t(rgb), t(logits)


def _scan(mg, pts, sem, rgb):
    scan = raydrop.LaserScan(H=mg.H, W=mg.W, fov_up=10.67, fov_down=-30.67)
    scan.set_points(pts, semantic=sem, rgb=rgb)
    scan.do_range_projection()
    return scan


def test_depth_filter_vs_reference():
    mg, gold, pts, sem, _, _ = _inputs()
    m = raydrop.depth_filter(pts, sem, return_mask=True, width=1, threshold=1).cpu().numpy()
    assert (m != gold['filter_sem']).sum() <= 2        # a neighbour at distance == radius up to rounding
    m = raydrop.depth_filter(pts, return_mask=True, width=5).cpu().numpy()
    assert (m != gold['filter_plain']).sum() <= 2
    kept = raydrop.depth_filter(pts, width=5)
    assert kept.shape == (int(m.sum()), 3)
    with pytest.raises(RuntimeError):
        raydrop.depth_filter(pts[:100])                 # not a 32-beam sweep


def test_range_projection_vs_reference():
    mg, gold, pts, sem, rgb, _ = _inputs()
    scan = _scan(mg, pts, sem, rgb)
    px, py = scan.proj_x.cpu().numpy(), scan.proj_y.cpu().numpy()
    moved = (px != gold['proj_x']) | (py != gold['proj_y'])
    assert moved.sum() <= 8, int(moved.sum())          # border cases only, and by one pixel
    assert np.abs(px - gold['proj_x']).max() <= 1 and np.abs(py - gold['proj_y']).max() <= 1
    assert np.array_equal(scan.unproj_range.cpu().numpy(), gold['unproj_range'])      # sqrt of a 3-term sum: exact
    idx = scan.proj_idx.cpu().numpy()
    same = idx == gold['proj_idx']
    assert (~same).sum() <= 24, int((~same).sum())     # the pixels the moved points left / entered
    for k in ('proj_range', 'proj_semantic', 'proj_mask'):
        a, b = getattr(scan, k).cpu().numpy(), gold[k]
        assert np.array_equal(a[same], b[same]), k
    for k in ('proj_xyz', 'proj_rgb'):
        a, b = getattr(scan, k).cpu().numpy(), gold[k]
        assert np.array_equal(a[same], b[same]), k
    # the nearest point wins a pixel; the reference's `proj_idx > 0` mask (point 0 never counts) is reproduced
    assert float(scan.proj_mask.sum()) == float((scan.proj_idx > 0).sum())
    occ = idx >= 0
    assert np.array_equal(scan.proj_range.cpu().numpy()[occ], scan.unproj_range.cpu().numpy()[idx[occ]])


def test_drop_selection_vs_reference():
    mg, gold, pts, sem, rgb, logits = _inputs()
    scan = _scan(mg, pts, sem, rgb)
    fmask = torch.from_numpy(gold['filter_sem']).cuda()       # the reference's own filter mask: isolates the selection
    rp, rl = raydrop.drop_rays(logits, scan, pts, sem, fmask, mask_thre=0.5)
    rp, rl = rp.cpu().numpy(), rl.cpu().numpy()
    want_p, want_l = gold['remain_points'], gold['remain_labels']
    assert abs(rp.shape[0] - want_p.shape[0]) <= 8, (rp.shape, want_p.shape)
    # order preserved: the survivors are a subsequence of the input cloud, equal to the reference's up to the
    # handful of border points
    got = {tuple(r) for r in rp.tolist()}
    want = {tuple(r) for r in want_p.tolist()}
    assert len(got ^ want) <= 16, len(got ^ want)
    assert not np.any(rl == 10) and not np.any((rl == 0) & (rp[:, 2] < -3))
    src = pts.cpu().numpy()
    pos = {tuple(r): i for i, r in enumerate(src.tolist())}
    order = [pos[tuple(r)] for r in rp.tolist()]
    assert all(a <= b for a, b in zip(order, order[1:]))       # (<=: the duplicated return maps to one key)
    # nothing survives an impossible threshold; empty clouds are fine
    rp0, _ = raydrop.drop_rays(logits, scan, pts, sem, fmask, mask_thre=1.5)
    assert rp0.shape[0] == 0
