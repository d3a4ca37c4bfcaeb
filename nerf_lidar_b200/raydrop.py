"""Stage-3 ray-drop, the parts around the U-Net, on the GPU -- `R/` = NeRF_LiDAR/NeRF_Lidar_code/:

  depth_filter(points, points_semantic, ...)   R/src/depth_filter.py:4-31
  LaserScan(H, W, fov_up, fov_down)            R/src/lidar_utils.py:57-275 (set_points / do_range_projection and the
                                               proj_* / unproj_range attributes the drop step reads)
  drop_rays(logits, scan, points, labels, ...) R/src/drop_simulation_rays.py:104-140 (save_near=False)

Same names, arguments and attribute layout as the reference, CUDA tensors instead of numpy arrays; kernels in
csrc/raydrop.cu.  The U-Net (R/src/unet/) is not built: its [2, H, W] logits are an input."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from ._lib import NlbRangeImage, check, f32, load, ptr, stream


def depth_filter(points: torch.Tensor, points_semantic: Optional[torch.Tensor] = None, return_mask: bool = False,
                 threshold: int = 1, radius: float = 1, width: int = 3, beams: int = 32):
    """R/src/depth_filter.py: `points` is a beam-major sweep [beams * W, 3]."""
    pts = f32(points).reshape(-1, 3)
    n = pts.shape[0]
    if n % beams:
        raise RuntimeError(f'depth_filter: {n} points are not a {beams}-beam sweep')
    sem = None if points_semantic is None else f32(points_semantic).reshape(-1)
    mask = torch.empty(n, dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        check(load().nlb_depth_filter(ptr(pts), ptr(sem), beams, n // beams, int(width), float(radius), int(threshold),
                                      ptr(mask), stream()))
    mask = mask.bool()
    return mask if return_mask else pts[mask].reshape(-1, 3)


class LaserScan:
    """R/src/lidar_utils.py:57-275 -- spherical range-image projection of a point cloud."""

    def __init__(self, project: bool = False, H: int = 64, W: int = 1024, fov_up: float = 3.0, fov_down: float = -25.0,
                 device='cuda'):
        self.project, self.proj_H, self.proj_W = project, H, W
        self.proj_fov_up, self.proj_fov_down = fov_up, fov_down
        self.device = torch.device(device)
        self.points = self.semantic = self.rgb = None

    def set_points(self, points, remissions=None, semantic=None, rgb=None):
        self.points = f32(points).reshape(-1, 3)
        self.semantic = None if semantic is None else f32(semantic).reshape(-1)
        self.rgb = None if rgb is None else f32(rgb).reshape(-1, 3)
        if self.project:
            self.do_range_projection()

    def size(self):
        return 0 if self.points is None else self.points.shape[0]

    __len__ = size

    def do_range_projection(self):
        H, W, dev = self.proj_H, self.proj_W, self.points.device
        n = self.points.shape[0]
        z = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
        self.proj_range, self.proj_xyz, self.proj_semantic = z(H, W), z(H, W, 3), z(H, W)
        self.proj_rgb, self.proj_idx, self.proj_mask = z(H, W, 3), z(H, W, dt=torch.int32), z(H, W)
        self.proj_x, self.proj_y, self.unproj_range = z(n, dt=torch.int32), z(n, dt=torch.int32), z(n)
        lib = load()
        ws = torch.empty(lib.nlb_range_projection_workspace_bytes(H, W), dtype=torch.uint8, device=dev)
        img = NlbRangeImage(ptr(self.proj_range), ptr(self.proj_xyz), ptr(self.proj_semantic), ptr(self.proj_rgb),
                            ptr(self.proj_idx), ptr(self.proj_mask))
        with torch.cuda.device(dev):
            check(lib.nlb_range_projection(ptr(self.points), ptr(self.semantic), ptr(self.rgb), n, H, W,
                                           float(self.proj_fov_up), float(self.proj_fov_down), ptr(self.proj_x),
                                           ptr(self.proj_y), ptr(self.unproj_range), C.byref(img), ptr(ws), stream()))


def drop_rays(pred_logits: torch.Tensor, laser_scan: LaserScan, points: torch.Tensor, points_semantic: torch.Tensor,
              depth_filter_mask: Optional[torch.Tensor] = None, mask_thre: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """The selection of R/src/drop_simulation_rays.py:104-140 (save_near=False) for one sweep: U-Net logits [2,H,W]
    -> (remain_points [m,3], remain_labels [m]), order preserved.  One host read (the survivor count) sizes the
    returned views."""
    pts, lab = f32(points).reshape(-1, 3), f32(points_semantic).reshape(-1)
    n, H, W = pts.shape[0], laser_scan.proj_H, laser_scan.proj_W
    lg = f32(pred_logits).reshape(2, H, W)
    fm = None if depth_filter_mask is None else depth_filter_mask.reshape(-1).to(torch.uint8).contiguous()
    out_p, out_l = torch.empty(n, 3, device=pts.device), torch.empty(n, device=pts.device)
    count = torch.zeros(1, dtype=torch.int32, device=pts.device)
    lib = load()
    ws = torch.empty(lib.nlb_raydrop_select_workspace_bytes(n), dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        check(lib.nlb_raydrop_select(ptr(lg), float(mask_thre), ptr(laser_scan.proj_mask), ptr(laser_scan.proj_x),
                                     ptr(laser_scan.proj_y), ptr(fm), ptr(pts), ptr(lab), n, H, W, ptr(out_p), ptr(out_l),
                                     ptr(count), ptr(ws), stream()))
    m = int(count.item())
    return out_p[:m], out_l[:m]
