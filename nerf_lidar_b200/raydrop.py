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


# ----------------------------------------------------------------------------- the U-Net (R/src/unet/)
import torch.nn as nn

from ._lib import NlbUnetConv, NlbUnetWeights


class DoubleConv(nn.Module):
    """R/src/unet/unet_parts.py:8-26 (parameter container; evaluated by csrc/unet.cu)."""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid_channels = mid_channels or out_channels
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(mid_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))


class Down(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))


class Up(nn.Module):
    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
            self.conv = DoubleConv(in_channels, out_channels, in_channels // 2)
        else:
            self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
            self.conv = DoubleConv(in_channels, out_channels)


class OutConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)


class UNet(nn.Module):
    """R/src/unet/unet_model.py:6-47: same constructor, sub-module names and state-dict keys (reference checkpoints
    load as they are).  Inference (`.eval()`) runs on csrc/unet.cu; training mode evaluates the same layers with
    torch (see `_forward_torch`)."""

    def __init__(self, n_channels, n_classes, bilinear=False, regression=False):
        super().__init__()
        self.n_channels, self.n_classes, self.bilinear, self.regression = n_channels, n_classes, bilinear, regression
        self.tf32 = None      # None: follow torch.backends.cudnn.allow_tf32; True / False: force
        self.inc = DoubleConv(n_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        factor = 2 if bilinear else 1
        self.down4 = Down(512, 1024 // factor)
        self.up1 = Up(1024, 512 // factor, bilinear)
        self.up2 = Up(512, 256 // factor, bilinear)
        self.up3 = Up(256, 128 // factor, bilinear)
        self.up4 = Up(128, 64, bilinear)
        self.outc = OutConv(64, n_classes)
        if regression:
            self.outr = OutConv(64, 1)

    def _forward_torch(self, x):
        """Training mode (R/src/model/ray_drop_train.py:73-125 trains this network with torch.optim.Adam): the layers
        as torch evaluates them -- batch statistics in the BatchNorms, autograd, cuDNN -- exactly the reference's
        `UNet.forward` (unet_model.py:31-47, unet_parts.py:59-69).  Outside the hot path; inference runs on
        csrc/unet.cu."""
        import torch.nn.functional as F
        dc = lambda m, t: m.double_conv(t)
        xs = [dc(self.inc, x)]
        for d in (self.down1, self.down2, self.down3, self.down4):
            xs.append(dc(d.maxpool_conv[1], d.maxpool_conv[0](xs[-1])))
        y = xs[-1]
        for u, skip in zip((self.up1, self.up2, self.up3, self.up4), xs[-2::-1]):
            y = u.up(y)
            dy, dx = skip.shape[2] - y.shape[2], skip.shape[3] - y.shape[3]
            y = F.pad(y, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
            y = dc(u.conv, torch.cat([skip, y], dim=1))
        logits = self.outc.conv(y)
        return (logits, torch.sigmoid(self.outr.conv(y))) if self.regression else logits

    @staticmethod
    def _fold(dc: DoubleConv, keep, tf32: bool):
        """(conv weight, BatchNorm folded to scale / shift, operand-layout copy of the weight for the TF32 tensor-core
        path) of both halves of a DoubleConv."""
        out = []
        for ci, bi in ((0, 1), (3, 4)):
            conv, bn = dc.double_conv[ci], dc.double_conv[bi]
            scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias.detach() - bn.running_mean * scale
            t = [f32(conv.weight.detach()), f32(scale), f32(shift)]
            packed = None
            oc, c = t[0].shape[:2]
            if tf32 and oc % 64 == 0 and c % 32 == 0:
                packed = torch.empty_like(t[0])
                with torch.cuda.device(packed.device):
                    check(load().nlb_unet_pack_conv(ptr(t[0]), oc, c, ptr(packed), stream()))
                t.append(packed)
            keep += t
            out.append(NlbUnetConv(ptr(t[0]), ptr(t[1]), ptr(t[2]), ptr(packed)))
        return out

    def forward(self, x: torch.Tensor):
        """logits [N, n_classes, H, W], or (logits, reg [N, 1, H, W]) with regression=True."""
        if self.training:
            return self._forward_torch(x)
        with torch.no_grad():
            return self._forward_kernels(x)

    def _forward_kernels(self, x: torch.Tensor):
        x = f32(x)
        N, cin, H, W = x.shape
        if cin != self.n_channels:
            raise RuntimeError(f'UNet: expected {self.n_channels} input channels, got {cin}')
        # folded BatchNorm terms / weight pointers, rebuilt only when a parameter or buffer changed (18 layers x a
        # handful of tiny launches would otherwise cost more than the convolutions)
        # precision: the reference's Conv2d layers follow torch.backends.cudnn.allow_tf32 (True by default: TF32
        # tensor-core convolutions), and so does this module unless `self.tf32` says otherwise
        tf32 = torch.backends.cudnn.allow_tf32 if self.tf32 is None else bool(self.tf32)
        version = (tf32,) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        cached = self.__dict__.get('_nlb_folded')
        if cached is None or cached[0] != version:
            keep = []
            w = NlbUnetWeights()
            w.inc[0], w.inc[1] = self._fold(self.inc, keep, tf32)
            for i, d in enumerate((self.down1, self.down2, self.down3, self.down4)):
                w.down[i][0], w.down[i][1] = self._fold(d.maxpool_conv[1], keep, tf32)
            for i, u in enumerate((self.up1, self.up2, self.up3, self.up4)):
                w.up[i][0], w.up[i][1] = self._fold(u.conv, keep, tf32)
                if not self.bilinear:
                    t = [f32(u.up.weight.detach()), f32(u.up.bias.detach())]
                    keep += t
                    w.up_weight[i], w.up_bias[i] = ptr(t[0]), ptr(t[1])
                    cin_t, cout_t = t[0].shape[:2]
                    if tf32 and cout_t % 64 == 0 and cin_t % 32 == 0:
                        packed = torch.empty_like(t[0])
                        with torch.cuda.device(packed.device):
                            check(load().nlb_unet_pack_convtranspose(ptr(t[0]), cin_t, cout_t, ptr(packed), stream()))
                        keep.append(packed)
                        w.up_packed[i] = ptr(packed)
            t = [f32(self.outc.conv.weight.detach().reshape(self.n_classes, 64)), f32(self.outc.conv.bias.detach())]
            keep += t
            w.outc_weight, w.outc_bias = ptr(t[0]), ptr(t[1])
            if self.regression:
                t = [f32(self.outr.conv.weight.detach().reshape(1, 64)), f32(self.outr.conv.bias.detach())]
                keep += t
                w.outr_weight, w.outr_bias = ptr(t[0]), ptr(t[1])
            w.bilinear, w.n_classes = int(bool(self.bilinear)), int(self.n_classes)
            self.__dict__['_nlb_folded'] = cached = (version, w, keep)
        w = cached[1]
        lib = load()
        ws = torch.empty(lib.nlb_unet_workspace_bytes(N, H, W) // 4, device=x.device)
        out = torch.empty(N, self.n_classes, H, W, device=x.device)
        reg = torch.empty(N, 1, H, W, device=x.device) if self.regression else None
        with torch.cuda.device(x.device):
            check(lib.nlb_unet_forward(ptr(x), C.byref(w), N, cin, H, W, ptr(out), ptr(reg), ptr(ws), stream()))
        return (out, reg) if self.regression else out
