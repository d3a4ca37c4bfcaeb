"""On-GPU ray generation and batch assembly.

The reference assembles every training batch on the host with numpy inside an 8-worker DataLoader
(Z/train.py:111-118 -> Z/internal/datasets.py:352-403 `next_train`, :707-749 `_next_train`, :431-535
`_make_ray_batch`, :585-640 `_make_lidar_ray_batch`).  Here the cameras, images / labels and the LiDAR tables
stay resident in HBM; pixels are drawn with the device RNG, rays come from the CUDA kernels in
csrc/raygen.cu, labels are device gathers, and nothing crosses PCIe per step.

Function names, argument meaning and return keys follow the reference:
  pixels_to_rays / cast_ray_batch   Z/internal/camera_utils.py:454-564 / :567-617
  get_directions                    Z/internal/lidar_utils.py:559-568
  cast_lidar_ray_batch              Z/internal/lidar_utils.py:8-33
There is no CPU path: CUDA tensors in, CUDA tensors out (the library raises otherwise)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from ._lib import NlbRayOut, check, load, ptr, stream


def _i32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError('nerf_lidar_b200: expected a CUDA tensor (there is no CPU path)')
    return t.to(torch.int32).contiguous()


def _f64(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError('nerf_lidar_b200: expected a CUDA tensor (there is no CPU path)')
    return t.to(torch.float64).contiguous()


def _alloc(n: int, device, imageplane: bool = True) -> Dict[str, torch.Tensor]:
    out = {k: torch.empty(n, 3, device=device) for k in ('origins', 'directions', 'viewdirs', 'base_x', 'base_y')}
    out['radii'] = torch.empty(n, 1, device=device)
    out['imageplane'] = torch.empty(n, 2, device=device) if imageplane else None
    return out


def _out_struct(o: Dict[str, torch.Tensor]) -> NlbRayOut:
    return NlbRayOut(ptr(o['origins']), ptr(o['directions']), ptr(o['viewdirs']), ptr(o['radii']),
                     ptr(o['imageplane']), ptr(o['base_x']), ptr(o['base_y']))


def pixels_to_rays(pix_x_int, pix_y_int, pixtocams, camtoworlds, cam_idx=None, distortion_params=None,
                   pixtocam_ndc=None):
    """camera_utils.pixels_to_rays for a perspective camera.  `pixtocams` [3,3] or [ncam,3,3], `camtoworlds`
    [3,4] or [ncam,3,4]; with stacked matrices `cam_idx` (same shape as the pixels) selects one per ray --
    the reference indexes the stacks on the host before the call (`batch_index`, camera_utils.py:591).
    Returns (origins, directions, viewdirs, radii, imageplane, base_x, base_y), pixel shape + [3|1|2]."""
    if distortion_params is not None or pixtocam_ndc is not None:
        raise NotImplementedError('raygen: lens distortion / NDC cameras are not built (the nuScenes loader uses neither)')
    shape = tuple(pix_x_int.shape)
    px, py = _i32(pix_x_int).reshape(-1), _i32(pix_y_int).reshape(-1)
    n = px.numel()
    p2c, c2w = _f64(pixtocams).reshape(-1, 3, 3), _f64(camtoworlds).reshape(-1, 3, 4)
    ci = None
    if cam_idx is not None:
        ci = _i32(torch.broadcast_to(cam_idx, shape)).reshape(-1)
    o = _alloc(n, px.device)
    st = _out_struct(o)
    check(load().nlb_camera_rays(ptr(px), ptr(py), ptr(ci), ptr(p2c), p2c.shape[0], ptr(c2w), c2w.shape[0], n,
                                 C.byref(st), stream()))
    r = lambda k, w: o[k].reshape(shape + (w,))
    return (r('origins', 3), r('directions', 3), r('viewdirs', 3), r('radii', 1), r('imageplane', 2),
            r('base_x', 3), r('base_y', 3))


def cast_ray_batch(cameras, pixels: Dict[str, torch.Tensor], camtype=None, patch_size=None) -> Dict[str, Optional[torch.Tensor]]:
    """camera_utils.cast_ray_batch: `cameras` = (pixtocams, camtoworlds, distortion_params, pixtocam_ndc),
    `pixels` carries pix_x_int / pix_y_int / cam_idx[...,1] and the per-ray metadata that is passed through."""
    pixtocams, camtoworlds, distortion_params, pixtocam_ndc = cameras
    cam_idx = pixels['cam_idx'][..., 0]
    origins, directions, viewdirs, radii, imageplane, base_x, base_y = pixels_to_rays(
        pixels['pix_x_int'], pixels['pix_y_int'], pixtocams, camtoworlds, cam_idx=cam_idx,
        distortion_params=distortion_params, pixtocam_ndc=pixtocam_ndc)
    return dict(origins=origins, directions=directions, viewdirs=viewdirs, radii=radii, imageplane=imageplane,
                lossmult=pixels.get('lossmult'), near=pixels.get('near'), far=pixels.get('far'),
                cam_idx=pixels.get('cam_idx'), exposure_idx=pixels.get('exposure_idx'),
                exposure_values=pixels.get('exposure_values'), base_x=base_x, base_y=base_y)


def get_directions(vertical_angles, horizontal_angles, device='cuda') -> torch.Tensor:
    """lidar_utils.get_directions: [len(vertical) * len(horizontal), 3] fp32, beam-major; vertical angles in
    degrees, horizontal angles in radians (as the reference's callers pass them)."""
    el = torch.as_tensor(vertical_angles, dtype=torch.float64).to(device).contiguous()
    az = torch.as_tensor(horizontal_angles, dtype=torch.float64).to(device).contiguous()
    out = torch.empty(el.numel() * az.numel(), 3, device=el.device)
    check(load().nlb_lidar_directions(ptr(el), el.numel(), ptr(az), az.numel(), ptr(out), stream()))
    return out


def cast_lidar_ray_batch(lidar_origins: torch.Tensor, lidar_directions: torch.Tensor,
                         pixels: Dict[str, torch.Tensor]) -> Dict[str, Optional[torch.Tensor]]:
    """lidar_utils.cast_lidar_ray_batch incl. its two input quirks (viewdirs over the GLOBAL norm,
    base_x = base_y = directions)."""
    from ._lib import f32
    o_in, d_in = f32(lidar_origins).reshape(-1, 3), f32(lidar_directions).reshape(-1, 3)
    n = d_in.shape[0]
    o = _alloc(n, d_in.device)
    ws = torch.empty(1, dtype=torch.float64, device=d_in.device)
    st = _out_struct(o)
    check(load().nlb_lidar_rays(ptr(o_in), ptr(d_in), n, ptr(ws), C.byref(st), stream()))
    return dict(origins=o['origins'], directions=o['directions'], viewdirs=o['viewdirs'], radii=o['radii'],
                imageplane=o['imageplane'], lossmult=pixels.get('lossmult'), near=pixels.get('near'),
                far=pixels.get('far'), cam_idx=pixels.get('cam_idx'), exposure_idx=pixels.get('exposure_idx'),
                exposure_values=pixels.get('exposure_values'), base_x=o['base_x'], base_y=o['base_y'])


_VECTOR_KEYS = ('origins', 'directions', 'viewdirs', 'radii', 'imageplane', 'base_x', 'base_y', 'lossmult', 'near',
                'far', 'cam_idx', 'rgb', 'timestamp')


class GpuRayLoader:
    """Device-resident replacement of the reference's training Dataset / DataLoader pair for the hot path's
    inputs: `next_train()` returns the batch dict of datasets.py:352-403 -- batch_size/4 rays in patch x patch
    blocks, the rest random pixels, plus batch_size/lidar_batch_ratio EXTRA LiDAR rays -- with every tensor
    already in HBM.

    images [ncam,H,W,3], depths / semantics / masks [ncam,H,W] (optional), cameras as `cast_ray_batch` takes
    them, `lidar_depends` = (distances [Nl], origins [Nl,3], directions [Nl,3], intensity [Nl] or None), all
    CUDA tensors.  Sampling follows `_next_train` (datasets.py:707-749, BatchingMethod.ALL_IMAGES): uniform
    patch corners / pixels / camera indices / LiDAR indices, drawn with the device generator."""

    def __init__(self, images, pixtocams, camtoworlds, near: float, far: float, depths=None, semantics=None,
                 masks=None, lidar_depends=None, timestamps=None, lidar_timestamps=None, lidar_frame_count: int = 1,
                 batch_size: int = 8192, patch_size: int = 32, lidar_batch_ratio: int = 4,
                 num_border_pixels_to_mask: int = 0, seed: int = 0):
        if not images.is_cuda:
            raise RuntimeError('GpuRayLoader: the dataset tensors must be resident on the GPU')
        self.images, self.depths, self.semantics, self.masks = images, depths, semantics, masks
        self.n_examples, self.height, self.width = images.shape[:3]
        self.cameras = (_f64(pixtocams), _f64(camtoworlds), None, None)
        self.near, self.far = float(near), float(far)
        self.lidar_depends = lidar_depends
        self.timestamps, self.lidar_timestamps = timestamps, lidar_timestamps
        self.lidar_frame_count = lidar_frame_count
        self.batch_size, self.patch_size, self.lidar_batch_ratio = batch_size, patch_size, lidar_batch_ratio
        self.border = num_border_pixels_to_mask
        self.gen = torch.Generator(device=images.device)
        self.gen.manual_seed(seed)
        dev = images.device
        yy, xx = torch.meshgrid(torch.arange(patch_size, device=dev), torch.arange(patch_size, device=dev), indexing='ij')
        self._patch_dx, self._patch_dy = xx, yy   # camera_utils.pixel_coordinates(patch, patch), 'xy' indexing

    def _randint(self, lo, hi, shape):
        return torch.randint(lo, hi, shape, device=self.images.device, generator=self.gen)

    def _make_ray_batch(self, pix_x_int, pix_y_int, cam_idx, patch_sample: bool):
        full = lambda v: torch.full(tuple(pix_x_int.shape) + (1,), float(v), device=pix_x_int.device)
        cam_b = torch.broadcast_to(cam_idx, pix_x_int.shape)
        pixels = dict(pix_x_int=pix_x_int, pix_y_int=pix_y_int, lossmult=full(1.0), near=full(self.near),
                      far=full(self.far), cam_idx=cam_b[..., None])
        batch = cast_ray_batch(self.cameras, pixels)
        batch['cam_idx'] = batch['cam_idx'].float()
        batch['rgb'] = self.images[cam_b, pix_y_int, pix_x_int]
        if self.depths is not None:
            batch['depth'] = self.depths[cam_b, pix_y_int, pix_x_int]
        if self.semantics is not None:
            batch['semantic'] = self.semantics[cam_b, pix_y_int, pix_x_int]
        ones = torch.ones(tuple(pix_x_int.shape), device=pix_x_int.device)
        batch['mask'] = self.masks[cam_b, pix_y_int, pix_x_int] if self.masks is not None else ones
        batch['timestamp'] = (self.timestamps[cam_b][..., None] if self.timestamps is not None
                              else torch.zeros_like(batch['near']))
        batch['lidar_mask'] = torch.zeros_like(ones)
        batch['patch_mask'] = ones.clone() if patch_sample else torch.zeros_like(ones)
        if self.lidar_depends is not None and self.lidar_depends[3] is not None:
            batch['intensity'] = torch.zeros_like(ones)
        return batch

    def _make_lidar_ray_batch(self, lidar_idx, lidar_frame_idx):
        dist, origins, directions, intensity = self.lidar_depends
        n = lidar_idx.numel()
        dev = lidar_idx.device
        full = lambda v: torch.full((n, 1), float(v), device=dev)
        pixels = dict(lossmult=full(1.0), near=full(self.near), far=full(self.far),
                      cam_idx=(self.n_examples + lidar_frame_idx).reshape(-1, 1).float())
        batch = cast_lidar_ray_batch(origins[lidar_idx], directions[lidar_idx], pixels)
        batch['rgb'] = torch.zeros(n, 3, device=dev)
        batch['depth'] = dist[lidar_idx].reshape(-1)
        batch['semantic'] = torch.full((n,), 255.0, device=dev)
        batch['mask'] = torch.ones(n, device=dev)
        if intensity is not None:
            batch['intensity'] = intensity[lidar_idx].reshape(-1)
        batch['lidar_mask'] = torch.ones(n, device=dev)
        batch['patch_mask'] = torch.zeros(n, device=dev)
        batch['timestamp'] = (self.lidar_timestamps[lidar_frame_idx].reshape(-1, 1) if self.lidar_timestamps is not None
                              else torch.zeros(n, 1, device=dev))
        return batch

    def _next_train(self, batch_size, patch_size, lidar_batch=0):
        if lidar_batch > 0:
            frame = self._randint(0, self.lidar_frame_count, (lidar_batch,))
            idx = self._randint(0, self.lidar_depends[0].shape[0], (lidar_batch,))
            return self._make_lidar_ray_batch(idx, frame)
        num_patches = batch_size // patch_size ** 2
        lower, upper = self.border, self.border + patch_size - 1
        px = self._randint(lower, self.width - upper, (num_patches, 1, 1))
        py = self._randint(lower, self.height - upper, (num_patches, 1, 1))
        if patch_size > 1:
            px, py = px + self._patch_dx, py + self._patch_dy
        cam = self._randint(0, self.n_examples, (num_patches, 1, 1))
        return self._make_ray_batch(px, py, cam, patch_sample=patch_size > 1)

    def next_train(self) -> Dict[str, torch.Tensor]:
        lidar_batch = self.batch_size // self.lidar_batch_ratio if (self.lidar_depends is not None and self.lidar_batch_ratio > 0) else 0
        if self.patch_size == 1:
            parts = [self._next_train(self.batch_size, 1)]
        else:
            patch_batch = self.batch_size // 4
            parts = [self._next_train(patch_batch, self.patch_size), self._next_train(self.batch_size - patch_batch, 1)]
        if lidar_batch:
            parts.append(self._next_train(self.batch_size, 1, lidar_batch=lidar_batch))
        batch = {}
        for key, first in parts[0].items():
            if first is None or any(p.get(key) is None for p in parts):
                continue
            # per-ray vectors keep their last axis, per-ray scalars are flattened (datasets.py:377-399)
            batch[key] = torch.cat([p[key].reshape(-1, p[key].shape[-1]) if key in _VECTOR_KEYS else p[key].reshape(-1)
                                    for p in parts], 0)
        return batch
