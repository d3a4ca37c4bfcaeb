"""Synthetic nuScenes-shaped inputs and random-init weights for the zipnerf hot
path (there is no dataset or checkpoint in the build environment).

Everything here is numpy with a PCG64 generator so the same seed gives the same
bits in the build container and on the GPU box (golden fixtures depend on it).

Ray schema = the batch dict the reference's data layer hands to Model.forward
(reference: internal/camera_utils.py:454-617 `pixels_to_rays`/`cast_ray_batch`,
internal/lidar_utils.py:8-33 `cast_lidar_ray_batch`, internal/datasets.py:352-403
batch composition, :483-535 / :618-638 labels).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

# nuScenes front camera and scene scale used throughout SURVEY.md 8(d)
IMG_W, IMG_H, FOCAL = 1600, 900, 1266.0
SCENE_SCALE = 1.0 / 60.0
NEAR, FAR = 2.0 * SCENE_SCALE, 500.0 * SCENE_SCALE

# 32-beam elevation table (reference: internal/lidar_utils.py:36-38)
LIDAR_ELEVATIONS_DEG = sorted([
    -30.67, -9.33, -29.33, -8.00, -28.00, -6.67, -26.67, -5.33, -25.33, -4.00, -24.00, -2.67,
    -22.67, -1.33, -21.33, 0.00, -20.00, 1.33, -18.67, 2.67, -17.33, 4.00, -16.00, 5.33,
    -14.67, 6.67, -13.33, 8.00, -12.00, 9.33, -10.67, 10.67])


def _pose(rng: np.random.Generator, n: int):
    """Ego poses: origins on the x axis in [-1,1], small y/z noise; cameras look
    along +x with a small yaw/pitch perturbation (OpenGL camera axes)."""
    o = np.stack([rng.uniform(-1, 1, n), rng.normal(0, 0.01, n), rng.normal(0, 0.01, n)], -1)
    yaw = rng.normal(0, 0.05, n)
    pitch = rng.normal(0, 0.02, n)
    # camera -z looks along world +x; camera x = world -y; camera y = world z
    fwd = np.stack([np.cos(yaw) * np.cos(pitch), np.sin(yaw) * np.cos(pitch), np.sin(pitch)], -1)
    up0 = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up0)
    right /= np.linalg.norm(right, axis=-1, keepdims=True)
    up = np.cross(right, fwd)
    R = np.stack([right, up, -fwd], -1)  # columns = camera axes in world
    return o, R


def _pix_to_rays(px, py, R, o):
    """pixels_to_rays for a pinhole camera (camera_utils.py:489-564)."""

    def cam_dir(x, y):
        d = np.stack([(x + 0.5 - IMG_W / 2) / FOCAL, (y + 0.5 - IMG_H / 2) / FOCAL, np.ones_like(x)], -1)
        return d * np.array([1.0, -1.0, -1.0])  # OpenCV -> OpenGL

    def rot(d):
        return np.einsum('nij,nj->ni', R, d)

    d0, dx, dy = rot(cam_dir(px, py)), rot(cam_dir(px + 1, py)), rot(cam_dir(px, py + 1))
    viewdirs = d0 / np.linalg.norm(d0, axis=-1, keepdims=True)
    ex, ey = dx - d0, dy - d0
    nx, ny = np.linalg.norm(ex, axis=-1), np.linalg.norm(ey, axis=-1)
    radii = (0.5 * (nx + ny))[:, None] * 2 / math.sqrt(12)
    return dict(origins=o, directions=d0, viewdirs=viewdirs, radii=radii,
                base_x=ex / nx[:, None], base_y=ey / ny[:, None])


def camera_rays(rng: np.random.Generator, n_patch_rays: int, n_pixel_rays: int, patch: int = 32):
    """`n_patch_rays` rays in contiguous patch x patch blocks followed by
    `n_pixel_rays` random pixels (datasets.py:356-366)."""
    parts = []
    n_patches = n_patch_rays // (patch * patch)
    if n_patches:
        o, R = _pose(rng, n_patches)
        x0 = rng.integers(0, IMG_W - patch, n_patches)
        y0 = rng.integers(0, IMG_H - patch, n_patches)
        yy, xx = np.meshgrid(np.arange(patch), np.arange(patch), indexing='ij')
        px = (x0[:, None, None] + xx).reshape(-1).astype(np.float64)
        py = (y0[:, None, None] + yy).reshape(-1).astype(np.float64)
        rep = patch * patch
        parts.append(_pix_to_rays(px, py, np.repeat(R, rep, 0), np.repeat(o, rep, 0)))
    if n_pixel_rays:
        o, R = _pose(rng, n_pixel_rays)
        px = rng.integers(0, IMG_W, n_pixel_rays).astype(np.float64)
        py = rng.integers(0, IMG_H, n_pixel_rays).astype(np.float64)
        parts.append(_pix_to_rays(px, py, R, o))
    out = {k: np.concatenate([p[k] for p in parts], 0) for k in parts[0]}
    out['patch_mask'] = np.concatenate([np.ones(n_patches * patch * patch), np.zeros(n_pixel_rays)])
    return out


def lidar_directions(width: int = 1084):
    """32 x `width` unit directions (lidar_utils.py:60 azimuth, :559-568
    get_directions: right / forward / up)."""
    az = np.linspace(270.0, -90.0, width) / 180.0 * np.pi
    el = np.array(LIDAR_ELEVATIONS_DEG) / 180.0 * np.pi
    th, ph = np.meshgrid(el, az, indexing='ij')
    d = np.stack([np.cos(th) * np.sin(ph), np.cos(th) * np.cos(ph), np.sin(th)], -1)
    return d.reshape(-1, 3)


def lidar_rays(rng: np.random.Generator, n: Optional[int] = None, width: int = 1084):
    """LiDAR rays with the reference's input quirks (lidar_utils.py:8-33):
    viewdirs = directions / GLOBAL Frobenius norm, base_x = base_y = directions,
    radii = 5e-4.  n=None -> one full 32 x width sweep from one origin."""
    dirs = lidar_directions(width)
    if n is None:
        o, _ = _pose(rng, 1)
        o = np.repeat(o, dirs.shape[0], 0)
    else:
        dirs = dirs[rng.integers(0, dirs.shape[0], n)]
        o, _ = _pose(rng, n)
    return dict(origins=o, directions=dirs, viewdirs=dirs / np.linalg.norm(dirs),
                radii=np.full((dirs.shape[0], 1), 5e-4), base_x=dirs.copy(), base_y=dirs.copy(),
                patch_mask=np.zeros(dirs.shape[0]))


def sensor_index(batch, num_cams: int = 1):
    """`glo_idx` of the nuScenes loader (Z/internal/datasets.py:632): the index of the sensor a ray belongs to
    in LearnPose's table -- here camera 0 for the image rays, the first LiDAR slot for the sweep rays."""
    lidar = np.asarray(batch['lidar_mask']).reshape(-1) > 0
    return np.where(lidar, num_cams, 0).astype(np.int32)


def _finish(rays: Dict[str, np.ndarray], rng, lidar_mask: np.ndarray, labels: bool):
    n = rays['origins'].shape[0]
    b = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in rays.items()}
    b['near'] = np.full((n, 1), NEAR, np.float32)
    b['far'] = np.full((n, 1), FAR, np.float32)
    b['cam_idx'] = np.where(lidar_mask > 0, -1, 0).astype(np.float32)[:, None]
    b['lossmult'] = np.ones((n, 1), np.float32)
    b['timestamp'] = np.zeros((n, 1), np.float32)
    b['lidar_mask'] = lidar_mask.astype(np.float32)
    if labels:
        lm = lidar_mask > 0
        rgb = rng.uniform(0, 1, (n, 3)).astype(np.float32)
        rgb[lm] = 0
        b['rgb'] = rgb
        depth = rng.uniform(NEAR * 2, FAR * 0.5, n).astype(np.float32)
        depth[(~lm) & (rng.uniform(0, 1, n) < 0.5)] = 0  # sparse projected depth on pixels
        b['depth'] = depth
        sem = rng.integers(0, 19, n).astype(np.float32)
        sem[rng.uniform(0, 1, n) < 0.05] = 255
        sem[lm] = 255
        b['semantic'] = sem
        b['intensity'] = np.where(lm, rng.uniform(0, 1, n), 0).astype(np.float32)
        # dataset mask (datasets.py:492,505,624): 1 = static background / LiDAR return, 0 = pixels on moving
        # objects, which Z/train.py:287,307 exclude from the losses when Config.instance_obj is off
        # (drawn last so that the arrays above keep the values the golden fixtures were generated with)
        b['mask'] = np.where(lm | (rng.uniform(0, 1, n) >= 0.1), 1.0, 0.0).astype(np.float32)
    return b


def make_train_batch(batch_size: int = 8192, seed: int = 0, lidar_batch_ratio: int = 4,
                     patch: int = 32, labels: bool = True) -> Dict[str, np.ndarray]:
    """Per-rank training batch exactly as datasets.py:352-403 composes it:
    batch_size/4 patch rays + 3/4 random pixels + batch_size/ratio EXTRA LiDAR
    rays (8192 -> 10240 rays through the model)."""
    rng = np.random.default_rng(seed)
    n_patch = (batch_size // 4) // (patch * patch) * patch * patch
    cam = camera_rays(rng, n_patch, batch_size - n_patch, patch)
    n_lidar = batch_size // lidar_batch_ratio if lidar_batch_ratio > 0 else 0
    if n_lidar:
        lid = lidar_rays(rng, n_lidar)
        rays = {k: np.concatenate([cam[k], lid[k]], 0) for k in cam}
    else:
        rays = cam
    lm = np.concatenate([np.zeros(batch_size), np.ones(n_lidar)])
    return _finish(rays, rng, lm, labels)


def make_lidar_sweep(seed: int = 0, width: int = 1084) -> Dict[str, np.ndarray]:
    """BASELINE config 3: one full 32 x width sweep (render_lidar.py:106-114)."""
    rng = np.random.default_rng(seed)
    rays = lidar_rays(rng, None, width)
    return _finish(rays, rng, np.ones(rays['origins'].shape[0]), labels=False)


def make_camera_frame(seed: int = 0, height: int = IMG_H, width: int = IMG_W) -> Dict[str, np.ndarray]:
    """One full pinhole frame, rays in row-major pixel order."""
    rng = np.random.default_rng(seed)
    o, R = _pose(rng, 1)
    yy, xx = np.meshgrid(np.arange(height), np.arange(width), indexing='ij')
    n = height * width
    rays = _pix_to_rays(xx.reshape(-1).astype(np.float64), yy.reshape(-1).astype(np.float64),
                        np.repeat(R, n, 0), np.repeat(o, n, 0))
    rays['patch_mask'] = np.zeros(n)
    return _finish(rays, rng, np.zeros(n), labels=False)


def to_torch(batch: Dict[str, np.ndarray], device='cpu', pin: bool = False):
    out = {}
    for k, v in batch.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        if pin:
            t = t.pin_memory()
        out[k] = t.to(device, non_blocking=pin) if str(device) != 'cpu' else t
    return out


# ----------------------------------------------------------------------------
# random-init weights with the reference's state-dict key names
# ----------------------------------------------------------------------------
def grid_offsets(num_levels, base_resolution=16, per_level_scale=2.0, log2_hashmap_size=21,
                 input_dim=3, align_corners=False):
    """Level sizing of GridEncoder.__init__ (gridencoder/grid.py:120-141)."""
    offsets, resolutions, offset = [], [], 0
    cap = 2 ** log2_hashmap_size
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        res = res if align_corners else res + 1
        n = int(np.ceil(min(cap, res ** input_dim) / 8) * 8)
        resolutions.append(res)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, np.int32), np.array(resolutions, np.int32)


def _linear(rng, out_f, in_f, kaiming_relu=False):
    bw = math.sqrt(6.0 / in_f) if kaiming_relu else 1.0 / math.sqrt(in_f)
    bb = 1.0 / math.sqrt(in_f)
    return (rng.uniform(-bw, bw, (out_f, in_f)).astype(np.float32),
            rng.uniform(-bb, bb, (out_f,)).astype(np.float32))


def init_state_dict(seed: int = 0, table_std: float = 1e-4, use_intensity: bool = True,
                    small_tables: bool = False) -> Dict[str, torch.Tensor]:
    """State dict of the nuscenes_single.gin model (77 656 777 parameters) with
    nn.Linear-style uniform init and tables ~ U(-table_std, table_std)
    (grid.py:151-153).  `small_tables` uses log2_hashmap_size=15 for CPU tests."""
    rng = np.random.default_rng(seed)
    sd = {}
    log2 = 15 if small_tables else 21

    def table(prefix, L, C):
        offs, res = grid_offsets(L, log2_hashmap_size=log2)
        rows = int(offs[-1])
        sd[prefix + 'encoder.embeddings'] = torch.from_numpy(
            (rng.random((rows, C), dtype=np.float32) * 2 - 1) * np.float32(table_std))
        sd[prefix + 'encoder.offsets'] = torch.from_numpy(offs)
        sd[prefix + 'encoder.grid_sizes'] = torch.from_numpy(res)

    def lin(name, out_f, in_f, kaiming_relu=False):
        w, b = _linear(rng, out_f, in_f, kaiming_relu)
        sd[name + '.weight'], sd[name + '.bias'] = torch.from_numpy(w), torch.from_numpy(b)

    table('nerf_mlp.', 10, 4)
    lin('nerf_mlp.density_layer.0', 64, 40)
    lin('nerf_mlp.density_layer.2', 256, 64)
    lin('nerf_mlp.lin_second_stage_0', 256, 283, True)
    lin('nerf_mlp.lin_second_stage_1', 256, 539, True)
    lin('nerf_mlp.rgb_layer', 3, 256)
    lin('nerf_mlp.sem_layer.0', 64, 256)
    lin('nerf_mlp.sem_layer.2', 19, 64)
    if use_intensity:
        lin('nerf_mlp.intensity_layer.0', 64, 256)
        lin('nerf_mlp.intensity_layer.2', 1, 64)
    for i, L in enumerate((6, 8)):
        table(f'prop_mlp_{i}.', L, 1)
        lin(f'prop_mlp_{i}.density_layer.0', 64, L)
        lin(f'prop_mlp_{i}.density_layer.2', 1, 64)
    return sd


def make_rand_inputs(n_rays: int, samples=(64, 64, 32), n_multi: int = 7, seed: int = 0):
    """Injected random draws of one forward pass with rand=True: per level the
    single-jitter u[N,1] (stepfun.py:215-216) and the multisample rotation noise
    [N,S,7] (render.py:149-150), both U[0,1)."""
    rng = np.random.default_rng(seed + 7919)
    out = []
    for S in samples:
        out.append(dict(jitter=rng.random((n_rays, 1), dtype=np.float32),
                        deg=rng.random((n_rays, S, n_multi), dtype=np.float32)))
    return out
