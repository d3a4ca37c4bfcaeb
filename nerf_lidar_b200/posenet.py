"""Pose refinement around the hot path: the per-sensor pose corrections of Z/internal/posenet_v2.py:78-122
(`LearnPose`) and their application to a ray batch (Z/train.py:200-240).

While `start_step < step < end_step` the corrections are trained: origins / directions / viewdirs / base_x /
base_y become functions of the correction, and the hot path has to return gradients w.r.t. the ray geometry
(`nlb_encode_input_backward`, `nlb_prop_input_backward`, the |d| term of the compositing and the view-direction
columns of the NerfMLP, see ops.py).  After the window the trained corrections are applied without gradients;
before it the batch is used as loaded.

State-dict keys (`r`, `t`, `init_c2w`) are the reference's, so its `posenet_ckpt_*` files load unchanged."""
from typing import Dict, Optional

import torch
import torch.nn as nn

from .configs import Config

RAY_ROTATED = ('directions', 'viewdirs', 'base_x', 'base_y')


def so3_exp(r: torch.Tensor) -> torch.Tensor:
    """Axis-angle [N,3] -> rotation [N,3,3] by Rodrigues' formula, with the reference's 1e-15 guard on the angle
    (posenet_v2.py:44-54): I + sin(a)/a K + (1-cos(a))/a^2 K^2."""
    x, y, z = r.unbind(-1)
    o = torch.zeros_like(x)
    K = torch.stack([torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], -2)
    a = (r.norm(dim=-1) + 1e-15)[:, None, None]
    eye = torch.eye(3, dtype=r.dtype, device=r.device).expand_as(K)
    return eye + (torch.sin(a) / a) * K + ((1 - torch.cos(a)) / a ** 2) * (K @ K)


class LearnPose(nn.Module):
    """One (axis-angle, translation) correction per camera and per LiDAR (posenet_v2.py:78-122)."""

    def __init__(self, num_cams: int, num_lidars: int = 1, learn_R: bool = True, learn_t: bool = True,
                 init_c2w=None, t_ratio: float = 1.):
        super().__init__()
        self.num_cams, self.num_lidars, self.t_ratio = num_cams, num_lidars, t_ratio
        self.init_c2w = None
        if init_c2w is not None:
            self.init_c2w = nn.Parameter(torch.as_tensor(init_c2w), requires_grad=False)
        n = num_cams + num_lidars
        self.r = nn.Parameter(torch.zeros(n, 3), requires_grad=learn_R)
        self.t = nn.Parameter(torch.zeros(n, 3), requires_grad=learn_t)

    def forward(self, cam_id: torch.Tensor, transform_only: bool = False) -> torch.Tensor:
        """[B] sensor indices -> [B,4,4] correction (composed with init_c2w when one was given)."""
        R = so3_exp(self.r)
        top = torch.cat([R, (self.t * self.t_ratio)[:, :, None]], -1)
        bottom = torch.zeros(top.shape[0], 1, 4, dtype=top.dtype, device=top.device)
        bottom[:, 0, 3] = 1.
        # index_select, not c2w[cam_id]: the backward of advanced indexing sorts the 10 240 ray indices and
        # serialises on the few sensor rows (measured 1.7 ms of a 9 ms step); index_select's is one index_add_
        c2w = torch.cat([top, bottom], 1).index_select(0, cam_id.long())
        if not transform_only and self.init_c2w is not None:
            c2w = c2w @ self.init_c2w.index_select(0, cam_id.long())
        return c2w


class Track_opt(nn.Module):
    """Per object and track entry a yaw and a centre correction (posenet_v2.py:65-76; same parameter names)."""

    def __init__(self, bboxes: torch.Tensor, learn_R: bool = True, learn_t: bool = True):
        super().__init__()
        n_obj, n_ts, _ = bboxes.shape
        self.init_bboxes = bboxes
        self.opt_r = nn.Parameter(torch.zeros(n_obj, n_ts, 1), requires_grad=learn_R)
        self.opt_t = nn.Parameter(torch.zeros(n_obj, n_ts, 3), requires_grad=learn_t)
        self.tracks = bboxes

    def forward(self):
        return self.opt_r, self.opt_t


def refined_track(tracknet: Track_opt, device=None) -> torch.Tensor:
    """Z/train.py:251-256: track[:, :, :3] += opt_t, track[:, :, 3:4] += opt_r (a new tensor carrying the graph)."""
    raw = tracknet.tracks.to(device if device is not None else tracknet.opt_r.device)
    return torch.cat([raw[:, :, :3] + tracknet.opt_t, raw[:, :, 3:4] + tracknet.opt_r, raw[:, :, 4:]], -1)


def track_window(config: Config, step: int) -> Optional[str]:
    """'train' inside track_start_opt < step < track_start_opt + 5000, 'apply' after it, None before / when off."""
    if not config.track_refine:
        return None
    if config.track_start_opt < step < config.track_start_opt + 5000:
        return 'train'
    if step > config.track_start_opt + 5000:
        return 'apply'
    return None


def create_tracknet(config: Config, tracks: torch.Tensor, device=None):
    """(tracknet, optimizer, lr_fn) of Z/internal/train_utils.py:303-326."""
    from .train import learning_rate_decay
    net = Track_opt(tracks.to(device) if device is not None else tracks)
    if device is not None:
        net = net.to(device)
    start = config.track_start_opt

    def lr_fn(step):
        return learning_rate_decay(step - start, config.tn_lr_init, config.tn_lr_final, config.max_steps - start,
                                   config.lr_delay_steps, config.lr_delay_mult)

    params = list(net.parameters())
    on_gpu = params[0].is_cuda
    lr = torch.tensor(config.tn_lr_init, device=params[0].device) if on_gpu else config.tn_lr_init
    opt = torch.optim.Adam(params, lr=lr, betas=(config.adam_beta1, config.adam_beta2), eps=config.adam_eps,
                           capturable=on_gpu)
    return net, opt, lr_fn


def refine_rays(batch: Dict[str, torch.Tensor], posenet: LearnPose) -> Dict[str, torch.Tensor]:
    """Z/train.py:208-221: origins += t, every direction-like field rotated by R (row-wise R v).  Returns a
    new dictionary; run under torch.no_grad() after the window."""
    out = dict(batch)
    pose = posenet(batch['glo_idx'].reshape(-1))
    R, t = pose[:, :3, :3], pose[:, :3, 3]
    out['origins'] = batch['origins'] + t
    keys = RAY_ROTATED + (('normals',) if 'normals' in batch else ())
    for k in keys:
        out[k] = (batch[k].reshape(-1, 1, 3) * R).sum(-1)
    return out


def pose_window(config: Config, step: int) -> Optional[str]:
    """'train' inside the refinement window, 'apply' after it, None before it or with pose_refine off."""
    if not config.pose_refine:
        return None
    if config.start_step < step < config.end_step:
        return 'train'
    if step > config.end_step:
        return 'apply'
    return None


def create_posenet(num_poses: int, config: Config, num_lidars: int = 0, device=None):
    """(posenet, optimizer, lr_fn) of Z/internal/train_utils.py:278-301.  On a CUDA device the Adam is built
    capturable with its learning rate in a device scalar, so the trainer can record the window's step -- the
    corrections' forward, the ray-geometry gradients and this update -- in its CUDA graph."""
    from .train import learning_rate_decay
    net = LearnPose(num_poses, num_lidars=num_lidars, t_ratio=config.t_ratio, learn_R=config.learn_R,
                    learn_t=config.learn_t)
    if device is not None:
        net = net.to(device)

    def lr_fn(step):
        return learning_rate_decay(step - config.start_step, config.pn_lr_init, config.pn_lr_final,
                                   config.end_step - config.start_step, config.lr_delay_steps, config.lr_delay_mult)

    params = [p for p in net.parameters() if p.requires_grad]
    on_gpu = bool(params) and params[0].is_cuda
    lr = torch.tensor(config.pn_lr_init, device=params[0].device) if on_gpu else config.pn_lr_init
    opt = torch.optim.Adam(params, lr=lr, betas=(config.adam_beta1, config.adam_beta2), eps=config.adam_eps,
                           capturable=on_gpu)
    return net, opt, lr_fn
