"""ORACLE (test infrastructure, not product code) -- CPU restatement of one
optimisation step of the reference's nuScenes run around the hot path:
losses of Z/train.py:283-455 + Z/internal/train_utils.py:55-181, autograd
backward, NaN scrub (train_utils.py:251-253) and torch.optim.Adam
(train_utils.py:256-275).  Used by tests (gradient / update parity of the fused
kernels) and by bench.py's cpu_baseline / `--impl reference` arm.

Masks follow Z/train.py:286-327: the dataset mask gates the rgb / depth / semantic terms and
the smoothness edges (`instance_obj=True` clears it, as the shipped gin does)."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from oracle import zipnerf_oracle as zo


def _interp_quad_masks(x, xp, fpdf, fcdf):
    """math.sorted_interp_quad, Z/internal/math.py:111-131 (mask formulation)."""
    mask = x[..., None, :] >= xp[..., :, None]

    def find(v):
        v0 = torch.max(torch.where(mask, v[..., None], v[..., :1, None]), -2).values
        v1 = torch.min(torch.where(~mask, v[..., None], v[..., -1:, None]), -2).values
        return v0, v1

    fpdf0, fpdf1 = find(fpdf)
    fcdf0, _ = find(fcdf)
    xp0, xp1 = find(xp)
    off = torch.clip(torch.nan_to_num((x - xp0) / (xp1 - xp0), 0), 0, 1)
    return fcdf0 + (x - xp0) * (fpdf0 + fpdf1 * off + fpdf0 * (1 - off)) / 2


def blur_stepfun(x, y, r):
    """Z/internal/stepfun.py:425-433."""
    xr, idx = torch.sort(torch.cat([x - r, x + r], -1))
    z = torch.zeros_like(y[..., :1])
    y1 = (torch.cat([y, z], -1) - torch.cat([z, y], -1)) / (2 * r)
    y2 = torch.cat([y1, -y1], -1).take_along_dim(idx[..., :-1], -1)
    yr = torch.cumsum((xr[..., 1:] - xr[..., :-1]) * torch.cumsum(y2, -1), -1).clamp_min(0)
    return xr, torch.cat([torch.zeros_like(yr[..., :1]), yr], -1)


def anti_interlevel(history, pulse_width=(0.03, 0.003), mult=0.01):
    """Z/internal/train_utils.py:134-172."""
    c = history[-1]['sdist'].detach()
    w = history[-1]['weights'].detach()
    wn = (w / (c[..., 1:] - c[..., :-1])).clamp_max(10)
    total = 0.
    for i, h in enumerate(history[:-1]):
        cp, wp = h['sdist'], h['weights']
        c_, w_ = blur_stepfun(c, wn, pulse_width[i])
        area = 0.5 * (w_[..., 1:] + w_[..., :-1]) * (c_[..., 1:] - c_[..., :-1])
        cdf = torch.cat([torch.zeros_like(area[..., :1]), torch.cumsum(area, -1)], -1)
        ws = torch.diff(_interp_quad_masks(cp, c_, w_, cdf), dim=-1)
        total = total + ((ws - wp).clamp_min(0) ** 2 / (wp + 1e-5)).mean()
    return mult * total


def distortion(history, mult=0.005):
    """Z/internal/stepfun.py:297-307 + train_utils.py:175-181."""
    t, w = history[-1]['sdist'], history[-1]['weights']
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    inter = torch.sum(w * torch.sum(w[..., None, :] * dut, -1), -1)
    intra = torch.sum(w ** 2 * (t[..., 1:] - t[..., :-1]), -1) / 3
    return mult * (inter + intra).mean()


def _edge_aware(rgb, x, eps, channel_sum, mask):
    """train_utils.edge_aware_loss_v2 / edge_aware_loss_for_semantic with `mask` [P,h,w]
    (Z/internal/train_utils.py:329-348,412-431)."""
    x = x / (x.mean(1, True).mean(2, True) + eps)
    gx = torch.abs(x[:, :, :-1] - x[:, :, 1:])
    gy = torch.abs(x[:, :-1] - x[:, 1:])
    if channel_sum:
        gx, gy = gx.sum(-1).unsqueeze(-1), gy.sum(-1).unsqueeze(-1)
    mx = mask[:, :, :-1] * mask[:, :, 1:]
    my = mask[:, :-1, :] * mask[:, 1:, :]
    rx = torch.mean(torch.abs(rgb[:, :, :-1] - rgb[:, :, 1:]), 3, keepdim=True)
    ry = torch.mean(torch.abs(rgb[:, :-1] - rgb[:, 1:]), 3, keepdim=True)
    sx = gx[mx > 0] * torch.exp(-rx[mx > 0])
    sy = gy[my > 0] * torch.exp(-ry[my > 0])
    return sx.mean() + sy.mean()


def losses(batch, rend, history, step, num_patch, patch_size=32, end_step=5000, start_step=0,
           use_intensity=True, instance_obj=False, lidar_supervision=True, only_lidar_supervision=False,
           pose_refine=True, regularisers=True) -> Dict[str, torch.Tensor]:
    """Z/train.py:283-455 for use_semantic=True, depth_loss (the statements of the training loop, in order)."""
    final = rend[-1]
    mask = batch['mask'] == 0                      # train.py:287 "only apply loss on mask == 0"
    if instance_obj:
        mask = torch.zeros_like(mask)              # train.py:288-289
    patch = batch['patch_mask'] == 1
    lidar = batch['lidar_mask'] == 1
    rgb_mask = (mask == 0) & ~patch                # train.py:307
    depth_mask = (batch['depth'] > 0) & rgb_mask
    sem_mask = (batch['semantic'] != 255) & rgb_mask
    if lidar_supervision:                          # train.py:313-319
        rgb_mask = rgb_mask & ~lidar
        depth_mask = depth_mask | lidar
        sem_mask = sem_mask & ~lidar
        if only_lidar_supervision:
            depth_mask = depth_mask & lidar
    refine = pose_refine and start_step < step < int(0.6 * end_step)
    out = {}
    lm = rgb_mask[:, None].float().expand(-1, 3)
    resid = (final['rgb'] - batch['rgb'][..., :3]) ** 2
    out['data'] = (lm * torch.sqrt(resid + 0.001 ** 2)).sum() / lm.sum()
    dep_lam = 0. if refine else (0.4 if step > end_step else 0.1)
    d = final['depth'][depth_mask] - batch['depth'][depth_mask]
    thre = torch.quantile(torch.abs(d), 0.9)
    out['depth'] = dep_lam * torch.log(torch.abs(d[d < thre]) + 1).mean()
    if num_patch > 0:
        shape = (num_patch, patch_size, patch_size)
        mask_patch = torch.where(mask[patch].reshape(*shape) > 0, 0, 1)   # train.py:363-364
        dep = final['depth'][patch].reshape(*shape, -1)
        rgbp = batch['rgb'][patch].reshape(*shape, -1)
        out['d_smo'] = torch.nan_to_num(0.01 * _edge_aware(rgbp, dep, 1e-7, False, mask_patch))
        semp = final['semantic'][patch].reshape(*shape, -1)
        out['s_smo'] = torch.nan_to_num(0.01 * _edge_aware(rgbp, semp, 1e-5, True, mask_patch))
    sem_lam = 0. if refine else (0.04 if step > end_step else 0.01)
    if sem_mask.sum() > 0:
        out['sem'] = sem_lam * nn.NLLLoss()(torch.log(final['semantic'][sem_mask] + 1e-6), batch['semantic'][sem_mask].long())
    else:
        out['sem'] = torch.tensor(0.) * sem_lam
    if use_intensity:
        out['int'] = 0.1 * (final['intensity'].reshape(-1) - batch['intensity'].reshape(-1))[lidar].pow(2).mean()
    if regularisers:
        out['interlevel'] = anti_interlevel(history)
        out['distortion'] = distortion(history)
    if 'hash_decay' in final:
        out['hash_decay'] = final['hash_decay']
    return out


def learning_rate(step, lr_init=0.01, lr_final=0.001, max_steps=25000, delay_steps=5000, delay_mult=1e-8):
    """Z/internal/math.py:54-86."""
    delay = delay_mult + (1 - delay_mult) * math.sin(0.5 * math.pi * min(max(step / delay_steps, 0), 1))
    t = min(max(step / max_steps, 0), 1)
    return delay * math.exp(t * (math.log(lr_final) - math.log(lr_init)) + math.log(lr_init))


class RefTrainer:
    """Parameters as leaf tensors + torch.optim.Adam(betas=.9/.99, eps=1e-15)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], max_steps=25000):
        self.p = {}
        for k, v in state_dict.items():
            self.p[k] = v.clone().requires_grad_(True) if v.is_floating_point() else v
        self.params = [v for v in self.p.values() if v.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=0.01, betas=(0.9, 0.99), eps=1e-15)
        self.max_steps = max_steps

    def step(self, batch, rand_inputs, step: int, num_patch: int):
        for g in self.opt.param_groups:
            g['lr'] = learning_rate(step, max_steps=self.max_steps)
        self.opt.zero_grad()
        train_frac = float(np.clip((step - 1) / (self.max_steps - 1), 0, 1))
        rend, hist = zo.model_forward(self.p, batch, rand_inputs, train_frac, True, training=True)
        ls = losses(batch, rend, hist, step, num_patch)
        loss = sum(ls.values())
        loss.backward()
        for q in self.params:
            if q.grad is not None:
                q.grad.nan_to_num_()
        self.opt.step()
        return {k: float(v) for k, v in ls.items()}, float(loss)
