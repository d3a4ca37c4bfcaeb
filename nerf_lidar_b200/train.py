"""One optimisation step of the zipnerf nuScenes run around the hot path
(reference: Z/train.py:174-470, Z/internal/train_utils.py:55-275).

The forward/backward of the path itself runs on the libnlb200 kernels through
`Model.forward`.  The loss terms are SURVEY.md section 8(f) "next #2" (training-step
remainder): fused kernels (csrc/render_losses.cu, csrc/losses.cu), free of host
synchronisation; the dense per-table passes (hash-decay gradient, NaN scrub, Adam,
zero_grad) are ONE fused kernel per table (csrc/adam.cu).

Data parallelism: every rank runs the step on its own rays; table and
MLP gradients are summed with one all-reduce per buffer before the optimizer pass
(`grad_scale = 1/world` reproduces DDP's mean).  The hash-decay gradient is a pure
function of the replicated parameters and is applied after the reduction."""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, parallel

# development switch (A/B timing): the loss dictionary summed and differentiated by torch scalar kernels
_FUSED_LOSSES = os.environ.get('NLB_LOSS_UNFUSED', '') in ('', '0')
from .configs import Config
from .models import Model


# ----------------------------------------------------------------------------- schedules
def log_lerp(t, v0, v1):
    lv0, lv1 = math.log(v0), math.log(v1)
    return math.exp(min(max(t, 0.), 1.) * (lv1 - lv0) + lv0)


def learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1.):
    """Z/internal/math.py:54-86."""
    if lr_delay_steps > 0:
        delay = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0), 1))
    else:
        delay = 1.
    return delay * log_lerp(step / max_steps, lr_init, lr_final)


# ----------------------------------------------------------------------------- losses
def anti_interlevel_loss(ray_history, config: Config):
    """train_utils.py:134-172 on the fused kernel (csrc/losses.cu)."""
    from . import ops
    last = ray_history[-1]
    total = 0.
    for i, rr in enumerate(ray_history[:-1]):
        per_ray = ops.interlevel_per_ray(last['sdist'], last['weights'], rr['sdist'], rr['weights'],
                                         config.pulse_width[i])
        total = total + per_ray.sum() / (rr['weights'].shape[0] * rr['weights'].shape[1])
    return config.anti_interlevel_loss_mult * total


def distortion_loss(ray_history, config: Config):
    """stepfun.lossfun_distortion (Z/internal/stepfun.py:297-307) on the final level,
    fused kernel (csrc/losses.cu)."""
    from . import ops
    t, w = ray_history[-1]['sdist'], ray_history[-1]['weights']
    return config.distortion_loss_mult * ops.distortion_per_ray(t, w).mean()


def _supervision_setup(renderings, config: Config, step: int, num_patch: int):
    """(rendering inputs, kernel configuration, loss name -> index into [data, depth, sem, int, d_smo, s_smo]) of
    the supervision terms of Z/train.py:283-455; the step-dependent multipliers follow Z/train.py:330-371."""
    final = renderings[-1]
    refine = config.pose_refine and config.start_step < step < int(0.6 * config.end_step)
    dep_lam = 0. if refine else (0.4 if step > config.end_step else 0.1)
    sem_lam = 0. if refine else (0.04 if step > config.end_step else 0.01)
    use_sem = bool(config.use_semantic) and 'semantic' in final
    use_int = bool(config.use_intensity) and 'intensity' in final
    rend = dict(rgb=final['rgb'], depth=final['depth'], semantic=final['semantic'] if use_sem else None,
                intensity=final['intensity'] if use_int else None)
    cfg = dict(num_patch=num_patch, patch_size=config.patch_size, lidar_supervision=config.lidar_supervision,
               only_lidar_supervision=config.only_lidar_supervison, charb=config.data_loss_type != 'mse',
               charb_padding=config.charb_padding, depth_mult=dep_lam if config.depth_loss else 0.,
               sem_mult=sem_lam, int_mult=0.1, smooth_mult=0.01,
               instance_obj=bool(getattr(config, 'instance_obj', False)))
    index = {'data': 0}
    if config.depth_loss:
        index['depth'] = 1
    if num_patch > 0:
        index['d_smo'] = 4
        if use_sem:
            index['s_smo'] = 5
    if use_sem:
        index['sem'] = 2
    if use_int:
        index['int'] = 3
    return rend, cfg, index


def compute_losses(batch: Dict[str, torch.Tensor], renderings, ray_history, config: Config, step: int,
                   num_patch: int) -> Dict[str, torch.Tensor]:
    """The loss dictionary of Z/train.py:283-455 for the nuScenes camera+LiDAR run on the fused
    kernels: supervision terms in csrc/render_losses.cu (incl. the dataset mask `batch['mask']`,
    Z/train.py:286-327), regularisers in csrc/losses.cu.  Every entry carries its own autograd graph (the
    trainer uses `compute_losses_fused`, which forms the sums and seeds the backward pass in three launches)."""
    from . import ops
    rend, cfg, index = _supervision_setup(renderings, config, step, num_patch)
    vals = ops.render_losses(rend, batch, cfg)
    losses = {k: vals[i] for k, i in index.items()}
    if config.anti_interlevel_loss_mult > 0:
        losses['interlevel'] = anti_interlevel_loss(ray_history, config)
    if config.distortion_loss_mult > 0:
        losses['distortion'] = distortion_loss(ray_history, config)
    return losses


def compute_losses_fused(batch: Dict[str, torch.Tensor], renderings, ray_history, config: Config, step: int,
                         num_patch: int, extra_values: Optional[Dict[str, torch.Tensor]] = None):
    """The same dictionary as `compute_losses` with its sums formed on the device in two launches
    (ops.interlevel_total, ops.main_loss): returns (values, main, prop) -- `values` the detached entries plus
    'hash_decay' (when the forward reported it) and 'loss' = the step's total, `main` / `prop` the two scalars to
    back-propagate (prop is None without an anti-interlevel loss).  `extra_values`: reported entries without a
    gradient that count into the total (latent_reg)."""
    from . import ops
    rend, cfg, index = _supervision_setup(renderings, config, step, num_patch)
    last = ray_history[-1]
    prop = None
    if config.anti_interlevel_loss_mult > 0 and len(ray_history) > 1:
        prop = ops.interlevel_total(last['sdist'], last['weights'], [rr['sdist'] for rr in ray_history[:-1]],
                                    [rr['weights'] for rr in ray_history[:-1]], config.pulse_width,
                                    config.anti_interlevel_loss_mult)
    extra_values = extra_values or {}
    used = [i in index.values() for i in range(6)]
    main, vals, extras = ops.main_loss(rend, last['weights'], last['sdist'], batch, cfg, used,
                                       max(float(config.distortion_loss_mult), 0.), prop,
                                       renderings[-1].get('hash_decay'), list(extra_values.values()))
    values = {k: vals[i] for k, i in index.items()}
    if prop is not None:
        values['interlevel'] = prop.detach()
    if config.distortion_loss_mult > 0:
        values['distortion'] = extras[0]
    for k, v in extra_values.items():
        values[k] = v
    if 'hash_decay' in renderings[-1]:
        values['hash_decay'] = extras[2]
    values['loss'] = extras[1]
    return values, main, prop


# ----------------------------------------------------------------------------- trainer
class Trainer:
    """Owns the optimizer state and runs train steps.  Tables keep persistent
    gradient buffers (`param._nlb_grad`) that the scatter kernels add into and the
    fused Adam pass clears; all dense-layer parameters live in ONE flat buffer so
    their Adam update is one launch and their all-reduce one message."""

    def __init__(self, model: Model, config: Config, world: int = 1, rank: int = 0):
        self.model, self.config, self.world, self.rank = model, config, world, rank
        self.tables = []
        dense = []

        def arena_for(n, dev):
            """Padded flat buffers of a parameter group: parameters and gradients whole (chunk * world floats,
            the collectives exchange equal 16-byte-aligned chunks), Adam moments for this rank's chunk only."""
            chunk = parallel.shard_chunk(n, world)
            lo = min(rank * chunk, n)
            cnt = min(chunk, n - lo)
            z = lambda k: torch.zeros(k, device=dev)
            return dict(n=n, chunk=chunk, lo=lo, cnt=cnt, param=z(chunk * world), grad=z(chunk * world), m=z(chunk), v=z(chunk))

        for name, p in model.named_parameters():
            if name.endswith('encoder.embeddings'):
                enc = model.get_submodule(name.rsplit('.', 2)[0]).encoder
                ar = arena_for(p.numel(), p.device)
                ar['param'][:ar['n']].copy_(p.data.reshape(-1))
                p.data = ar['param'][:ar['n']].view_as(p)
                p._nlb_grad = ar['grad'][:ar['n']].view_as(p)
                offs = enc.offsets.tolist()
                counts = torch.tensor([(offs[i + 1] - offs[i]) * enc.level_dim for i in range(enc.num_levels)],
                                      dtype=torch.float32, device=p.device)
                self.tables.append(dict(name=name, param=p, enc=enc, grad=p._nlb_grad, arena=ar, m=ar['m'], v=ar['v'],
                                        sumsq=torch.zeros(enc.num_levels, device=p.device), counts=counts,
                                        offsets=(C.c_int32 * (enc.num_levels + 1))(*offs)))
            else:
                dense.append(p)
        # largest table first: it is the first gradient to be final in the backward pass
        self.tables.sort(key=lambda t: -t['param'].numel())
        n = sum((p.numel() + 3) // 4 * 4 for p in dense)
        dev = dense[0].device
        self.flat_arena = arena_for(n, dev)
        self.flat = self.flat_arena['param'][:n]
        self.flat_grad = self.flat_arena['grad'][:n]
        self.flat_m, self.flat_v = self.flat_arena['m'], self.flat_arena['v']   # this rank's chunk
        off = 0
        for p in dense:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
            p._nlb_grad = p.grad   # the NerfMLP weight-gradient kernels add into it directly (ops._NerfMLP)
            off += (k + 3) // 4 * 4
        self.dense = dense
        self.decay = config.hash_decay_mults if config.hash_decay_mults > 0 else 0.
        self.hash_decay_value = torch.zeros((), device=dev)  # persistent: written in place (CUDA-graph safe)
        self._one = torch.ones((), device=dev)
        self._hd_coef = None
        # per-level sums of squares of all tables in one buffer: ONE small all-reduce per step in data-parallel runs
        self._sumsq_all = torch.zeros(sum(t['enc'].num_levels for t in self.tables), device=dev)
        o = 0
        for t in self.tables:
            t['sumsq'] = self._sumsq_all[o:o + t['enc'].num_levels]
            o += t['enc'].num_levels
        self._graphs = {}
        self._static = None
        # Every training step -- eager, warm-up or captured -- is issued on this stream.  Autograd binds a
        # leaf's AccumulateGrad node to the stream of its first use (the PropMLP weight gradients still travel
        # through autograd); a step first run eagerly on the legacy default stream and later captured on
        # another one made the engine join the default stream into the capture
        # (cudaErrorStreamCaptureIsolation: tests/dp_worker.py, eager step followed by the graphed one).
        self.stream = torch.cuda.Stream(device=dev)
        # data parallel: collectives still in flight when a step returns (the all-gather of the NeRF table's
        # updated rows travels under the next step's proposal levels); Model.forward and every trainer entry
        # point wait for them first
        self._pending = []
        self.posenet = self.pn_optimizer = self.pn_lr_fn = None
        self.tracknet = self.tn_optimizer = self.tn_lr_fn = None
        model.__dict__['_nlb_sync'] = self.sync
        model.train()
        model.training = True

    def lr(self, step: int) -> float:
        c = self.config
        return learning_rate_decay(step, c.lr_init, c.lr_final, c.max_steps, c.lr_delay_steps, c.lr_delay_mult)

    # The proposal networks receive gradients only from the anti-interlevel loss (the
    # sampling weights are detached: Model.stop_level_grad) and that loss sees the final
    # level detached, so the backward pass splits into two independent halves:
    #   main -> NeRF table + NerfMLP        prop -> proposal tables + PropMLPs.
    # In data-parallel runs the NeRF table's all-reduce (77 % of the gradient bytes) is
    # issued between the two and travels while the proposal half is still computing.
    PROP_LOSSES = ('interlevel',)

    def forward_losses(self, batch, step: int, num_patch: Optional[int] = None, rand_inputs=None, curr_track=None):
        """Forward pass and loss dictionary; returns (losses, loss_main, loss_prop)."""
        c = self.config
        train_frac = float(np.clip((step - 1) / (c.max_steps - 1), 0, 1))
        if num_patch is None:
            num_patch = (c.batch_size_per_rank // 4) // (c.patch_size ** 2) if hasattr(c, 'batch_size_per_rank') else 0
        renderings, ray_history = self.model(True, batch, train_frac, True, zero_glo=False, sample_n=c.sample_n_train,
                                             sample_m=c.sample_m_train, step=step, max_step=c.max_steps,
                                             curr_track=curr_track, rand_inputs=rand_inputs)
        latents = getattr(self.model, 'latent_vector_dict', None)
        extra = {}
        if getattr(c, 'latent_size', 0) > 0 and latents is not None:
            # Z/train.py:394-399 with train_utils.latentReg (Z/internal/train_utils.py:456-457): the reference
            # rebuilds the sum with torch.tensor([...]), which drops the graph -- a reported value, no gradient
            with torch.no_grad():
                extra['latent_reg'] = sum(c.latent_reg * torch.norm(v) for v in latents.values())
        if not _FUSED_LOSSES:
            return self._forward_losses_unfused(batch, renderings, ray_history, step, num_patch, extra)
        # the sums of the dictionary, the step's total and (in the backward pass) every gradient seed come from
        # three launches; the returned values are detached views of their outputs
        return compute_losses_fused(batch, renderings, ray_history, c, step, num_patch, extra)

    def _forward_losses_unfused(self, batch, renderings, ray_history, step, num_patch, extra):
        """The same step with the dictionary summed by torch (A/B timing: NLB_LOSS_UNFUSED=1)."""
        losses = compute_losses(batch, renderings, ray_history, self.config, step, num_patch)
        losses.update(extra)
        main = sum(v for k, v in losses.items() if k not in self.PROP_LOSSES)
        prop = [v for k, v in losses.items() if k in self.PROP_LOSSES]
        prop = sum(prop) if prop else None
        # values only from here on: the returned dictionary must not keep the autograd graph (and with it the
        # AccumulateGrad nodes, which are bound to the stream of their first use) alive across steps / captures
        losses = {k: v.detach() for k, v in losses.items()}
        total = main.detach() if prop is None else main.detach() + prop.detach()
        if 'hash_decay' in renderings[-1]:
            # value only (its gradient is fused into Adam); cloned because the optimizer pass
            # overwrites the persistent buffer with the updated tables' value
            losses['hash_decay'] = renderings[-1]['hash_decay'].clone()
            total = total + losses['hash_decay']
        losses['loss'] = total
        return losses, main, prop

    def _nerf_tables(self):
        return [t for t in self.tables if 'prop' not in t['name']]

    def _prop_tables(self):
        return [t for t in self.tables if 'prop' in t['name']]

    # ------------------------------------------------------------------------- pose refinement (Z/train.py:200-240,464-466)
    def attach_posenet(self, net, optimizer, lr_fn):
        """`posenet.create_posenet(...)`'s triple.  Inside the window (start_step < step < end_step) every step
        refines the batch's rays with gradients, back-propagates into the corrections through the hot path's
        ray-geometry gradients and steps `optimizer`; afterwards the corrections are applied without gradients."""
        self.posenet, self.pn_optimizer, self.pn_lr_fn = net, optimizer, lr_fn

    def attach_tracknet(self, net, optimizer, lr_fn):
        """`posenet.create_tracknet(...)`'s triple (Z/train.py:99-103,244-266,467-471): inside
        track_start_opt < step < track_start_opt + 5000 the object tracks are refined -- `Model.forward` gets
        `curr_track` with a graph, the object branch returns the gradient of the interpolated poses -- afterwards the
        refined tracks are applied without gradients."""
        self.tracknet, self.tn_optimizer, self.tn_lr_fn = net, optimizer, lr_fn

    def _track_mode(self, step: int):
        if getattr(self, 'tracknet', None) is None:
            return None
        from . import posenet as pn
        return pn.track_window(self.config, step)

    def _current_track(self, mode):
        from . import posenet as pn
        if mode == 'train':
            self.tn_optimizer.zero_grad(set_to_none=True)
            return pn.refined_track(self.tracknet, self.flat.device)
        if mode == 'apply':
            with torch.no_grad():
                return pn.refined_track(self.tracknet, self.flat.device)
        return None

    def _side_step(self, net, optimizer):
        """Gradient mean over the ranks (the reference wraps the side networks in DDP), clipping, Adam."""
        params = [p for p in net.parameters() if p.grad is not None]
        if self.world > 1:
            for p in params:
                torch.distributed.all_reduce(p.grad)
                p.grad.div_(self.world)
        c = self.config
        if c.grad_max_val > 0:  # train_utils.clip_gradients (Z/internal/train_utils.py:223-232)
            torch.nn.utils.clip_grad_value_(params, c.grad_max_val)
        if c.grad_max_norm > 0:
            torch.nn.utils.clip_grad_norm_(params, c.grad_max_norm)
        optimizer.step()

    def _pose_mode(self, step: int):
        if getattr(self, 'posenet', None) is None:
            return None
        from . import posenet as pn
        return pn.pose_window(self.config, step)

    def _set_pose_lr(self, step: int):
        """Outside any capture: the learning rate is a device scalar the captured Adam reads."""
        for mode, opt, fn in ((self._pose_mode(step), self.pn_optimizer, self.pn_lr_fn),
                              (self._track_mode(step), self.tn_optimizer, self.tn_lr_fn)):
            if mode != 'train':
                continue
            lr = float(fn(step))
            for group in opt.param_groups:
                if torch.is_tensor(group['lr']):
                    group['lr'].fill_(lr)
                else:
                    group['lr'] = lr

    def _refined(self, batch, mode):
        from . import posenet as pn
        if mode == 'train':
            self.pn_optimizer.zero_grad(set_to_none=True)
            return pn.refine_rays(batch, self.posenet)
        if mode == 'apply':
            with torch.no_grad():
                return pn.refine_rays(batch, self.posenet)
        return batch

    def _pose_step(self, mode, track_mode=None):
        if mode == 'train':
            self._side_step(self.posenet, self.pn_optimizer)
        if track_mode == 'train':
            self._side_step(self.tracknet, self.tn_optimizer)

    def sync(self):
        """Waits (on the current stream) for the collectives a data-parallel step left in flight: afterwards the
        replicated parameters are complete on this rank."""
        if self._pending:
            parallel.wait_all(self._pending)
            self._pending = []

    def train_step(self, batch: Dict[str, torch.Tensor], step: int, num_patch: Optional[int] = None,
                   rand_inputs=None) -> Dict[str, torch.Tensor]:
        """One eager training step, issued on the trainer's own stream (the caller's stream waits for it)."""
        self.sync()
        dev = self.flat.device
        cur = torch.cuda.current_stream(dev)
        if cur == self.stream or torch.cuda.is_current_stream_capturing():
            if not torch.cuda.is_current_stream_capturing():
                self._set_pose_lr(step)
            return self._train_step(batch, step, num_patch, rand_inputs)
        self._set_pose_lr(step)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            out = self._train_step(batch, step, num_patch, rand_inputs)
        cur.wait_stream(self.stream)
        return out

    def _backward(self, *losses, retain_graph: bool = False):
        """Back-propagates the given scalars (None entries are skipped) from a cached unit seed: `loss.backward()`
        fills a fresh ones tensor per root and `main + prop` is one more launch."""
        roots = [l for l in losses if l is not None]
        torch.autograd.backward(roots, [self._one] * len(roots), retain_graph=retain_graph)

    def _train_step(self, batch, step, num_patch=None, rand_inputs=None):
        pose, tmode = self._pose_mode(step), self._track_mode(step)
        batch = self._refined(batch, pose)
        losses, main, prop = self.forward_losses(batch, step, num_patch, rand_inputs, self._current_track(tmode))
        if self.world == 1:
            # (running the NeRF table's optimizer pass on a side stream beside the proposal backward was
            # measured: 7.30-7.35 ms per step against 7.23 ms in order -- the two contend for L2)
            self._backward(main, prop)
            self.optimizer_step(step)
            self._pose_step(pose, tmode)
            return losses
        # the refined rays feed both halves: the first backward must keep the posenet's part of the graph
        self._backward(main, retain_graph=pose == 'train' and prop is not None)
        early = self.reduce_scatter_gradients(self._nerf_tables())
        if prop is not None:
            self._backward(prop)
        late = self.reduce_scatter_gradients(self._prop_tables(), dense=True)
        parallel.wait_all(early + late)
        self.optimizer_step(step, reduce=False)
        self._pose_step(pose, tmode)
        return losses

    # ---- data parallel: reduce-scatter of the gradients -> this rank's slice of the optimizer pass -> all-gather
    def _arenas(self, tables, dense=False):
        return [t['arena'] for t in tables] + ([self.flat_arena] if dense else [])

    def reduce_scatter_gradients(self, tables, dense=False):
        """Starts the gradient sums (what DDP's all-reduce does at Z/train.py:459, first half) and returns the
        handles.  Afterwards only this rank's chunk of each gradient buffer holds the sum."""
        h = []
        for ar in self._arenas(tables, dense):
            h += parallel.reduce_scatter_async(ar['grad'], ar['chunk'], self.rank)
        return h

    def all_gather_parameters(self, tables, dense=False):
        h = []
        for ar in self._arenas(tables, dense):
            h += parallel.all_gather_async(ar['param'], ar['chunk'], self.rank)
        return h

    def optimizer_tables(self, step: int, tables):
        """Fused hash-decay + NaN scrub + Adam + zero-grad pass over this rank's slice of the given tables."""
        c = self.config
        lr = self.lr(step)
        scale = 1.0 / self.world
        lib = _lib.load()
        st = _lib.stream()
        if len(tables) == len(self.tables):
            self._sumsq_all.zero_()          # (views of one buffer: a single fill)
        for t in tables:
            enc, ar = t['enc'], t['arena']
            decay = 0. if (c.obj_nodecay and 'obj' in t['name']) else self.decay
            t['decay'] = decay
            if len(tables) != len(self.tables):
                t['sumsq'].zero_()
            if self.world > 1:
                # the chunks of the other ranks hold their own partial sums: clear them for the next step's scatter
                g = ar['grad']
                g[:ar['lo']].zero_()
                g[ar['lo'] + ar['cnt']:].zero_()
            with _lib.timed('adam_table'):
                _lib.check(lib.nlb_adam_table_step_range(
                    ar['param'].data_ptr(), ar['grad'].data_ptr(), ar['m'].data_ptr(), ar['v'].data_ptr(), t['offsets'],
                    enc.num_levels, enc.level_dim, float(decay), float(lr), c.adam_beta1, c.adam_beta2, c.adam_eps,
                    int(step), scale, t['sumsq'].data_ptr(), ar['lo'], ar['cnt'], st))

    def optimizer_finish(self, step: int):
        """Dense-layer Adam pass (this rank's slice), the all-gather of the updated parameters and the
        bookkeeping after all table passes."""
        c = self.config
        lib = _lib.load()
        fa = self.flat_arena
        if self.world > 1:
            g = fa['grad']
            g[:fa['lo']].zero_()
            g[fa['lo'] + fa['cnt']:].zero_()
        if fa['cnt'] > 0:
            _lib.check(lib.nlb_adam_step(fa['param'].data_ptr() + 4 * fa['lo'], fa['grad'].data_ptr() + 4 * fa['lo'],
                                         fa['m'].data_ptr(), fa['v'].data_ptr(), fa['cnt'], float(self.lr(step)),
                                         c.adam_beta1, c.adam_beta2, c.adam_eps, int(step), 1.0 / self.world,
                                         _lib.stream()))
        self._mark_packed_stale()

    def publish(self, capture_safe_only: bool = False, defer_nerf: bool = False):
        """After the optimizer pass of a data-parallel step: all-gather of the updated parameters (and of the
        per-level sums of squares behind the reported hash-decay value).  With `defer_nerf` the NeRF table's
        gather (77 % of the bytes) is issued last and left in flight: the next step reads that table only after its
        two proposal levels, so the transfer hides under them (`sync()` waits for it)."""
        if self.world > 1:
            first = self._prop_tables() if defer_nerf else self.tables
            h = self.all_gather_parameters(first, dense=True)
            if parallel.is_dist():
                import torch.distributed as dist
                h.append(dist.all_reduce(self._sumsq_all, async_op=True))
            parallel.wait_all(h)
            if defer_nerf:
                self._pending += self.all_gather_parameters(self._nerf_tables())
        self._update_hash_decay_value()

    def _update_hash_decay_value(self):
        # Model.hash_decay_loss of the UPDATED tables (mean over levels of the per-level mean square): the next
        # forward reports it as renderings[-1]['hash_decay'] without re-reading 310 MB of tables.  One launch: the
        # per-level sums of squares (all tables, one buffer) against decay / (levels x values per level)
        if self.decay <= 0 or not any(t.get('decay', 0) > 0 for t in self.tables):
            return
        if self._hd_coef is None:
            coef = []
            for t in self.tables:
                L = t['enc'].num_levels
                on = 1.0 if t.get('decay', 0) > 0 else 0.0
                coef.append(on * self.decay / (L * t['counts']))
            self._hd_coef = torch.cat(coef).contiguous()
        from . import ops
        ops.weighted_sums([(self._sumsq_all, self._hd_coef, 1.0, 0)], self.hash_decay_value)
        self.model._hash_decay_value = self.hash_decay_value

    def _mark_packed_stale(self):
        """The dense parameters changed through raw pointers (torch's version counters do not see it): the
        packed tensor-core operand images are stale.  Called by the optimizer pass and after every REPLAY of a
        captured step -- a replay runs the optimizer kernels but none of this Python."""
        for m in self.model.modules():
            if hasattr(m, '_nlb_dirty'):
                m._nlb_dirty = True

    def optimizer_step(self, step: int, reduce: bool = True, publish: bool = True):
        if reduce and self.world > 1:
            parallel.wait_all(self.reduce_scatter_gradients(self.tables, dense=True))
        self.optimizer_tables(step, self.tables)
        self.optimizer_finish(step)
        if publish:
            self.publish()

    # ------------------------------------------------------------------------- CUDA-graph step
    def _regime(self, step: int):
        """Everything a captured step bakes in besides the dynamic scalars: the
        step-dependent loss multipliers of Z/train.py:330-371 change only at the
        pose-refinement window boundaries."""
        c = self.config
        refine = c.pose_refine and c.start_step < step < int(0.6 * c.end_step)
        return (bool(refine), step > c.end_step, self._pose_mode(step), self._track_mode(step))

    def _write_dynamic(self, step: int):
        c = self.config
        train_frac = float(np.clip((step - 1) / (c.max_steps - 1), 0, 1))
        slope = self.model.anneal_slope
        anneal = (slope * train_frac) / ((slope - 1) * train_frac + 1) if slope > 0 else 1.
        out2 = (C.c_float * 2)()
        _lib.check(_lib.load().nlb_adam_bias_terms(float(self.lr(step)), c.adam_beta1, c.adam_beta2, int(step), out2))
        # fill_ passes the value as a kernel argument at enqueue time.  (An asynchronous copy from ONE pinned
        # staging buffer reads it when the copy executes: the host runs many replayed steps ahead of the device
        # and would overwrite the scalars of a step that has not started yet.)
        for i, v in enumerate((anneal, out2[0], out2[1])):
            self._static['dyn'][i:i + 1].fill_(float(v))

    def train_step_graphed(self, batch: Dict[str, torch.Tensor], step: int, num_patch: Optional[int] = None,
                           rand_inputs=None) -> Dict[str, torch.Tensor]:
        """`train_step` replayed as ONE CUDA graph (single process) or as three graphs --
        forward + main backward, proposal backward, optimizer pass -- around the eagerly
        issued NCCL gradient all-reduces (data parallel; the NeRF table's reduction overlaps
        the proposal backward): the step is ~200 launches and host-bound when issued eagerly.  `batch` (device or pinned-host tensors) is copied
        into static device buffers; the per-step scalars (anneal, learning rate, Adam bias
        corrections) travel through the library's dynamic-scalar buffer
        (nlb_set_dynamic_scalars); the graph is re-captured when the batch layout or the
        loss-multiplier regime changes.  Returns the loss dictionary (static tensors,
        overwritten by the next call)."""
        dev = self.flat.device
        if self.world > 1 and 'train' in (self._pose_mode(step), self._track_mode(step)):
            # the corrections' gradient all-reduce sits between the backward and their Adam step: the
            # data-parallel window (a fifth of the run at most) is issued eagerly
            return self.train_step(batch, step, num_patch, rand_inputs)
        key = (self._regime(step), num_patch, tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(batch.items())),
               rand_inputs is not None)
        if self._static is None:
            self._static = dict(dyn=torch.zeros(4, device=dev))
        st = self._static
        if st.get('batch_key') != key[2]:
            st['batch'] = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in batch.items()}
            st['batch_key'] = key[2]
            st.pop('rand', None)
            self._graphs.clear()
        for k, v in batch.items():
            st['batch'][k].copy_(v, non_blocking=True)
        srand = None
        if rand_inputs is not None:  # injected random draws (parity tests): static copies as well
            if 'rand' not in st:
                st['rand'] = [{k: torch.empty_like(v, device=dev) for k, v in r.items()} for r in rand_inputs]
            for dst, src in zip(st['rand'], rand_inputs):
                for k, v in src.items():
                    dst[k].copy_(v, non_blocking=True)
            srand = st['rand']
        self._write_dynamic(step)
        self._set_pose_lr(step)
        lib = _lib.load()
        entry = self._graphs.get(key)
        if entry is None:
            # warm up eagerly on a side stream (allocator, cuBLAS handles, kernel attributes), then capture
            lib.nlb_set_dynamic_scalars(st['dyn'].data_ptr())
            try:
                out = self.train_step(st['batch'], step, num_patch, srand)   # on self.stream
                out = {k: v.clone() for k, v in out.items()}
                g = torch.cuda.CUDAGraph()
                g_prop = g_opt = g_nerf = None
                if self.world == 1:
                    with torch.cuda.graph(g, stream=self.stream):
                        captured = self._train_step(st['batch'], step, num_patch, srand)
                else:
                    # the collectives stay outside the graphs: proposal levels | NeRF level + main backward |
                    # proposal backward | optimizer.  The first cut sits where the forward first reads the NeRF
                    # table (Model.forward calls the hook): the previous step's all-gather of that table is
                    # waited for between the two replays.
                    g_nerf = torch.cuda.CUDAGraph()
                    ctx = [torch.cuda.graph(g, stream=self.stream)]
                    cut = []

                    def hook():
                        if not cut:
                            ctx.pop().__exit__(None, None, None)
                            ctx.append(torch.cuda.graph(g_nerf, pool=g.pool(), stream=self.stream))
                            ctx[0].__enter__()
                            cut.append(True)
                    self.model.__dict__['_nlb_before_nerf_table'] = hook
                    ctx[0].__enter__()
                    try:
                        captured, main, prop = self.forward_losses(st['batch'], step, num_patch, srand)
                        self._backward(main)
                    finally:
                        self.model.__dict__.pop('_nlb_before_nerf_table', None)
                        ctx.pop().__exit__(None, None, None)
                    if not cut:
                        raise RuntimeError('train_step_graphed: the forward never reached the NeRF level')
                    if prop is not None:
                        g_prop = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g_prop, pool=g.pool(), stream=self.stream):
                            self._backward(prop)
                    g_opt = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_opt, pool=g.pool(), stream=self.stream):
                        self.optimizer_step(step, reduce=False, publish=False)
                    del main, prop
            finally:
                lib.nlb_set_dynamic_scalars(None)
            # capturing records the step without running it: the eager warm-up step above WAS this call's step
            self._graphs[key] = (g, g_nerf if self.world > 1 else None, g_prop, g_opt, captured)
            return out
        g, g_nerf, g_prop, g_opt, captured = entry
        if g_opt is None:
            g.replay()
        else:
            g.replay()                      # proposal levels: prop tables + dense parameters (gathered, waited)
            self.sync()                     # the NeRF table's rows of the previous step have arrived
            g_nerf.replay()                 # NeRF level, losses, main backward
            early = self.reduce_scatter_gradients(self._nerf_tables())
            if g_prop is not None:
                g_prop.replay()
            late = self.reduce_scatter_gradients(self._prop_tables(), dense=True)
            parallel.wait_all(early + late)
            g_opt.replay()
            self.publish(defer_nerf=True)
        self._mark_packed_stale()
        return captured
