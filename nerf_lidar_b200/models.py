"""`Model` / `NerfMLP` / `PropMLP` / `render_image` with the reference's public
interface (Z/internal/models.py:30-177 ctor, :239-576 forward, :1379-1507
render_image) on the B200 kernels.

Scope: the static-scene zipnerf path of nuscenes_single.gin --
3 sampling levels, hash-grid PropMLPs, NerfMLP with semantic + intensity heads,
opaque background, power-transformation ray warp -- and, with Config.instance_obj, the dynamic-object
branch (Z/internal/models.py:92-177 ctor, :306-315,401-477 forward; ObjMLP per class with split shape /
texture latents) on csrc/obj.cu (poses are constants: track refinement is outside).

State-dict keys are the reference's (nerf_mlp.encoder.embeddings,
nerf_mlp.density_layer.0.weight, prop_mlp_0.encoder.offsets, ...), so reference
checkpoints load with `load_state_dict(strict=False)` exactly as
Z/internal/checkpoints.py:26-55 does."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .configs import Config, configurable
from .gridencoder import GridEncoder


def set_kwargs(self, kwargs):
    for k, v in kwargs.items():
        setattr(self, k, v)


class MLP(nn.Module):
    """Parameter container + head evaluation of Z/internal/models.py:796-1263
    (only the options the zipnerf nuScenes path uses)."""
    bottleneck_width: int = 256
    net_depth_viewdirs: int = 2
    net_width_viewdirs: int = 256
    skip_layer_dir: int = 0
    num_rgb_channels: int = 3
    deg_view: int = 4
    density_bias: float = -1.
    rgb_premultiplier: float = 1.
    rgb_bias: float = 0.
    rgb_padding: float = 0.001
    disable_density_normals: bool = False
    disable_rgb: bool = False
    warp_fn = 'contract'
    grid_level_interval: int = 2
    grid_level_dim: int = 4
    grid_base_resolution: int = 16
    grid_disired_resolution: int = 8192
    grid_log2_hashmap_size: int = 21
    class_num: int = 19
    use_semantic: bool = False
    analytic_gradient: bool = True
    use_intensity: bool = False
    no_sem_layer: bool = True
    re_weights: bool = True
    density_init: bool = False
    fixed_semantic: bool = False
    class_type: int = 255
    obj_mode: bool = False
    complex_decoder: bool = False
    latent_size: int = 0
    split_latent: bool = False
    mlp_dtype = torch.bfloat16  # operand type of the dense layers (fp32 accumulation)

    def __init__(self, **kwargs):
        super().__init__()
        set_kwargs(self, kwargs)
        if not self.disable_density_normals:
            raise NotImplementedError('density normals (Ref-NeRF options) are outside the zipnerf hot path; '
                                      'bind disable_density_normals=True as nuscenes_single.gin does')
        if self.use_semantic and self.no_sem_layer and not self.disable_rgb and not self.fixed_semantic:
            raise NotImplementedError('no_sem_layer=True (semantic = bottleneck slice) is not built; '
                                      'nuscenes_single.gin binds Config.no_sem_layer=False')
        if self.obj_mode or self.complex_decoder:
            raise NotImplementedError('obj_mode / complex_decoder density heads are not built '
                                      '(nuscenes_single.gin binds ObjMLP.obj_mode=False)')
        self.grid_num_levels = int(np.log(self.grid_disired_resolution / self.grid_base_resolution)
                                   / np.log(self.grid_level_interval)) + 1
        self.encoder = GridEncoder(input_dim=3, num_levels=self.grid_num_levels, level_dim=self.grid_level_dim,
                                   base_resolution=self.grid_base_resolution,
                                   desired_resolution=self.grid_disired_resolution,
                                   log2_hashmap_size=self.grid_log2_hashmap_size, gridtype='hash',
                                   align_corners=False)
        width = self.encoder.output_dim
        # latent codes of the per-class object networks (Z/internal/models.py:907-912,953-955): with
        # split_latent the first half (shape) joins the grid features, the second (texture) the view branch
        if self.latent_size > 0:
            width += self.latent_size // 2 if self.split_latent else self.latent_size
        self.density_layer = nn.Sequential(nn.Linear(width, 64), nn.ReLU(),
                                           nn.Linear(64, 1 if self.disable_rgb else self.bottleneck_width))
        if self.density_init:
            self.density_layer[2].bias.data[0] = self.density_layer[2].bias.data[0] + 0.1
        if not self.disable_rgb:
            dim_dir_enc = 3 + 2 * 3 * self.deg_view
            in_rgb = self.bottleneck_width + dim_dir_enc
            if self.split_latent:
                in_rgb += self.latent_size // 2
            last = in_rgb
            for i in range(self.net_depth_viewdirs):
                lin = nn.Linear(last, self.net_width_viewdirs)
                torch.nn.init.kaiming_uniform_(lin.weight)
                self.register_module(f'lin_second_stage_{i}', lin)
                last = self.net_width_viewdirs
                if i == self.skip_layer_dir:
                    last += in_rgb
            self.rgb_layer = nn.Linear(last, self.num_rgb_channels)
            if not self.no_sem_layer and not self.fixed_semantic:
                self.sem_layer = nn.Sequential(nn.Linear(self.bottleneck_width, 64), nn.ReLU(),
                                               nn.Linear(64, self.class_num))
            if self.use_intensity:
                self.intensity_layer = nn.Sequential(nn.Linear(self.bottleneck_width, 64), nn.ReLU(),
                                                     nn.Linear(64, 1))

    # -- dense head of the NeRF level ------------------------------------------------
    def dir_enc(self, viewdirs: torch.Tensor) -> torch.Tensor:
        """coord.pos_enc(viewdirs, 0, deg_view, append_identity=True), Z/internal/coord.py:199-210."""
        scales = 2 ** torch.arange(0, self.deg_view, device=viewdirs.device)
        sx = (viewdirs[..., None, :] * scales[:, None]).reshape(*viewdirs.shape[:-1], -1)
        return torch.cat([viewdirs, torch.sin(torch.cat([sx, sx + 0.5 * math.pi], dim=-1))], dim=-1)

    def _check_fused_shapes(self):
        """The fused tcgen05 kernels (csrc/nerf_mlp.cu) are compiled for ONE architecture -- the NerfMLP of
        nuscenes_single.gin -- and take raw pointers: any gin binding that changes a layer shape or one of the
        scalars baked into their epilogues must fail here, not read out of bounds or return wrong numbers."""
        if getattr(self, '_nlb_shapes_ok', False):
            return
        want = {'density_layer.0.weight': (64, 40), 'density_layer.2.weight': (256, 64),
                'sem_layer.0.weight': (64, 256), 'sem_layer.2.weight': (19, 64),
                'intensity_layer.0.weight': (64, 256), 'intensity_layer.2.weight': (1, 64),
                'lin_second_stage_0.weight': (256, 283), 'lin_second_stage_1.weight': (256, 539),
                'rgb_layer.weight': (3, 256)}
        have = {k: tuple(v.shape) for k, v in self.named_parameters()}
        if not self.use_intensity:
            # no intensity head (the configuration of the dynamic-object branch): its rows of the fused
            # sem | intensity layers are packed as zeros and, in training, stay a frozen all-zero head
            want = {k: v for k, v in want.items() if not k.startswith('intensity_layer')}
        bad = [f'{k}: {have.get(k)} (built for {v})' for k, v in want.items() if have.get(k) != v]
        scalars = dict(deg_view=4, net_depth_viewdirs=2, skip_layer_dir=0, density_bias=-1., rgb_premultiplier=1.,
                       rgb_bias=0., rgb_padding=0.001, class_num=19)
        bad += [f'{k}={getattr(self, k)!r} (built for {v!r})' for k, v in scalars.items() if getattr(self, k) != v]
        if not self.use_semantic or self.no_sem_layer:
            bad.append('the semantic head (sem_layer) is required (Config.use_semantic, Config.no_sem_layer=False)')
        if self.mlp_dtype != torch.bfloat16:
            bad.append(f'mlp_dtype={self.mlp_dtype} (the dense layers run with bf16 operands, fp32 accumulation)')
        if bad:
            raise NotImplementedError('NerfMLP: the fused tensor-core kernels are built for the nuscenes_single.gin '
                                      'architecture only; unsupported: ' + '; '.join(bad))
        self._nlb_shapes_ok = True

    def heads(self, feat: torch.Tensor, viewdirs: torch.Tensor, S: int) -> Dict[str, torch.Tensor]:
        """features[N*S, 40] -> density / rgb / semantic / intensity
        (Z/internal/models.py:996-997,1116-1251) on the tcgen05 / TMEM kernels: forward, and in training
        also the data-gradient chain and the weight gradients.  Dense layers run with bf16 operands and fp32
        accumulation (bf16 MLP is a build decision, tolerance 1e-3 vs the fp32
        reference); activations are evaluated in fp32.  There is no other implementation in the product
        (the plain-torch evaluation the parity tests compare with lives in tests/helpers.py)."""
        self._check_fused_shapes()
        if torch.is_grad_enabled():
            return ops.nerf_mlp_train(self, feat, viewdirs, S)
        return ops.nerf_mlp_forward(self, feat, viewdirs, S)

    def forward(self, *a, **k):
        raise NotImplementedError('MLP modules are evaluated through Model.forward (fused kernels); '
                                  'the stand-alone MLP.forward(means, stds) entry is outside the hot-path scope')


@configurable
class NerfMLP(MLP):
    pass


@configurable
class PropMLP(MLP):
    pass


@configurable
class ObjMLP(MLP):
    """Per-class (or per-instance) network of a dynamic object (Z/internal/models.py:1271-1273): evaluated by
    csrc/obj.cu on the sample points that fall inside the object's box."""
    pass


def query_class(class_name: str) -> int:
    """obj_utils.query_class (Z/internal/obj_utils.py:498-508)."""
    if 'human' in class_name:
        return 11
    if 'truck' in class_name or 'trailer' in class_name or 'construction' in class_name:
        return 14
    if 'bus' in class_name:
        return 15
    if 'car' in class_name:
        return 13
    return 255


@configurable
class Model(nn.Module):
    """Z/internal/models.py:30-58 class attributes (gin-configurable)."""
    num_prop_samples = (64, 64)
    num_nerf_samples: int = 32
    num_levels: int = 3
    bg_intensity_range = (1., 1.)
    anneal_slope: float = 10
    stop_level_grad: bool = True
    use_viewdirs: bool = True
    raydist_fn = 'contract'
    single_jitter: bool = True
    dilation_multiplier: float = 0.5
    dilation_bias: float = 0.0025
    num_glo_features: int = 0
    near_anneal_rate = None
    near_anneal_init: float = 0.95
    single_mlp: bool = False
    distinct_prop: bool = True
    resample_padding: float = 0.0
    opaque_background: bool = False
    power_lambda: float = -1.5
    std_scale: float = 0.35
    prop_desired_grid_size = [512, 2048]
    training: bool = False
    # kwargs of the dynamic-object branch (no class attributes: a registered ParameterDict must not be shadowed):
    #   bboxes = (obj_info {track: [T, 9]}, obj_type_info {track: class name})      datasets.py:1457
    #   latent_vector_dict = {'obj_latent_<track>': Parameter[latent_size]}         train_utils.create_latent

    def __init__(self, config: Optional[Config] = None, **kwargs):
        super().__init__()
        set_kwargs(self, kwargs)
        self.config = config if config is not None else Config()
        config = self.config
        if self.raydist_fn != 'power_transformation' or self.single_mlp or not self.distinct_prop \
                or self.num_glo_features > 0 or not self.single_jitter or self.near_anneal_rate is not None \
                or not self.stop_level_grad or self.bg_intensity_range[0] != self.bg_intensity_range[1]:
            raise NotImplementedError('only the nuscenes_single.gin zipnerf configuration is built: '
                                      "raydist_fn='power_transformation', distinct proposal MLPs, no GLO, "
                                      'single_jitter, stop_level_grad, constant background')
        self.nerf_mlp = NerfMLP(use_semantic=config.use_semantic, analytic_gradient=config.analytic_gradient,
                                use_intensity=config.use_intensity, no_sem_layer=config.no_sem_layer)
        for i in range(self.num_levels - 1):
            self.register_module(f'prop_mlp_{i}', PropMLP(grid_disired_resolution=self.prop_desired_grid_size[i]))
        self.instance_obj = False
        self.tracks = None
        if getattr(config, 'instance_obj', False):
            self._init_objects(config)

    def _init_objects(self, config):
        """Z/internal/models.py:92-177: one ObjMLP per class (latent mode: a latent code per track) or per track
        (instance mode), and the track table [n_obj, T, 9]."""
        if getattr(self, 'bboxes', None) is None:
            raise ValueError('Config.instance_obj=True needs Model(bboxes=(obj_info, obj_type_info)) as Z/train.py:88 passes it')
        if config.symmetrize:
            raise NotImplementedError('Config.symmetrize (training-time symmetry loss of the object branch) is not built')
        if config.use_intensity:
            raise NotImplementedError(
                'Config.instance_obj with Config.use_intensity: ObjMLP has no intensity head, and the reference\'s '
                'merge loop (Z/internal/models.py:461-472) cannot overwrite the intensity key with None')
        obj_info, self.obj_type_info = self.bboxes
        latent_mode = not (config.latent_size == 0 and not config.fuse_render)
        self.mlp_type = ['latent' if latent_mode else 'instance' for _ in obj_info]
        tracks = []
        for track_id, bbox_infos in obj_info.items():
            class_type = self.obj_type_info[track_id]
            class_id = query_class(class_type)
            obj_mlp = ObjMLP(deg_view=2, grid_level_interval=2, grid_base_resolution=16, warp_fn=None,
                             re_weights=False, fixed_semantic=True, use_semantic=config.use_semantic,
                             class_type=class_id, latent_size=config.latent_size if latent_mode else 0,
                             **({'grid_level_dim': 2} if latent_mode else {}))
            if self.mlp_type[track_id] == 'instance':
                self.register_module(f'obj_mlp_{track_id}', obj_mlp)
            self.register_module(f'obj_mlp_{class_id}_fusion' if 'fusion' in class_type else f'obj_mlp_{class_id}', obj_mlp)
            tracks.append(np.asarray(bbox_infos))
        self.tracks = torch.from_numpy(np.stack(tracks)).float()          # init_tracks (models.py:180-183)
        self.instance_obj = True
        latents = self.__dict__.pop('latent_vector_dict', None)
        if latent_mode and latents is None:
            raise ValueError('latent mode (Config.latent_size > 0) needs Model(latent_vector_dict=...) as Z/train.py:82-88 builds it')
        self.latent_vector_dict = nn.ParameterDict(latents) if latents is not None else None

    def _obj_network(self, track_id: int):
        """(ObjMLP, latent or None) of a track (Z/internal/models.py:425-437)."""
        if self.mlp_type[track_id] == 'instance':
            return self.get_submodule(f'obj_mlp_{track_id}'), None
        class_type = self.obj_type_info[track_id]
        class_id = query_class(class_type)
        name = f'obj_mlp_{class_id}_fusion' if 'fusion' in class_type else f'obj_mlp_{class_id}'
        return self.get_submodule(name), self.latent_vector_dict[f'obj_latent_{track_id}']

    # Z/internal/models.py:203-223 (value only: its gradient is applied analytically
    # inside the fused optimizer pass, see train.py / csrc/adam.cu)
    @torch.no_grad()
    def hash_decay_loss(self) -> torch.Tensor:
        total = 0.
        for name, param in sorted(self.named_parameters(), key=lambda x: x[0]):
            if 'encoder' in name:
                enc = self.get_submodule(name.split('.')[0]).encoder
                offs = enc.offsets.tolist()
                sq = param.detach() ** 2
                per = torch.stack([sq[offs[l]:offs[l + 1]].mean(0) for l in range(enc.num_levels)])
                total = total + per.mean()
        return self.config.hash_decay_mults * total

    def forward(self, rand, batch, train_frac, compute_extras, zero_glo=True, sample_n=7, sample_m=3, step=0,
                max_step=25000, curr_track=None, rand_inputs: Optional[Sequence[Dict[str, torch.Tensor]]] = None):
        """Same contract as the reference: returns (renderings, ray_history).
        `rand_inputs` (extension) injects the random draws -- per level a dict with
        'jitter'[N,1] and 'deg'[N,S,7] in U[0,1) -- instead of torch.rand, so that
        parity tests see the reference's jitter."""
        if sample_n != 7 or sample_m != 3:
            raise NotImplementedError('the fused kernels implement the n=7, m=3 hexagonal multisample')
        sync = self.__dict__.get('_nlb_sync')     # a data-parallel trainer may still be gathering the NeRF table
        if sync is not None and not (batch['origins'].is_cuda and torch.cuda.is_current_stream_capturing()):
            sync()
        rays = ops.RayBundle(batch)
        N, dev = rays.N, rays.device
        near, far = batch['near'], batch['far']
        viewdirs = ops.f32(batch['viewdirs'])
        sdist = weights = None
        prod_num_samples = 1
        renderings: List[Dict] = []
        ray_history: List[Dict] = []
        anneal = (self.anneal_slope * train_frac) / ((self.anneal_slope - 1) * train_frac + 1) \
            if self.anneal_slope > 0 else 1.
        bg = float(self.bg_intensity_range[0])
        draws = None
        if rand and rand_inputs is None:
            # all uniform draws of the pass from ONE generator launch (per level: jitter [N,1] | rotation noise [N,S,7])
            counts = [N * (1 + 7 * (self.num_prop_samples[i] if i < self.num_levels - 1 else self.num_nerf_samples))
                      for i in range(self.num_levels)]
            draws = torch.rand(sum(counts), device=dev).split(counts)
        obj_pose = None
        if self.instance_obj:
            track = curr_track if curr_track is not None else self.tracks
            if track is not None:     # obj_utils.get_pose: per ray, per track
                if torch.is_grad_enabled() and track.requires_grad:
                    # track refinement (Z/train.py:244-257): the blend stays in torch so the gradient of the poses
                    # (csrc/obj.cu k_obj_backward) reaches the track corrections
                    obj_pose = ops.obj_pose_torch(ops.f32(batch['timestamp']), track.to(dev))
                else:
                    obj_pose = ops.obj_pose(batch['timestamp'], track.to(dev))
        for i_level in range(self.num_levels):
            is_prop = i_level < (self.num_levels - 1)
            S = self.num_prop_samples[i_level] if is_prop else self.num_nerf_samples
            dilation = self.dilation_bias + self.dilation_multiplier * 1.0 / prod_num_samples
            prod_num_samples *= S
            use_dilation = (self.dilation_bias > 0 or self.dilation_multiplier > 0) and i_level > 0
            jitter = deg = None
            if rand:
                if rand_inputs is not None:
                    jitter, deg = rand_inputs[i_level]['jitter'], ops.f32(rand_inputs[i_level]['deg'])
                else:
                    jitter, deg = draws[i_level][:N].view(N, 1), draws[i_level][N:].view(N, S, 7)
            sdist, tdist = ops.resample_level(sdist, weights, near, far, S, use_dilation, dilation, anneal, jitter,
                                              bool(rand), self.power_lambda, self.resample_padding)
            if is_prop:
                mlp = self.get_submodule(f'prop_mlp_{i_level}')
                density = ops.prop_level(tdist, deg, mlp, rays, self.std_scale)
                res = dict(density=density, rgb=None, semantic=None, intensity=None)
            else:
                hook = self.__dict__.get('_nlb_before_nerf_table')   # data-parallel trainer: see Trainer.train_step_graphed
                if hook is not None:
                    hook()
                feat = ops.nerf_encode(tdist, deg, self.nerf_mlp.encoder, rays, self.std_scale)
                res = self.nerf_mlp.heads(feat, viewdirs, S)
            obj_mask = None
            if obj_pose is not None:
                apply = ops.obj_apply_train if torch.is_grad_enabled() else ops.obj_apply
                obj_mask = apply(self, res, tdist, rays, viewdirs, obj_pose, is_prop)
            sem = res['semantic'] if (not is_prop and self.config.use_semantic) else None
            inten = res['intensity'] if (not is_prop and self.config.use_intensity) else None
            comp = ops.composite(res['density'], tdist, rays.directions, far, res['rgb'], sem, inten, bg,
                                 self.opaque_background, bool(compute_extras))
            weights = comp['weights']
            rendering = dict(rgb=comp['rgb'], depth=comp['depth'])
            if sem is not None:
                rendering['semantic'] = comp['semantic']
            if inten is not None:
                rendering['intensity'] = comp['intensity']
            if res['rgb'] is None:
                zero = self.__dict__.get('_nlb_zero')
                if zero is None or zero.device != dev:
                    zero = self.__dict__['_nlb_zero'] = torch.zeros(1, device=dev)
                res['rgb'] = zero.expand(N, S, 3)
            if compute_extras:
                rendering['acc'] = comp['acc']
                rendering['distance_mean'] = comp['distance_mean']
                pct = comp['distance_percentiles']
                rendering['distance_percentile_5'] = pct[:, 0]
                rendering['distance_median'] = pct[:, 1]
                rendering['distance_percentile_95'] = pct[:, 2]
                n = self.config.vis_num_rays
                rendering['ray_sdist'] = sdist[:n]
                rendering['ray_weights'] = weights[:n]
                rendering['ray_rgbs'] = res['rgb'][:n]
            if obj_mask is not None:   # Z/internal/models.py:543-545
                res['obj_mask'] = res['instance_mask'] = obj_mask
                rendering['obj_mask'] = rendering['instance_mask'] = obj_mask.sum(-1) > 0
            renderings.append(rendering)
            res['sdist'], res['weights'], res['tdist'] = sdist, weights, tdist
            ray_history.append(res)
        if compute_extras:
            final_rgb = torch.sum(renderings[-1]['ray_rgbs'] * renderings[-1]['ray_weights'][..., None], dim=-2)
            for r in renderings[:-1]:
                r['ray_rgbs'] = torch.broadcast_to(final_rgb[:, None, :], r['ray_rgbs'].shape)
        if self.config.hash_decay_mults > 0 and self.training:
            cached = getattr(self, '_hash_decay_value', None)  # produced by the fused optimizer pass
            renderings[-1]['hash_decay'] = cached if cached is not None else self.hash_decay_loss()
        return renderings, ray_history


class _SingleProcess:
    """Stand-in for accelerate.Accelerator when rendering on one GPU."""
    process_index = 0
    num_processes = 1
    is_main_process = True


def _chunk_outputs(renderings, ray_history, return_weights):
    """What render_image keeps of one chunk (Z/internal/models.py:1440-1452): the final level's rendering, the
    ray_* visualisation leaves of every level, and the final weights on request."""
    out = dict(renderings[-1])
    for k in renderings[0]:
        if k.startswith('ray_'):
            out[k] = [r[k] for r in renderings]
    if return_weights:
        out['weights'] = ray_history[-1]['weights']
    return out


def _forward_chunk(model, rand, chunk, train_frac, return_weights=False):
    """One chunk of rays through Model.forward -> _chunk_outputs.  Deterministic rendering (rand=False) on CUDA is
    replayed as a CUDA graph per chunk shape: ~60 launches per chunk are host-bound when issued eagerly (the
    32 x 1084 LiDAR sweep is three chunks).  The anneal slope travels through the library's dynamic-scalar
    buffer.  The returned tensors live in the graph's static buffers: the caller copies them out before the
    next replay (stream order makes that safe)."""
    first = next(iter(chunk.values()))
    if rand or not first.is_cuda or first.shape[0] == 0 or not getattr(model, 'graph_render', True):
        return _chunk_outputs(*model(rand, chunk, train_frac=train_frac, compute_extras=True, zero_glo=True),
                              return_weights)
    from . import _lib
    dev = first.device
    key = tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(chunk.items()))
    # a graph holds raw parameter pointers: in-place updates (optimizer, load_state_dict) are seen at replay time,
    # re-allocated storage (model.to(...), re-assigned parameters) invalidates every captured graph
    storage = tuple(p.data_ptr() for p in model.parameters())
    if model.__dict__.get('_render_storage') != storage:
        model.__dict__['_render_storage'] = storage
        model.__dict__['_render_graphs'] = {}
        model.__dict__['_render_dyn'] = torch.zeros(4, device=dev)
    cache = model.__dict__.setdefault('_render_graphs', {})
    dyn = model.__dict__['_render_dyn']
    slope = model.anneal_slope
    # by value at enqueue time (a pinned staging buffer would be overwritten by the next call while this one's
    # copies are still queued)
    dyn[0:1].fill_(float((slope * train_frac) / ((slope - 1) * train_frac + 1) if slope > 0 else 1.))
    entry = cache.get(key)
    lib = _lib.load()
    if entry is None:
        static = {k: v.clone() for k, v in chunk.items()}
        lib.nlb_set_dynamic_scalars(dyn.data_ptr())
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                model(rand, static, train_frac=train_frac, compute_extras=True, zero_glo=True)  # warm-up
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outputs = model(rand, static, train_frac=train_frac, compute_extras=True, zero_glo=True)
        finally:
            lib.nlb_set_dynamic_scalars(None)
        entry = cache[key] = (graph, static, outputs)
    graph, static, outputs = entry
    # the graph reads the NerfMLP's packed bf16 operand images, which are refreshed OUTSIDE it: repack (in place,
    # same buffer) when the parameters changed since the last pack -- optimizer steps, load_state_dict
    from . import ops
    for m in model.modules():
        if hasattr(m, '_nlb_packed'):
            ops.nerf_mlp_pack(m)
    for k, v in chunk.items():
        static[k].copy_(v, non_blocking=True)
    graph.replay()
    return _chunk_outputs(*outputs, return_weights)


@torch.no_grad()
def render_image(model, accelerator, batch, rand, config, train_frac=1, verbose=True, return_weights=False,
                 image=True, render_instance=False, instance_id=None):
    """Z/internal/models.py:1379-1507 with one change of schedule: every rank takes
    ONE contiguous slice of the whole ray set, renders it in local chunks of
    config.render_chunk_size and the packed per-ray outputs are all-gathered once
    at the end (the reference pads, slices and all-gathers every leaf of every
    chunk).  `accelerator` needs process_index / num_processes (an
    accelerate.Accelerator works; None = single process); torch.distributed is used
    for the gather when it is initialised."""
    if render_instance:
        raise NotImplementedError('render_instance (object-only rendering) is SURVEY 8(f) "next"')
    from . import parallel
    acc = accelerator if accelerator is not None else _SingleProcess()
    was_training = bool(model.training)
    model.eval()
    model.training = False
    if image:
        height, width = batch['origins'].shape[:2]
        num_rays = height * width
    else:
        num_rays = batch['origins'].shape[0]
    batch = {k: v.reshape((num_rays, -1)) for k, v in batch.items() if v is not None}
    world, rank = int(acc.num_processes), int(acc.process_index)
    lo, hi = parallel.shard_range(num_rays, world, rank)
    rendering, bundles = {}, {}
    for idx0 in range(lo, hi, config.render_chunk_size):
        idx1 = min(hi, idx0 + config.render_chunk_size)
        chunk = {k: v[idx0:idx1] for k, v in batch.items()}
        out = _forward_chunk(model, rand, chunk, train_frac, return_weights)
        for k, v in out.items():
            if isinstance(v, list):     # ray_* bundles: vis_num_rays rays of every chunk, every level
                bundles.setdefault(k, []).append([z.clone() for z in v])
                continue
            if k not in rendering:      # per-ray leaves go straight into this rank's result buffer (no cat)
                rendering[k] = v.new_empty((hi - lo,) + tuple(v.shape[1:]))
            rendering[k][idx0 - lo:idx1 - lo].copy_(v)
    for k, per_chunk in bundles.items():
        rendering[k] = [torch.cat([c[i] for c in per_chunk]) for i in range(len(per_chunk[0]))]
    if world > 1:
        rendering = parallel.gather_rendering(rendering, num_rays, world, rank)
    for k, z in rendering.items():
        if not k.startswith('ray_') and 'hash' not in k:
            rendering[k] = z.reshape((height, width) + z.shape[1:]) if image else z.reshape(num_rays, -1)
    keys = [k for k in rendering if k.startswith('ray_')]
    if keys:
        n = rendering[keys[0]][0].shape[0]
        ray_idx = torch.randperm(n)[:config.vis_num_rays]
        for k in keys:
            rendering[k] = [r[ray_idx.to(r.device)] for r in rendering[k]]
    if was_training:
        model.train()
        model.training = True
    return rendering
