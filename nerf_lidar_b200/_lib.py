"""ctypes binding of libnlb200.so (include/nlb200.h).

The product path fails loudly when the CUDA library is missing: there is no CPU
or PyTorch fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NLB_LIB') or os.path.join(_HERE, 'libnlb200.so')  # NLB_LIB: dev A/B builds only

c_f = C.c_void_p  # device pointers travel as void*


class NlbRays(C.Structure):
    _fields_ = [('tdist', c_f), ('origins', c_f), ('directions', c_f), ('radii', c_f), ('base_x', c_f),
                ('base_y', c_f), ('deg_noise', c_f), ('N', C.c_int), ('S', C.c_int), ('std_scale', C.c_float),
                ('points_cache', c_f), ('points_mode', C.c_int)]


class NlbTable(C.Structure):
    _fields_ = [('embeddings', c_f), ('offsets', c_f), ('grid_sizes', c_f), ('L', C.c_int), ('C', C.c_int),
                ('H', C.c_uint32), ('S', C.c_float), ('offsets_host', C.POINTER(C.c_int32))]


class NlbLossesIn(C.Structure):
    _fields_ = [(n, c_f) for n in ('rgb', 'depth', 'semantic', 'intensity', 't_rgb', 't_depth', 't_semantic',
                                   't_intensity', 'patch_mask', 'lidar_mask', 'ray_valid')] + \
               [(n, C.c_int) for n in ('N', 'K', 'num_patch', 'patch_size', 'lidar_supervision',
                                       'only_lidar_supervision', 'charb')] + \
               [(n, C.c_float) for n in ('charb_padding', 'depth_mult', 'sem_mult', 'int_mult', 'smooth_mult',
                                         'smo_scale_x', 'smo_scale_y')]


class NlbCompositeIn(C.Structure):
    _fields_ = [('density', c_f), ('tdist', c_f), ('directions', c_f), ('rgb', c_f), ('semantic', c_f),
                ('intensity', c_f), ('far', c_f), ('N', C.c_int), ('S', C.c_int), ('K', C.c_int),
                ('bg', C.c_float), ('opaque_background', C.c_int), ('compute_extras', C.c_int)]


class NlbCompositeOut(C.Structure):
    _fields_ = [('weights', c_f), ('rgb', c_f), ('depth', c_f), ('acc', c_f), ('semantic', c_f),
                ('intensity', c_f), ('distance_mean', c_f), ('distance_percentiles', c_f)]


class NlbCompositeGrad(C.Structure):
    _fields_ = [('g_weights', c_f), ('g_rgb', c_f), ('g_depth', c_f), ('g_acc', c_f), ('g_semantic', c_f),
                ('g_intensity', c_f)]


class NlbNerfMlpWeights(C.Structure):
    _fields_ = [(n, c_f) for n in ('W_d0', 'b_d0', 'W_d2', 'b_d2', 'W_s0', 'b_s0', 'W_s2', 'b_s2', 'W_i0', 'b_i0',
                                   'W_i2', 'b_i2', 'W_v0', 'b_v0', 'W_v1', 'b_v1', 'W_rgb', 'b_rgb')]


class NlbNerfMlpSaved(C.Structure):
    _fields_ = [(n, c_f) for n in ('h0', 'x', 'g', 'h1', 'h2', 'f0')]


class NlbNerfMlpGradIn(C.Structure):
    _fields_ = [(n, c_f) for n in ('g_density', 'g_rgb', 'g_semantic', 'g_intensity', 'density', 'rgb', 'semantic')]


class NlbNerfMlpGradOut(C.Structure):
    _fields_ = [(n, c_f) for n in ('d_rgb', 'd_v1', 'd_v0', 'd_hs1', 'd_g', 'd_x', 'd_h0')] + \
               [(n, C.c_int) for n in ('ld_v1', 'ld_v0', 'ld_g')]


class NlbBf16SumJob(C.Structure):
    _fields_ = [('x', C.c_void_p), ('rows', C.c_int64), ('cols', C.c_int), ('ld', C.c_int), ('group', C.c_int),
                ('out', C.c_void_p)]


class NlbSumTerm(C.Structure):
    _fields_ = [('x', C.c_void_p), ('w', C.c_void_p), ('n', C.c_int64), ('coef', C.c_float), ('out_index', C.c_int)]


class NlbScaleJob(C.Structure):
    _fields_ = [('src', C.c_void_p), ('src2', C.c_void_p), ('dst', C.c_void_p), ('n', C.c_int64), ('g', C.c_void_p),
                ('s', C.c_void_p), ('s2', C.c_void_p), ('coef', C.c_float), ('coef2', C.c_float)]


class NlbRayGrads(C.Structure):
    _fields_ = [(n, c_f) for n in ('origins', 'directions', 'base_x', 'base_y')]


class NlbRayOut(C.Structure):
    _fields_ = [(n, c_f) for n in ('origins', 'directions', 'viewdirs', 'radii', 'imageplane', 'base_x', 'base_y')]


class NlbObjMlp(C.Structure):
    _fields_ = [(n, c_f) for n in ('W_d0', 'b_d0', 'W_d2', 'b_d2', 'W_v0', 'b_v0', 'W_v1', 'b_v1', 'W_rgb', 'b_rgb',
                                   'latent')] + \
               [(n, C.c_int) for n in ('hidden', 'bottleneck', 'view_width', 'deg_view', 'latent_shape', 'latent_tex')] + \
               [(n, C.c_float) for n in ('density_bias', 'rgb_premultiplier', 'rgb_bias', 'rgb_padding')] + \
               [(n, C.c_int) for n in ('class_type', 'class_num')]


class NlbObjGrads(C.Structure):
    _fields_ = [(n, c_f) for n in ('g_W_d0', 'g_b_d0', 'g_W_d2', 'g_b_d2', 'g_W_v0', 'g_b_v0', 'g_W_v1', 'g_b_v1', 'g_W_rgb',
                                   'g_b_rgb', 'g_latent', 'g_table', 'g_pose')]


class NlbRangeImage(C.Structure):
    _fields_ = [(n, c_f) for n in ('proj_range', 'proj_xyz', 'proj_semantic', 'proj_rgb', 'proj_idx', 'proj_mask')]


class NlbUnetConv(C.Structure):
    _fields_ = [('weight', c_f), ('scale', c_f), ('shift', c_f), ('packed', c_f)]


class NlbUnetWeights(C.Structure):
    _fields_ = [('inc', NlbUnetConv * 2), ('down', (NlbUnetConv * 2) * 4), ('up', (NlbUnetConv * 2) * 4),
                ('up_weight', c_f * 4), ('up_bias', c_f * 4), ('up_packed', c_f * 4), ('outc_weight', c_f), ('outc_bias', c_f), ('outr_weight', c_f), ('outr_bias', c_f),
                ('bilinear', C.c_int), ('n_classes', C.c_int)]


_u32, _i, _f, _p = C.c_uint32, C.c_int, C.c_float, C.c_void_p

# name -> (restype, argtypes); mirrors include/nlb200.h one to one
SIGNATURES = {
    'nlb_last_error': (C.c_char_p, []),
    'nlb_version': (_i, []),
    'nlb_device_ok': (_i, []),
    'nlb_grid_encode_forward': (_i, [_p, _p, _p, _p, _u32, _u32, _u32, _u32, _f, _u32, _p, _u32, _i, _u32, _p]),
    'nlb_grid_encode_backward': (_i, [_p, _p, _p, _p, _p, _u32, _u32, _u32, _u32, _f, _u32, _p, _p, _u32, _i, _u32, _p]),
    'nlb_grad_total_variation': (_i, [_p, _p, _p, _p, _f, _u32, _u32, _u32, _u32, _f, _u32, _u32, _i, _p]),
    'nlb_grid_corner_indices': (_i, [_p, _p, _p, _u32, _u32, _u32, _f, _u32, _u32, _i, _p]),
    'nlb_resample': (_i, [_p, _p, _i, _i, _f, _f, _f, _p, _p, _f, _p, _p, _f, _i, _i, _p, _p, _p, _p]),
    'nlb_sorted_interp': (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    'nlb_sample_points': (_i, [C.POINTER(NlbRays), _p, _p]),
    'nlb_encode_forward': (_i, [C.POINTER(NlbRays), C.POINTER(NlbTable), _p, _p]),
    'nlb_encode_backward': (_i, [C.POINTER(NlbRays), C.POINTER(NlbTable), _p, _p, _p, _p]),
    'nlb_encode_backward_workspace_bytes': (C.c_size_t, [C.POINTER(NlbTable)]),
    'nlb_prop_forward': (_i, [C.POINTER(NlbRays), C.POINTER(NlbTable), _p, _p, _p, _p, _p, _p, _p]),
    'nlb_prop_backward': (_i, [C.POINTER(NlbRays), C.POINTER(NlbTable), _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    'nlb_prop_backward_workspace_bytes': (C.c_size_t, [_i, _i, C.POINTER(NlbTable)]),
    'nlb_encode_input_backward': (_i, [C.POINTER(NlbRays), C.POINTER(NlbTable), _p, C.POINTER(NlbRayGrads), _p]),
    'nlb_prop_input_backward': (_i, [C.POINTER(NlbRays), C.POINTER(NlbTable), _p, C.POINTER(NlbRayGrads), _p]),
    'nlb_composite_forward': (_i, [C.POINTER(NlbCompositeIn), C.POINTER(NlbCompositeOut), _p]),
    'nlb_composite_backward': (_i, [C.POINTER(NlbCompositeIn), _p, C.POINTER(NlbCompositeGrad), _p, _p, _p, _p, _p]),
    'nlb_nerf_mlp_packed_bytes': (C.c_size_t, []),
    'nlb_nerf_mlp_pack': (_i, [C.POINTER(NlbNerfMlpWeights), _p, _p]),
    'nlb_nerf_mlp_forward': (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p, C.POINTER(NlbNerfMlpSaved), _p]),
    'nlb_nerf_mlp_packed_transposed_bytes': (C.c_size_t, []),
    'nlb_nerf_mlp_pack_transposed': (_i, [C.POINTER(NlbNerfMlpWeights), _p, _p]),
    'nlb_nerf_mlp_backward': (_i, [C.POINTER(NlbNerfMlpGradIn), C.POINTER(NlbNerfMlpSaved), _i, _p, _p,
                                   C.POINTER(NlbNerfMlpGradOut), _p]),
    'nlb_nerf_mlp_wgrad': (_i, [C.POINTER(NlbNerfMlpSaved), C.POINTER(NlbNerfMlpGradOut), _i,
                                C.POINTER(NlbNerfMlpWeights), _p]),
    'nlb_nerf_mlp_wgrad_finish': (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _p, C.POINTER(NlbNerfMlpWeights), _p]),
    'nlb_colsum_bf16': (_i, [_p, C.c_int64, _i, _i, _p, _p]),
    'nlb_group_sum_bf16': (_i, [_p, C.c_int64, _i, _i, _i, _p, _p]),
    'nlb_bf16_sums': (_i, [C.POINTER(NlbBf16SumJob), _i, _p]),
    'nlb_weighted_sums': (_i, [C.POINTER(NlbSumTerm), _i, _p, _i, _p]),
    'nlb_scale_tensors': (_i, [C.POINTER(NlbScaleJob), _i, _p]),
    'nlb_debug_set_timeline': (_i, [_p]),
    'nlb_distortion_loss': (_i, [_p, _p, _i, _i, _p, _p, _p]),
    'nlb_interlevel_loss': (_i, [_p, _p, _i, _p, _p, _i, _f, _i, _p, _p, _p]),
    'nlb_adam_table_step': (_i, [_p, _p, _p, _p, C.POINTER(C.c_int32), _i, _i, _f, _f, _f, _f, _f, _i, _f, _p, _p]),
    'nlb_adam_table_step_range': (_i, [_p, _p, _p, _p, C.POINTER(C.c_int32), _i, _i, _f, _f, _f, _f, _f, _i, _f, _p,
                                       C.c_int64, C.c_int64, _p]),
    'nlb_render_losses_workspace_bytes': (C.c_size_t, []),
    'nlb_render_losses': (_i, [C.POINTER(NlbLossesIn)] + [_p] * 10),
    'nlb_set_dynamic_scalars': (_i, [_p]),
    'nlb_adam_bias_terms': (_i, [_f, _f, _f, _i, C.POINTER(C.c_float)]),
    'nlb_adam_step': (_i, [_p, _p, _p, _p, C.c_int64, _f, _f, _f, _f, _i, _f, _p]),
    'nlb_camera_rays': (_i, [_p, _p, _p, _p, _i, _p, _i, C.c_int64, C.POINTER(NlbRayOut), _p]),
    'nlb_lidar_directions': (_i, [_p, _i, _p, _i, _p, _p]),
    'nlb_lidar_rays': (_i, [_p, _p, C.c_int64, _p, C.POINTER(NlbRayOut), _p]),
    'nlb_depth_filter': (_i, [_p, _p, _i, _i, _i, _f, _i, _p, _p]),
    'nlb_range_projection_workspace_bytes': (C.c_size_t, [_i, _i]),
    'nlb_range_projection': (_i, [_p, _p, _p, _i, _i, _i, _f, _f, _p, _p, _p, C.POINTER(NlbRangeImage), _p, _p]),
    'nlb_raydrop_select_workspace_bytes': (C.c_size_t, [_i]),
    'nlb_raydrop_select': (_i, [_p, _f, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
    'nlb_unet_workspace_bytes': (C.c_size_t, [_i, _i, _i]),
    'nlb_unet_pack_conv': (_i, [_p, _i, _i, _p, _p]),
    'nlb_unet_pack_convtranspose': (_i, [_p, _i, _i, _p, _p]),
    'nlb_unet_forward': (_i, [_p, C.POINTER(NlbUnetWeights), _i, _i, _i, _i, _p, _p, _p, _p]),
    'nlb_obj_pose': (_i, [_p, _p, _i, _i, _i, _p, _p]),
    'nlb_obj_forward': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, C.POINTER(NlbTable), C.POINTER(NlbObjMlp), _p, _p, _p, _p, _p, _p]),
    'nlb_obj_backward': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, C.POINTER(NlbTable), C.POINTER(NlbObjMlp), _p, _p, _p,
                              C.POINTER(NlbObjGrads), _p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads libnlb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m nerf_lidar_b200.build` '
                '(or __graft_entry__.build()). nerf_lidar_b200 has no CPU/PyTorch fallback.')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


LAUNCHES = 0  # kernels launched through the ABI since import (bench.py reads deltas)


class KernelTimer:
    """CUDA-event timing of named ABI calls on the launching stream (bench.py's
    roofline line).  Enabled by assigning an instance to `_lib.TIMER`."""

    def __init__(self):
        self.events = {}

    def start(self, name):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        self.events.setdefault(name, []).append((e0, e1))
        e0.record(torch.cuda.current_stream())
        return e1

    def summary(self):
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.events.items()}


TIMER: Optional[KernelTimer] = None


class timed:
    """with timed('name'): <one ABI call>"""

    def __init__(self, name):
        self.name = name
        self.end = None

    def __enter__(self):
        if TIMER is not None:
            self.end = TIMER.start(self.name)

    def __exit__(self, *a):
        if self.end is not None:
            self.end.record(torch.cuda.current_stream())
        return False


def check(code: int):
    global LAUNCHES
    LAUNCHES += 1
    if code != 0:
        msg = load().nlb_last_error().decode()
        if code == -2:
            raise NotImplementedError(msg)
        raise RuntimeError(msg)


def ptr(t: Optional[torch.Tensor]):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('nerf_lidar_b200: expected a CUDA tensor (there is no CPU path)')
    if not t.is_contiguous():
        raise RuntimeError('nerf_lidar_b200: expected a contiguous tensor')
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError('nerf_lidar_b200: expected a CUDA tensor (there is no CPU path)')
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
