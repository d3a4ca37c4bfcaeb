"""Configuration: the subset of the reference's `Config` dataclass that the hot path
and its training step read (Z/internal/configs.py:22-212, same field names and
defaults), and a small reader for the reference's gin files.

`gin` is not installed in the build image, so `parse_config_files_and_bindings`
understands the only syntax the reference's configs use -- `Class.attribute =
<python literal>` lines (Z/configs/nuscenes_single.gin) -- and applies them to the
classes registered with `@configurable` (Model, NerfMLP, PropMLP, Config), with
skip_unknown semantics like Z/internal/configs.py:225-226."""
from __future__ import annotations

import ast
import dataclasses
import os
from typing import Dict, Iterable, List, Optional

_REGISTRY: Dict[str, type] = {}
_CONFIG_BINDINGS: Dict[str, object] = {}


def configurable(cls):
    _REGISTRY[cls.__name__] = cls
    return cls


@dataclasses.dataclass
class Config:
    seed: int = 0
    dataset_loader: str = 'llff'
    batch_size: int = 2 ** 16
    patch_size: int = 32
    lidar_supervision: bool = False
    only_lidar_supervison: bool = False
    lidar_batch_ratio: int = 4
    near: float = 2.
    far: float = 6.
    render_chunk_size: int = 16384
    vis_num_rays: int = 16
    max_steps: int = 25000
    data_loss_type: str = 'charb'
    charb_padding: float = 0.001
    data_loss_mult: float = 1.0
    data_coarse_loss_mult: float = 0.
    interlevel_loss_mult: float = 0.0
    anti_interlevel_loss_mult: float = 0.01
    pulse_width: tuple = (0.03, 0.003)
    hash_decay_mults: float = 0.1
    lr_init: float = 0.01
    lr_final: float = 0.001
    lr_delay_steps: int = 5000
    lr_delay_mult: float = 1e-8
    adam_beta1: float = 0.9
    adam_beta2: float = 0.99
    adam_eps: float = 1e-15
    grad_max_norm: float = 0.
    grad_max_val: float = 0.
    distortion_loss_mult: float = 0.005
    disable_multiscale_loss: bool = False
    zero_glo: bool = False
    sample_n_train: int = 7
    sample_m_train: int = 3
    sample_n_test: int = 7
    sample_m_test: int = 3
    pose_refine: bool = True
    t_ratio: float = 0.25
    pn_lr_init: float = 4e-5
    pn_lr_final: float = 2e-6
    start_step: int = 10000
    end_step: int = 20000
    learn_R: bool = True
    learn_t: bool = True
    track_refine: bool = False
    track_start_opt: int = 5000
    tn_lr_init: float = 1e-4
    tn_lr_final: float = 1e-5
    analytic_gradient: bool = True
    use_intensity: bool = False
    no_sem_layer: bool = True
    instance_obj: bool = False
    use_semantic: bool = True
    latent_size: int = 0
    latent_reg: float = 0.001
    obj_nodecay: bool = False
    depth_loss: bool = True
    sem_detach: bool = True
    symmetrize: bool = False
    fuse_render: bool = False

    def __post_init__(self):
        for k, v in _CONFIG_BINDINGS.items():
            if hasattr(self, k):
                setattr(self, k, v)


configurable(Config)


def _apply(lines: Iterable[str], skip_unknown: bool = True):
    for raw in lines:
        line = raw.split('#', 1)[0].strip()
        if not line or '=' not in line:
            continue
        lhs, rhs = [s.strip() for s in line.split('=', 1)]
        if '.' not in lhs:
            continue
        cls_name, attr = lhs.rsplit('.', 1)
        cls_name = cls_name.split('/')[-1]
        try:
            value = ast.literal_eval(rhs)
        except (ValueError, SyntaxError):
            value = rhs  # references (@fn) and bare identifiers are kept as strings
        if cls_name == 'Config':
            _CONFIG_BINDINGS[attr] = value
            continue
        cls = _REGISTRY.get(cls_name)
        if cls is None:
            if skip_unknown:
                continue
            raise KeyError(f'unknown configurable {cls_name}')
        setattr(cls, attr, value)


def parse_config_files_and_bindings(config_files: Optional[List[str]] = None,
                                    bindings: Optional[List[str]] = None, skip_unknown: bool = True):
    for path in config_files or []:
        if not os.path.exists(path):
            alt = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'gin', os.path.basename(path))
            path = alt if os.path.exists(alt) else path
        with open(path) as f:
            _apply(f.readlines(), skip_unknown)
    _apply(bindings or [], skip_unknown)


def clear_config():
    _CONFIG_BINDINGS.clear()


def load_config(gin_configs=None, gin_bindings=None) -> Config:
    """Z/internal/configs.py:223-229."""
    parse_config_files_and_bindings(gin_configs, gin_bindings, skip_unknown=True)
    return Config()


def nuscenes_single(use_intensity: bool = True, instance_obj: bool = False) -> Config:
    """nuscenes_single.gin (+ the two bindings SURVEY 8(d) adds for the static-scene
    hot path) applied programmatically; returns the resulting Config."""
    here = os.path.dirname(os.path.abspath(__file__))
    return load_config([os.path.join(here, 'gin', 'nuscenes_single.gin')],
                       [f'Config.use_intensity = {use_intensity}', f'Config.instance_obj = {instance_obj}'])
