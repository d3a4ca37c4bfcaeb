"""The gin reader and the Config dataclass of the drop-in (nerf_lidar_b200/configs.py) against the reference's own
files: its nuscenes_single.gin is read unchanged, and every field the subset defines has the reference's name and
default (Z/internal/configs.py:22-212).  Each case runs in a subprocess: bindings are process-global state, as
with gin itself."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/NeRF_LiDAR/zipnerf'


def _run(code, timeout=300):
    env = dict(os.environ, TORCHDYNAMO_DISABLE='1')
    r = subprocess.run([sys.executable, '-c', f'import sys; sys.path.insert(0, {ROOT!r})\n' + code],
                       capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_in_repo_gin_and_bindings():
    _run('''
from nerf_lidar_b200 import configs, models
cfg = configs.nuscenes_single()
assert cfg.instance_obj is False and cfg.use_intensity is True          # the two bindings of SURVEY 8(d)
assert cfg.use_semantic and cfg.no_sem_layer is False and cfg.lidar_supervision and cfg.lidar_batch_ratio == 4
assert (cfg.start_step, cfg.end_step, cfg.patch_size, cfg.latent_size) == (0, 5000, 32, 128)
assert models.Model.raydist_fn == 'power_transformation' and models.Model.opaque_background is True
assert models.PropMLP.grid_level_dim == 1 and models.PropMLP.disable_rgb and models.PropMLP.disable_density_normals
assert models.NerfMLP.disable_density_normals
# syntax the reference's files use: no spaces around '=', trailing comments, scoped names, unknown classes
configs.parse_config_files_and_bindings(None, ['Config.near=0.25  # metres', 'train/Config.far = 7',
                                               'ObjMLP.grid_level_dim = 2', 'Config.exp_name = test3'])
cfg = configs.Config()
assert cfg.near == 0.25 and cfg.far == 7
try:
    configs.parse_config_files_and_bindings(None, ['Nope.x = 1'], skip_unknown=False)
    raise SystemExit('unknown configurable accepted')
except KeyError:
    pass
configs.clear_config()
assert configs.Config().near == 2.0
print('ok')
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_reference_gin_file_reads_unchanged():
    _run(f'''
import os
from nerf_lidar_b200 import configs, models
ours = configs.nuscenes_single(use_intensity=False, instance_obj=True)
ours = dict(vars(ours))
cls_ours = {{(c.__name__, a): getattr(c, a) for c in (models.Model, models.PropMLP, models.NerfMLP)
            for a in ('raydist_fn', 'opaque_background', 'grid_level_dim', 'disable_rgb', 'disable_density_normals')
            if hasattr(c, a)}}
configs.clear_config()
ref = configs.load_config([os.path.join({REF!r}, 'configs', 'nuscenes_single.gin')], [])
assert dict(vars(ref)) == ours, {{k: (v, ours[k]) for k, v in vars(ref).items() if ours[k] != v}}
cls_ref = {{(c.__name__, a): getattr(c, a) for c in (models.Model, models.PropMLP, models.NerfMLP)
           for a in ('raydist_fn', 'opaque_background', 'grid_level_dim', 'disable_rgb', 'disable_density_normals')
           if hasattr(c, a)}}
assert cls_ref == cls_ours
print('ok')
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_config_fields_have_the_reference_defaults():
    _run('''
import dataclasses, importlib
from oracle import ref_shims
from nerf_lidar_b200 import configs
ref_shims.import_reference()
rc = importlib.import_module('internal.configs')
ref_fields = {f.name: f for f in dataclasses.fields(rc.Config)}
extra = {'depth_loss', 'fuse_render'}   # switches of this build, not reference fields
bad = []
for f in dataclasses.fields(configs.Config):
    if f.name in extra and f.name not in ref_fields:
        continue
    if f.name in ref_fields:
        r = ref_fields[f.name]
        rd = r.default if r.default is not dataclasses.MISSING else (r.default_factory() if r.default_factory is not dataclasses.MISSING else None)
    elif hasattr(rc.Config, f.name):
        rd = getattr(rc.Config, f.name)   # un-annotated class attributes of the reference (seed, pulse_width)
    else:
        bad.append((f.name, 'not in the reference Config'))
        continue
    if rd != f.default and not (isinstance(rd, (tuple, list)) and tuple(rd) == tuple(f.default)):
        bad.append((f.name, f.default, rd))
assert not bad, bad
print('ok')
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_schedules_match_the_reference_math():
    """train.learning_rate_decay / log_lerp against Z/internal/math.py:41-86 over the whole run (the gin's
    lr_init 0.01 -> lr_final 0.001 with the 5000-step delay)."""
    _run('''
import importlib, math, sys
sys.path.insert(0, '/root/reference/NeRF_LiDAR/zipnerf')
rm = importlib.import_module('internal.math')
from nerf_lidar_b200 import configs, train
c = configs.nuscenes_single()
for step in list(range(0, 200)) + list(range(200, 25001, 137)) + [25000, 30000]:
    want = float(rm.learning_rate_decay(step, c.lr_init, c.lr_final, c.max_steps, c.lr_delay_steps, c.lr_delay_mult))
    got = train.learning_rate_decay(step, c.lr_init, c.lr_final, c.max_steps, c.lr_delay_steps, c.lr_delay_mult)
    assert abs(got - want) <= 1e-12 * max(abs(want), 1e-30), (step, got, want)
    assert abs(train.learning_rate_decay(step, 1e-2, 1e-3, 25000) - float(rm.learning_rate_decay(step, 1e-2, 1e-3, 25000))) <= 1e-15
for t in (-0.5, 0., 0.3, 1., 1.7):
    assert abs(train.log_lerp(t, 0.03, 0.003) - float(rm.log_lerp(t, 0.03, 0.003))) <= 1e-15
print('ok')
''')


def test_fused_mlp_rejects_other_architectures():
    """The tcgen05 NerfMLP kernels are compiled for the nuscenes_single.gin architecture and take raw pointers:
    a gin binding that changes a layer shape or a baked-in scalar must raise before any launch."""
    import pytest
    import torch
    from nerf_lidar_b200 import configs, models
    cfg = configs.nuscenes_single()
    ok = models.Model(cfg).nerf_mlp
    ok._check_fused_shapes()
    for attr, value in (('bottleneck_width', 128), ('net_width_viewdirs', 128), ('class_num', 12), ('deg_view', 2),
                        ('density_bias', 0.0), ('rgb_padding', 0.01), ('grid_level_dim', 2), ('net_depth_viewdirs', 3),
                        ('mlp_dtype', torch.float32)):
        old = getattr(models.NerfMLP, attr)
        setattr(models.NerfMLP, attr, value)
        try:
            mlp = models.Model(cfg).nerf_mlp
            with pytest.raises(NotImplementedError, match='built for'):
                mlp.heads(torch.zeros(32, mlp.encoder.output_dim), torch.zeros(1, 3), 32)
        finally:
            setattr(models.NerfMLP, attr, old)
