"""ORACLE tooling (test infrastructure, never imported by the product).

Lets the reference's own `internal/models.py` be imported and executed on CPU in
THIS container (from /root/reference) so that golden vectors can be generated
with the reference's own Python.  /root/reference does not exist on the GPU box,
so nothing at test/bench time calls `import_reference()`; only
tests/golden/make_golden.py and tests/test_oracle_vs_reference.py (skipped when
the tree is absent) do.

Missing third-party modules are replaced by minimal stand-ins:
gin (identity decorators), accelerate, torch_scatter.segment_coo, pyquaternion,
skimage.metrics; `gridencoder.GridEncoder` is the torch-CPU restatement in
oracle/grid_oracle.py because the reference's encoder is CUDA-only.
"""
from __future__ import annotations

import os
import sys
import types
from contextlib import contextmanager

import torch

REF_ZIPNERF = '/root/reference/NeRF_LiDAR/zipnerf'


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ZIPNERF, 'internal'))


def _segment_coo(src, index, out=None, dim_size=None, reduce='sum'):
    """torch_scatter.segment_coo for a sorted 1-D index over dim 0."""
    if out is None:
        n = int(dim_size if dim_size is not None else int(index.max()) + 1)
        out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res = torch.zeros_like(out).index_add(0, index, src)
    if reduce == 'mean':
        cnt = torch.zeros(out.shape[0], dtype=src.dtype, device=src.device).index_add(
            0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp_min(1)
        res = res / cnt.reshape((-1,) + (1,) * (src.dim() - 1))
    return res


def install_stubs():
    os.environ.setdefault('TORCHDYNAMO_DISABLE', '1')
    if 'gin' not in sys.modules:
        gin = types.ModuleType('gin')

        def configurable(*args, **kwargs):
            if len(args) == 1 and callable(args[0]) and not kwargs:
                return args[0]
            return lambda f: f

        gin.configurable = configurable
        gin.config = types.SimpleNamespace(external_configurable=lambda f, module=None: f)
        gin.add_config_file_search_path = lambda p: None
        gin.parse_config_files_and_bindings = lambda *a, **k: None
        sys.modules['gin'] = gin
    if 'accelerate' not in sys.modules:
        acc = types.ModuleType('accelerate')
        acc.Accelerator = type('Accelerator', (), {})
        acc.utils = types.SimpleNamespace(send_to_device=lambda b, d: b)
        sys.modules['accelerate'] = acc
    if 'torch_scatter' not in sys.modules:
        ts = types.ModuleType('torch_scatter')
        ts.segment_coo = _segment_coo
        sys.modules['torch_scatter'] = ts
    if 'pyquaternion' not in sys.modules:
        pq = types.ModuleType('pyquaternion')
        pq.Quaternion = type('Quaternion', (), {})
        sys.modules['pyquaternion'] = pq
    if 'skimage' not in sys.modules:
        sk = types.ModuleType('skimage')
        skm = types.ModuleType('skimage.metrics')
        skm.structural_similarity = lambda *a, **k: 0.0
        skm.peak_signal_noise_ratio = lambda *a, **k: 0.0
        sk.metrics = skm
        sys.modules['skimage'] = sk
        sys.modules['skimage.metrics'] = skm
    for name in ('tensorboardX', 'imageio', 'mediapy', 'trimesh', 'matplotlib'):
        pass  # not imported by internal/models.py
    from oracle import grid_oracle
    ge = types.ModuleType('gridencoder')
    ge.GridEncoder = grid_oracle.GridEncoder
    sys.modules['gridencoder'] = ge


def import_reference():
    """Returns the reference's `internal.models` module (and friends) imported
    unmodified from /root/reference."""
    if not reference_available():
        raise RuntimeError('/root/reference is not present')
    install_stubs()
    if REF_ZIPNERF not in sys.path:
        sys.path.insert(0, REF_ZIPNERF)
    import importlib
    models = importlib.import_module('internal.models')
    return models


class RefConfig:
    """The Config fields read by Model/MLP (internal/configs.py:22-212) with the
    nuscenes_single.gin values for the static-scene hot path."""
    use_semantic = True
    analytic_gradient = True
    use_intensity = True
    no_sem_layer = False
    zero_glo = False
    instance_obj = False
    hash_decay_mults = 0.1
    symmetrize = False
    vis_num_rays = 16
    sem_detach = True
    obj_nodecay = True
    latent_size = 0
    fuse_render = False


def apply_gin_bindings(models):
    """Class-attribute equivalents of configs/nuscenes_single.gin:27-34."""
    models.Model.raydist_fn = 'power_transformation'
    models.Model.opaque_background = True
    models.PropMLP.disable_density_normals = True
    models.PropMLP.disable_rgb = True
    models.PropMLP.grid_level_dim = 1
    models.NerfMLP.disable_density_normals = True


@contextmanager
def injected_rand(queue):
    """Replace torch.rand / torch.rand_like by a FIFO of pre-drawn tensors so the
    reference and the kernels see identical jitter (stepfun.py:215-216,
    render.py:150)."""
    q = list(queue)
    real_rand, real_rand_like = torch.rand, torch.rand_like

    def fake_rand(*shape, **kw):
        t = q.pop(0)
        shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        assert tuple(t.shape) == shp, (t.shape, shp)
        return t

    def fake_rand_like(x, **kw):
        t = q.pop(0)
        assert t.shape == x.shape, (t.shape, x.shape)
        return t

    torch.rand, torch.rand_like = fake_rand, fake_rand_like
    try:
        yield q
    finally:
        torch.rand, torch.rand_like = real_rand, real_rand_like
