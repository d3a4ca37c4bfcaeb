"""The C-ABI library loads and exports every symbol include/nlb200.h declares
(no compute calls: there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'nlb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(nlb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from nerf_lidar_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in nlb200.h but not exported'
    # and the python binding table covers the same set
    assert sorted(_lib.SIGNATURES) == names


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'nerf_lidar_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f


def test_cuda_ops_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from nerf_lidar_b200 import ops
    with pytest.raises(RuntimeError):
        ops.sorted_interp(torch.zeros(1, 4), torch.zeros(1, 4), torch.zeros(1, 4))


def _lib():
    from nerf_lidar_b200 import _lib as L, build
    build.build()
    return L, L.load()


def test_argument_errors_are_reported_before_any_launch():
    """Error behaviour of the C ABI, checkable without a device: every entry point validates its arguments before
    touching CUDA and reports like the reference's bindings do (a RuntimeError there = NLB_EINVAL + message here:
    `GridEncoding: C must be 1, 2, 4, or 8`, gridencoder.cu:381-398; D > 3 is NLB_EUNSUPPORTED in this build)."""
    L, lib = _lib()
    EINVAL, EUNSUP = -1, -2
    # _gridencoder ABI: level_dim / input_dim checks come first, empty batches are accepted, null pointers are not
    assert lib.nlb_grid_encode_forward(None, None, None, None, 8, 3, 3, 4, 1.0, 16, None, 0, 0, 0, None) == EINVAL
    assert b'C must be 1, 2, 4, or 8' in lib.nlb_last_error()
    assert lib.nlb_grid_encode_forward(None, None, None, None, 8, 4, 2, 4, 1.0, 16, None, 0, 0, 0, None) == EUNSUP
    assert lib.nlb_grid_encode_forward(None, None, None, None, 0, 3, 2, 4, 1.0, 16, None, 0, 0, 0, None) == 0
    assert lib.nlb_grid_encode_forward(None, None, None, None, 8, 3, 2, 4, 1.0, 16, None, 0, 0, 0, None) == EINVAL
    assert b'null pointer' in lib.nlb_last_error()
    assert lib.nlb_grid_encode_backward(None, None, None, None, None, 8, 3, 5, 4, 1.0, 16, None, None, 0, 0, 0, None) == EINVAL
    # fused path: descriptors
    assert lib.nlb_encode_forward(None, None, None, None) == EINVAL
    assert b'null descriptor' in lib.nlb_last_error()
    offs = (ctypes.c_int32 * 3)(0, 4920, 4920 + 3000)          # second level: 3000 rows, hashed, not a power of two
    fake = ctypes.c_void_p(256)   # stands for device memory: never dereferenced on the host
    rays = L.NlbRays(fake, fake, fake, fake, fake, fake, None, 4, 8, 0.35, None, 0)
    tab = L.NlbTable(fake, fake, fake, 2, 1, 16, 1.0, offs)
    assert lib.nlb_encode_forward(ctypes.byref(rays), ctypes.byref(tab), fake, None) == EUNSUP
    assert b'power-of-two' in lib.nlb_last_error()
    tab_no_host = L.NlbTable(fake, fake, fake, 2, 1, 16, 1.0, None)
    assert lib.nlb_encode_forward(ctypes.byref(rays), ctypes.byref(tab_no_host), fake, None) == EINVAL
    assert b'offsets_host' in lib.nlb_last_error()
    tab_big = L.NlbTable(fake, fake, fake, 17, 1, 16, 1.0, offs)
    assert lib.nlb_encode_forward(ctypes.byref(rays), ctypes.byref(tab_big), fake, None) == EUNSUP
    rays_cache = L.NlbRays(fake, fake, fake, fake, fake, fake, None, 4, 8, 0.35, None, 1)   # mode 1 without a cache
    ok_offs = (ctypes.c_int32 * 3)(0, 4920, 4920 + 4096)
    tab_ok = L.NlbTable(fake, fake, fake, 2, 1, 16, 1.0, ok_offs)
    assert lib.nlb_encode_forward(ctypes.byref(rays_cache), ctypes.byref(tab_ok), fake, None) == EINVAL
    assert b'points_cache' in lib.nlb_last_error()
    # proposal level: PropMLP tables have level_dim 1
    tab_c4 = L.NlbTable(fake, fake, fake, 2, 4, 16, 1.0, ok_offs)
    assert lib.nlb_prop_forward(ctypes.byref(rays), ctypes.byref(tab_c4), fake, fake, fake, fake, fake, None, None) == EINVAL
    assert b'level_dim 1' in lib.nlb_last_error()
    # empty ray batch: accepted by the fused entry points with null data pointers
    empty = L.NlbRays(None, None, None, None, None, None, None, 0, 8, 0.35, None, 0)
    assert lib.nlb_encode_forward(ctypes.byref(empty), ctypes.byref(tab_ok), None, None) == 0
    assert lib.nlb_encode_backward(ctypes.byref(empty), ctypes.byref(tab_ok), None, None, None, None) == 0
    # compositing
    cin = L.NlbCompositeIn(fake, fake, fake, None, None, None, None, 4, 8, 0, 1.0, 1, 0)
    assert lib.nlb_composite_forward(ctypes.byref(cin), None, None) == EINVAL
    # workspace queries are pure host arithmetic
    assert lib.nlb_encode_backward_workspace_bytes(ctypes.byref(tab_ok)) >= 0
    assert lib.nlb_prop_backward_workspace_bytes(4, 8, ctypes.byref(tab_ok)) > 0


def test_gridencoder_backend_checks_like_the_reference_binding():
    """nerf_lidar_b200/_gridencoder.py = the module Z/gridencoder/grid.py imports as `_gridencoder`: the same three
    functions with the argument order of bindings.cpp:5-9, and the CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_*
    guards of gridencoder.cu:15-18 as RuntimeError (this is what a CPU box can exercise of them)."""
    import inspect
    import torch
    from nerf_lidar_b200 import _gridencoder as be
    assert list(inspect.signature(be.grid_encode_forward).parameters) == [
        'inputs', 'embeddings', 'offsets', 'outputs', 'B', 'D', 'C', 'L', 'S', 'H', 'dy_dx', 'gridtype',
        'align_corners', 'interp']
    assert list(inspect.signature(be.grid_encode_backward).parameters) == [
        'grad', 'inputs', 'embeddings', 'offsets', 'grad_embeddings', 'B', 'D', 'C', 'L', 'S', 'H', 'dy_dx',
        'grad_inputs', 'gridtype', 'align_corners', 'interp']
    assert list(inspect.signature(be.grad_total_variation).parameters) == [
        'inputs', 'embeddings', 'grad', 'offsets', 'weight', 'B', 'D', 'C', 'L', 'S', 'H', 'gridtype', 'align_corners']
    x, emb = torch.zeros(8, 3), torch.zeros(64, 2)
    offs, out = torch.zeros(3, dtype=torch.int32), torch.zeros(2, 8, 2)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        be.grid_encode_forward(x, emb, offs, out, 8, 3, 2, 2, 1.0, 16, None, 0, False, 0)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        be.grid_encode_backward(out, x, emb, offs, torch.zeros_like(emb), 8, 3, 2, 2, 1.0, 16, None, None, 0, False, 0)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        be.grad_total_variation(x, emb, torch.zeros_like(emb), offs, 1.0, 8, 3, 2, 2, 1.0, 16, 0, False)
