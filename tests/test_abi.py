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
