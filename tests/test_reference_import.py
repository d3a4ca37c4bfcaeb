"""The reference's own gridencoder/grid.py picks up nerf_lidar_b200/_gridencoder.py as
its `_gridencoder` backend with zero edits.  Runs only where the
reference tree is present (the build container), in a subprocess so sys.modules
stays clean."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/NeRF_LiDAR/zipnerf'


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_reference_grid_py_binds_to_our_backend():
    code = (
        "import sys, warnings; warnings.filterwarnings('ignore');"
        f"sys.path.insert(0, {os.path.join(ROOT, 'nerf_lidar_b200')!r}); sys.path.insert(0, {REF!r});"
        "import gridencoder.grid as g;"
        "import _gridencoder as be;"
        "assert g._backend is be, g._backend;"
        "assert be.__file__.endswith('nerf_lidar_b200/_gridencoder.py'), be.__file__;"
        "enc = g.GridEncoder(3, 6, 1, base_resolution=16, desired_resolution=512, log2_hashmap_size=21);"
        "assert enc.embeddings.shape[0] == 6606952;"
        "print('ok')"
    )
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-2000:]
