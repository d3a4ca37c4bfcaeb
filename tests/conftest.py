import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        skip = pytest.mark.skip(reason='no CUDA device')
        for it in items:
            if 'gpu' in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope='session')
def full_state_dict_visible():
    """nuscenes_single.gin-sized weights with visible table values (seed 12)."""
    from nerf_lidar_b200 import synthetic
    return synthetic.init_state_dict(seed=12, table_std=0.5)
