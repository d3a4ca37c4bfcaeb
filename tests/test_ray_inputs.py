"""The synthetic nuScenes-shaped rays (nerf_lidar_b200/synthetic.py: the bench's and the tests' inputs) against
the reference's OWN ray generation -- camera_utils.pixels_to_rays and lidar_utils.get_directions /
cast_lidar_ray_batch -- through tests/golden/rays_ref.npz, incl. the two LiDAR input quirks SURVEY 8(a) lists
(viewdirs divided by the GLOBAL Frobenius norm, base_x = base_y = directions)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from nerf_lidar_b200 import synthetic as sy

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden', 'rays_ref.npz')
REF = '/root/reference/NeRF_LiDAR/zipnerf'


def _f32_close(a, b):
    """Equal after the batch's float32 cast up to one ulp (the reference divides in float32 where synthetic.py
    divides in float64 and casts)."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.all(np.abs(a - b) <= np.spacing(np.abs(b)))


def test_camera_rays_match_pixels_to_rays():
    gold = np.load(GOLD)
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import make_ray_golden as mg
    o, R, px, py = mg.inputs()
    mine = sy._pix_to_rays(px.astype(np.float64), py.astype(np.float64), R, o)
    for k in ('origins', 'directions', 'viewdirs', 'radii', 'base_x', 'base_y'):
        want = gold['cam_' + k]
        assert mine[k].shape == want.shape, k
        assert np.abs(mine[k] - want).max() <= 1e-12, (k, np.abs(mine[k] - want).max())
        assert _f32_close(mine[k], want), k
    # radii = mean pixel footprint * 2 / sqrt(12) (camera_utils.py:559-562): ~ (1 / f) / sqrt(3) at the centre
    assert abs(float(gold['cam_radii'].mean()) * sy.FOCAL * np.sqrt(3) - 1) < 0.2


def test_lidar_sweep_matches_reference_including_its_quirks():
    gold = np.load(GOLD)
    lr = sy.lidar_rays(np.random.default_rng(0), None, 1084)
    d = gold['lidar_directions']
    assert d.shape == (32 * 1084, 3) and d.dtype == np.float32
    assert np.array_equal(lr['directions'].astype(np.float32), d)           # get_directions, bit for bit
    # the reference sums the 104 064 squares of the float32 array in float32, synthetic.py in float64: the global
    # norm (186.2...) and hence every component agree to 1e-6 relative, not to the ulp
    assert np.allclose(lr['viewdirs'], gold['lidar_viewdirs'], rtol=2e-6, atol=0)
    # quirk 1: ONE norm over the whole [N, 3] array, so |viewdirs| = 1 / sqrt(N), not 1
    assert abs(float(np.linalg.norm(gold['lidar_viewdirs'][0])) * np.sqrt(d.shape[0]) - 1) < 1e-5
    # quirk 2: both cone bases are the direction itself; radii are the constant 5e-4
    assert gold['lidar_base_is_directions'].all()
    assert np.array_equal(lr['base_x'], lr['directions']) and np.array_equal(lr['base_y'], lr['directions'])
    assert np.array_equal(lr['radii'], gold['lidar_radii'])


def test_batches_carry_the_reference_schema():
    """Keys / shapes of the batch dict Model.forward reads and the 8192 + 2048 composition
    (datasets.py:352-403)."""
    b = sy.make_train_batch(8192, seed=0)
    n = 8192 + 2048
    for k, w in dict(origins=3, directions=3, viewdirs=3, base_x=3, base_y=3, radii=1, near=1, far=1, cam_idx=1,
                     lossmult=1, timestamp=1).items():
        assert b[k].shape == (n, w) and b[k].dtype == np.float32, k
    assert int(b['lidar_mask'].sum()) == 2048 and int(b['patch_mask'].sum()) == 2048
    sweep = sy.make_lidar_sweep(seed=0)
    assert sweep['origins'].shape == (32 * 1084, 3) and np.ptp(sweep['origins'], axis=0).max() == 0


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_fixture_reproduces_from_the_live_reference():
    env = dict(os.environ, TORCHDYNAMO_DISABLE='1')
    r = subprocess.run([sys.executable, os.path.join(HERE, 'golden', 'make_ray_golden.py'), '--check'],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-2000:]
