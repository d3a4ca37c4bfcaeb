"""On-GPU ray generation (csrc/raygen.cu, nerf_lidar_b200/raygen.py) against the reference's own
pixels_to_rays / get_directions / cast_lidar_ray_batch through tests/golden/rays_ref.npz (generated from the
imported reference by tests/golden/make_ray_golden.py) and against synthetic.py's numpy restatement, which
tests/test_ray_inputs.py pins to the same fixture.  Bars: equal after the float32 cast up to one ulp
(float64 arithmetic on both sides; the reference's BLAS may order a 3-term sum differently)."""
import os
import sys

import numpy as np
import pytest
import torch

from nerf_lidar_b200 import raygen, synthetic as sy

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden', 'rays_ref.npz')


def _ulp_close(a, b, ulps=1):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return bool(np.all(np.abs(a - b) <= ulps * np.spacing(np.abs(b))))


def _golden_inputs():
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import make_ray_golden as mg
    o, R, px, py = mg.inputs()
    K = np.array([[sy.FOCAL, 0, sy.IMG_W / 2], [0, sy.FOCAL, sy.IMG_H / 2], [0, 0, 1.]])
    c2w = np.concatenate([R, o[:, :, None]], -1)
    return px, py, np.linalg.inv(K), c2w


def test_camera_rays_vs_reference_golden():
    gold = np.load(GOLD)
    px, py, p2c, c2w = _golden_inputs()
    n = px.shape[0]
    dev = 'cuda'
    got = raygen.pixels_to_rays(torch.from_numpy(px).to(dev), torch.from_numpy(py).to(dev),
                                torch.from_numpy(p2c).to(dev), torch.from_numpy(c2w).to(dev),
                                cam_idx=torch.arange(n, device=dev))
    names = ('origins', 'directions', 'viewdirs', 'radii', 'imageplane', 'base_x', 'base_y')
    for k, g in zip(names, got):
        if k == 'imageplane':
            continue
        want = gold['cam_' + k]
        assert tuple(g.shape) == want.shape and g.dtype == torch.float32, k
        assert _ulp_close(g.cpu().numpy(), want), (k, np.abs(g.cpu().numpy() - want.astype(np.float32)).max())
    # imageplane = camera-space xy of the pixel centre (camera_utils.py:527): (x + .5 - cx) / f, -(y + .5 - cy) / f
    ip = got[4].cpu().numpy()
    want_ip = np.stack([(px + 0.5 - sy.IMG_W / 2) / sy.FOCAL, -(py + 0.5 - sy.IMG_H / 2) / sy.FOCAL], -1)
    assert np.allclose(ip, want_ip, rtol=0, atol=1e-6)


def test_camera_rays_shared_matrices_and_patch_shapes():
    """One intrinsics matrix shared by all rays (`batch_index`), [P, 32, 32] pixel blocks, against synthetic.py."""
    rng = np.random.default_rng(3)
    o, R = sy._pose(rng, 4)
    c2w = np.concatenate([R, o[:, :, None]], -1)
    K = np.array([[sy.FOCAL, 0, sy.IMG_W / 2], [0, sy.FOCAL, sy.IMG_H / 2], [0, 0, 1.]])
    x0, y0 = rng.integers(0, sy.IMG_W - 32, 4), rng.integers(0, sy.IMG_H - 32, 4)
    yy, xx = np.meshgrid(np.arange(32), np.arange(32), indexing='ij')
    px, py = x0[:, None, None] + xx, y0[:, None, None] + yy
    cam = np.arange(4)[:, None, None]
    dev = 'cuda'
    got = raygen.cast_ray_batch((torch.from_numpy(np.linalg.inv(K)).to(dev), torch.from_numpy(c2w).to(dev), None, None),
                                dict(pix_x_int=torch.from_numpy(px).to(dev), pix_y_int=torch.from_numpy(py).to(dev),
                                     cam_idx=torch.from_numpy(np.broadcast_to(cam, px.shape).copy()).to(dev)[..., None]))
    want = sy._pix_to_rays(px.reshape(-1).astype(np.float64), py.reshape(-1).astype(np.float64),
                           np.repeat(R, 1024, 0), np.repeat(o, 1024, 0))
    for k in ('origins', 'directions', 'viewdirs', 'radii', 'base_x', 'base_y'):
        g = got[k]
        assert tuple(g.shape[:3]) == (4, 32, 32), k
        assert _ulp_close(g.reshape(-1, g.shape[-1]).cpu().numpy(), want[k]), k


def test_lidar_directions_and_rays_vs_reference_golden():
    gold = np.load(GOLD)
    az = np.linspace(270, -90, 1084) / 180 * np.pi
    d = raygen.get_directions(sy.LIDAR_ELEVATIONS_DEG, az)
    want = gold['lidar_directions']
    assert tuple(d.shape) == want.shape and d.dtype == torch.float32
    assert _ulp_close(d.cpu().numpy(), want)                      # float64 trigonometry, one float32 rounding
    # the batch: feed the reference's own table so that the comparison isolates cast_lidar_ray_batch
    dref = torch.from_numpy(want).cuda()
    b = raygen.cast_lidar_ray_batch(torch.zeros_like(dref), dref, {})
    assert torch.equal(b['directions'], dref) and torch.equal(b['base_x'], dref) and torch.equal(b['base_y'], dref)
    assert np.array_equal(b['radii'].cpu().numpy().astype(np.float64), gold['lidar_radii'].astype(np.float32).astype(np.float64))
    # the reference accumulates the global norm in float32 (np.linalg.norm of a float32 array), the kernel in
    # float64: 1e-6 relative, as for synthetic.py (tests/test_ray_inputs.py)
    assert np.allclose(b['viewdirs'].cpu().numpy(), gold['lidar_viewdirs'], rtol=2e-6, atol=0)
    assert abs(float(b['viewdirs'][0].norm()) * np.sqrt(want.shape[0]) - 1) < 1e-5   # quirk: |viewdirs| = 1/sqrt(N)


def test_empty_and_argument_errors():
    e = torch.empty(0, dtype=torch.int64, device='cuda')
    K = torch.eye(3, dtype=torch.float64, device='cuda')
    P = torch.eye(3, 4, dtype=torch.float64, device='cuda')
    out = raygen.pixels_to_rays(e, e, K, P)
    assert out[0].shape == (0, 3) and out[3].shape == (0, 1)
    one = torch.zeros(1, dtype=torch.int64, device='cuda')
    with pytest.raises(RuntimeError):      # two cameras, no cam_idx
        raygen.pixels_to_rays(one, one, K, torch.stack([P, P]))
    with pytest.raises(RuntimeError):      # host tensors: there is no CPU path
        raygen.pixels_to_rays(one.cpu(), one.cpu(), K, P)
    with pytest.raises(NotImplementedError):
        raygen.pixels_to_rays(one, one, K, P, distortion_params=dict(k1=0.1))


def test_gpu_ray_loader_batch_schema_and_labels():
    """next_train(): the 8192 + 2048 composition and key set of datasets.py:352-403, labels gathered at the
    drawn pixels, rays equal to the reference formula at those pixels."""
    dev = 'cuda'
    g = torch.Generator(device=dev); g.manual_seed(1)
    ncam, H, W = 3, 90, 160
    images = torch.rand(ncam, H, W, 3, device=dev, generator=g)
    depths = torch.rand(ncam, H, W, device=dev, generator=g)
    sem = torch.randint(0, 19, (ncam, H, W), device=dev, generator=g).float()
    masks = (torch.rand(ncam, H, W, device=dev, generator=g) > 0.1).float()
    rng = np.random.default_rng(0)
    o, R = sy._pose(rng, ncam)
    c2w = torch.from_numpy(np.concatenate([R, o[:, :, None]], -1)).to(dev)
    K = np.array([[126.6, 0, W / 2], [0, 126.6, H / 2], [0, 0, 1.]])
    p2c = torch.from_numpy(np.linalg.inv(K)).to(dev)
    ld = sy.lidar_directions(64)
    nl = ld.shape[0]
    lidar = (torch.rand(nl, device=dev, generator=g), torch.zeros(nl, 3, device=dev),
             torch.from_numpy(ld.astype(np.float32)).to(dev), torch.rand(nl, device=dev, generator=g))
    loader = raygen.GpuRayLoader(images, p2c, c2w, sy.NEAR, sy.FAR, depths=depths, semantics=sem, masks=masks,
                                 lidar_depends=lidar, batch_size=8192, patch_size=32, lidar_batch_ratio=4, seed=5)
    b = loader.next_train()
    n = 8192 + 2048
    for k, w in dict(origins=3, directions=3, viewdirs=3, base_x=3, base_y=3, radii=1, near=1, far=1, cam_idx=1,
                     lossmult=1, timestamp=1, rgb=3).items():
        assert tuple(b[k].shape) == (n, w) and b[k].dtype == torch.float32 and b[k].is_cuda, (k, b[k].shape)
    for k in ('depth', 'semantic', 'mask', 'lidar_mask', 'patch_mask', 'intensity'):
        assert tuple(b[k].shape) == (n,), (k, b[k].shape)
    assert int(b['lidar_mask'].sum()) == 2048 and int(b['patch_mask'].sum()) == 2048
    assert bool((b['patch_mask'][:2048] == 1).all())               # patch rays lead the batch (train.py:296-306)
    assert bool((b['semantic'][8192:] == 255).all()) and bool((b['rgb'][8192:] == 0).all())
    # camera rays: recover the pixel from imageplane-free quantities and compare the gathered colour
    cam = b['cam_idx'][:8192, 0].long()
    d_cam = torch.einsum('nji,nj->ni', c2w[cam][:, :, :3].float(), b['directions'][:8192])   # R^T d
    px = torch.round(d_cam[:, 0] / -d_cam[:, 2] * 126.6 + W / 2 - 0.5).long()
    py = torch.round(-d_cam[:, 1] / -d_cam[:, 2] * 126.6 + H / 2 - 0.5).long()
    assert bool((px >= 0).all() and (px < W).all() and (py >= 0).all() and (py < H).all())
    assert torch.equal(b['rgb'][:8192], images[cam, py, px])
    assert torch.equal(b['depth'][:8192], depths[cam, py, px]) and torch.equal(b['mask'][:8192], masks[cam, py, px])
    # the first patch is a contiguous 32 x 32 block of one camera
    assert int(px[:1024].max() - px[:1024].min()) == 31 and int(py[:1024].max() - py[:1024].min()) == 31
    assert int(cam[:1024].min()) == int(cam[:1024].max())
    # the batch steps through the model
    from nerf_lidar_b200 import configs, models
    cfg = configs.nuscenes_single()
    model = models.Model(cfg).cuda()
    model.load_state_dict({k: v.cuda() for k, v in sy.init_state_dict(seed=0, table_std=0.1).items()}, strict=False)
    with torch.no_grad():
        rend, _ = model(False, {k: v[:256].contiguous() for k, v in b.items()}, 1.0, True)
    assert bool(torch.isfinite(rend[-1]['rgb']).all())
