"""The callers either side of the hot path, chained on the device (SURVEY 8f #2 - #4 around 8a): batches from the
device-resident loader (raygen.GpuRayLoader) -> captured training steps (Trainer.train_step_graphed) -> LiDAR sweep
render (render_image) -> stage 3 (depth filter, range projection, ray-drop U-Net, drop selection).  No parity claim
here (every stage has its own test against the reference); this pins the interfaces between the stages."""
import numpy as np
import pytest
import torch

from nerf_lidar_b200 import configs, models, raydrop, raygen, synthetic as sy, train

pytestmark = pytest.mark.gpu


def test_loader_train_render_raydrop():
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    ncam, H, W = 4, 90, 160
    rng = np.random.default_rng(0)
    o, R = sy._pose(rng, ncam)
    c2w = torch.from_numpy(np.concatenate([R, o[:, :, None]], -1)).to(dev)
    K = np.array([[126.6, 0, W / 2], [0, 126.6, H / 2], [0, 0, 1.]])
    width = 256
    ld = sy.lidar_directions(width).astype(np.float32)
    nl = ld.shape[0]
    loader = raygen.GpuRayLoader(
        torch.rand(ncam, H, W, 3, device=dev, generator=g), torch.from_numpy(np.linalg.inv(K)).to(dev), c2w, sy.NEAR, sy.FAR,
        depths=torch.rand(ncam, H, W, device=dev, generator=g) * 3, semantics=torch.randint(0, 19, (ncam, H, W), device=dev, generator=g).float(),
        masks=torch.ones(ncam, H, W, device=dev),
        lidar_depends=(torch.rand(nl, device=dev, generator=g) * 3 + 0.1, torch.zeros(nl, 3, device=dev), torch.from_numpy(ld).to(dev),
                       torch.rand(nl, device=dev, generator=g)),
        batch_size=4096, patch_size=32, lidar_batch_ratio=4, seed=1)
    cfg = configs.nuscenes_single()
    model = models.Model(cfg, training=True).cuda()
    model.load_state_dict({k: v.cuda() for k, v in sy.init_state_dict(seed=0, table_std=0.05).items()}, strict=False)
    tr = train.Trainer(model, cfg)
    num_patch = (4096 // 4) // 1024
    losses = []
    for i in range(4):                       # the first call captures, the others replay; batches keep their shapes
        b = loader.next_train()
        assert b['origins'].shape[0] == 4096 + 1024
        out = tr.train_step_graphed(b, 6000 + i, num_patch)
        losses.append(float(out['loss']))
    assert all(np.isfinite(losses)), losses
    # LiDAR sweep render (BASELINE configs[2]) from rays generated on the device
    dirs = raygen.get_directions(sy.LIDAR_ELEVATIONS_DEG, np.linspace(270, -90, width) / 180 * np.pi)
    n = dirs.shape[0]
    full = lambda v: torch.full((n, 1), float(v), device=dev)
    sweep = raygen.cast_lidar_ray_batch(torch.zeros(n, 3, device=dev), dirs, dict(near=full(sy.NEAR), far=full(sy.FAR), lossmult=full(1.)))
    sweep = {k: v for k, v in sweep.items() if v is not None}
    sweep['cam_idx'], sweep['timestamp'] = full(-1), full(0)
    rend = models.render_image(model, None, sweep, False, cfg, image=False, verbose=False)
    depth = rend['depth'].reshape(-1)
    assert depth.shape[0] == n and bool(torch.isfinite(depth).all())
    # stage 3 on the rendered sweep (render_lidar.py writes points = origin + depth * direction and their labels)
    pts = (dirs * depth[:, None] / sy.SCENE_SCALE).contiguous()
    labels = rend['semantic'].argmax(-1).float()
    fm = raydrop.depth_filter(pts, labels, return_mask=True, width=1, threshold=1)
    scan = raydrop.LaserScan(H=32, W=width, fov_up=10.67, fov_down=-30.67)
    scan.set_points(pts, semantic=labels, rgb=rend['rgb'].reshape(-1, 3))
    scan.do_range_projection()
    feats = torch.cat([scan.proj_range[None], scan.proj_semantic[None], scan.proj_mask[None], scan.proj_rgb.permute(2, 0, 1)], 0)
    torch.manual_seed(0)
    net = raydrop.UNet(n_channels=6, n_classes=2, bilinear=True).cuda().eval()
    logits = net(feats[None].contiguous())[0]
    kept_p, kept_l = raydrop.drop_rays(logits, scan, pts, labels, fm, mask_thre=0.5)
    assert kept_p.shape[0] <= n and kept_p.shape[0] == kept_l.shape[0]
