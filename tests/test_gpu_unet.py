"""The ray-drop U-Net (csrc/unet.cu, nerf_lidar_b200.raydrop.UNet) against the reference's OWN torch U-Net
(NeRF_Lidar_code/src/unet/, eval mode, CPU fp32) -- tests/golden/unet_ref.npz from tests/golden/make_unet_golden.py,
both up-sampling modes, on a 6-channel 32 x 1024 feature map.  Bar: 2e-5 absolute on logits of magnitude 0.1-0.2
(fp32 on both sides, sums of up to 9216 terms in a different order)."""
import os
import sys

import numpy as np
import pytest
import torch

from nerf_lidar_b200 import raydrop

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))


def test_unet_mirror_has_the_reference_parameters():
    """CPU: same constructor order -> the same seeded initialisation and state-dict keys as the reference class."""
    import make_unet_golden as mg
    gold = np.load(mg.OUT)
    for bil in (True, False):
        net = mg.seeded(raydrop.UNet, bil)
        assert abs(float(mg.checksum(net)[0]) - float(gold[f'checksum_bilinear{int(bil)}'][0])) <= 1e-6 * float(mg.checksum(net)[0])
        keys = list(net.state_dict())
        assert 'inc.double_conv.0.weight' in keys and 'down4.maxpool_conv.1.double_conv.4.running_var' in keys
        assert ('up1.up.weight' in keys) == (not bil) and 'outc.conv.bias' in keys
    assert 'outr.conv.weight' in raydrop.UNet(6, 2, regression=True).state_dict()


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
@pytest.mark.parametrize('bilinear', [True, False])
def test_unet_forward_vs_reference(bilinear, precision):
    """Golden = the reference's own U-Net on CPU in fp32.  `fp32`: the FMA kernels, 2e-5 absolute on logits of 0.1-0.2.
    `tf32` (the module's default, like torch's `cudnn.allow_tf32 = True` for the reference's Conv2d on a GPU): the
    tcgen05 kernels read the operands with a 10-bit mantissa; torch's own TF32 run of the same network deviates
    8e-5..1e-4 from its fp32 run on these inputs (tools/unet_vs_torch.py), the bar here is 3e-4."""
    import make_unet_golden as mg
    gold = np.load(mg.OUT)
    net = mg.seeded(raydrop.UNet, bilinear).cuda()
    net.tf32 = precision == 'tf32'
    out = net(mg.image().cuda()).cpu().numpy()
    want = gold[f'logits_bilinear{int(bilinear)}']
    assert out.shape == want.shape
    bar = 2e-5 if precision == 'fp32' else 3e-4
    assert np.abs(out - want).max() <= bar, float(np.abs(out - want).max())
    # the decision the drop step takes on these logits (class 1 > class 0) agrees except within the bar of a tie
    flip = (out[0, 1] > out[0, 0]) != (want[0, 1] > want[0, 0])
    assert np.all(np.abs(want[0, 1] - want[0, 0])[flip] <= 2 * bar)
    if precision == 'tf32':
        # the tensor-core layers against the fp32 kernels on the same device: a different rounding of the same sums
        net.tf32 = False
        ref = net(mg.image().cuda()).cpu().numpy()
        assert 0 < np.abs(out - ref).max() <= bar


@pytest.mark.gpu
def test_unet_tf32_follows_torch_flag():
    """Precision follows torch.backends.cudnn.allow_tf32 like the reference's Conv2d layers; images whose width the
    128-pixel tile does not divide (W = 64 x k is fine, W = 48 is not) fall back to the fp32 kernels per layer."""
    import make_unet_golden as mg
    net = mg.seeded(raydrop.UNet, True).cuda()
    x = mg.image().cuda()
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = False
        a = net(x)
        net.tf32 = False
        assert torch.equal(a, net(x))
        net.tf32 = None
        torch.backends.cudnn.allow_tf32 = True
        b = net(x)
        assert not torch.equal(a, b) and float((a - b).abs().max()) < 3e-4
    finally:
        torch.backends.cudnn.allow_tf32 = old
    y = torch.randn(1, 6, 48, 208, device='cuda', generator=torch.Generator(device='cuda').manual_seed(5))
    net.tf32 = True
    t = net(y)
    net.tf32 = False
    assert float((t - net(y)).abs().max()) < 3e-4


@pytest.mark.gpu
def test_unet_batches_shapes_and_errors():
    import make_unet_golden as mg
    net = mg.seeded(raydrop.UNet, True).cuda()
    x = torch.randn(2, 6, 32, 64, device='cuda', generator=torch.Generator(device='cuda').manual_seed(3))
    both = net(x)
    assert both.shape == (2, 2, 32, 64)
    # images of a batch are independent (the split of the input channels depends on the batch size: the sums are
    # the same up to fp32 rounding of a different association, 3e-6 measured on the TF32 path, 5e-8 on the fp32 one)
    assert torch.allclose(both[1:], net(x[1:].contiguous()), atol=2e-5)
    net.tf32 = False
    assert torch.allclose(net(x)[1:], net(x[1:].contiguous()), atol=1e-6)
    net.tf32 = None
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 6, 24, 64, device='cuda'))                       # not a multiple of 16
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 5, 32, 64, device='cuda'))


@pytest.mark.gpu
@pytest.mark.parametrize('shape,bilinear', [((2, 6, 32, 1024), True), ((1, 6, 64, 2048), False), ((3, 6, 16, 128), True),
                                            ((1, 6, 48, 64), False)])
def test_unet_tf32_other_shapes(shape, bilinear):
    """The tensor-core path at other image shapes and batch sizes (batched sweeps, a 64-beam 2048-column image, the
    smallest image whose deepest level is still one 2 x 64 tile wide enough -- 16 x 128 is not: its levels below
    W = 64 run on the fp32 kernels -- and a height the 2-row tile does not divide) against the fp32 kernels."""
    import make_unet_golden as mg
    net = mg.seeded(raydrop.UNet, bilinear).cuda()
    x = torch.randn(*shape, device='cuda', generator=torch.Generator(device='cuda').manual_seed(shape[2] + shape[3]))
    net.tf32 = True
    a = net(x)
    net.tf32 = False
    b = net(x)
    assert a.shape == (shape[0], 2, shape[2], shape[3]) and torch.isfinite(a).all()
    scale = float(b.abs().max())
    assert float((a - b).abs().max()) <= 2e-3 * max(scale, 0.1), (float((a - b).abs().max()), scale)
    assert not torch.equal(a, b)


@pytest.mark.gpu
def test_unet_training_mode_and_regression_head():
    """Training mode evaluates the same layers with torch (autograd, batch statistics), as the reference trains this
    network (R/src/model/ray_drop_train.py:73-125); in eval mode the kernels agree with that torch evaluation, incl.
    the range-regression head sigmoid(outr(x)) of UNet(regression=True)."""
    import make_unet_golden as mg
    torch.manual_seed(0)
    net = raydrop.UNet(6, 2, bilinear=True, regression=True).cuda()
    x = mg.image().cuda()
    gt = (torch.rand(1, 32, 1024, device='cuda') > 0.5).long()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    net.train()
    losses = []
    for _ in range(3):
        logits, reg = net(x)
        loss = torch.nn.functional.cross_entropy(logits, gt) + (reg[:, 0] - 0.5).abs().mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0]
    net.eval()
    net.tf32 = False
    logits, reg = net(x)                                    # csrc/unet.cu on the updated weights and statistics
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            want_l, want_r = net._forward_torch(x)          # torch, eval-mode BatchNorm
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert reg.shape == (1, 1, 32, 1024) and float(reg.min()) >= 0 and float(reg.max()) <= 1
    scale = float(want_l.abs().max())
    assert float((logits - want_l).abs().max()) <= 1e-4 * max(scale, 1.0), (float((logits - want_l).abs().max()), scale)
    assert float((reg - want_r).abs().max()) <= 1e-5


@pytest.mark.gpu
def test_sweep_to_dropped_cloud_pipeline():
    """depth filter -> range projection -> U-Net on the projected features -> drop selection, all on the device."""
    import make_raydrop_golden as rg
    import make_unet_golden as mg
    pts, sem, rgb, _ = rg.inputs()
    t = lambda a: torch.from_numpy(a).cuda()
    pts, sem, rgb = t(pts), t(sem), t(rgb)
    fm = raydrop.depth_filter(pts, sem, return_mask=True, width=1, threshold=1)
    scan = raydrop.LaserScan(H=rg.H, W=rg.W, fov_up=10.67, fov_down=-30.67)
    scan.set_points(pts, semantic=sem, rgb=rgb)
    scan.do_range_projection()
    feats = torch.cat([scan.proj_range[None], scan.proj_semantic[None], scan.proj_mask[None], scan.proj_rgb.permute(2, 0, 1)], 0)
    net = mg.seeded(raydrop.UNet, True).cuda()
    logits = net(feats[None].contiguous())[0]
    kept_p, kept_l = raydrop.drop_rays(logits, scan, pts, sem, fm, mask_thre=0.5)
    assert 0 < kept_p.shape[0] < pts.shape[0] and kept_l.shape[0] == kept_p.shape[0]
    assert not bool((kept_l == 10).any())
