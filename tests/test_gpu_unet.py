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
    with pytest.raises(NotImplementedError):
        raydrop.UNet(6, 2, regression=True)


@pytest.mark.gpu
@pytest.mark.parametrize('bilinear', [True, False])
def test_unet_forward_vs_reference(bilinear):
    import make_unet_golden as mg
    gold = np.load(mg.OUT)
    net = mg.seeded(raydrop.UNet, bilinear).cuda()
    out = net(mg.image().cuda()).cpu().numpy()
    want = gold[f'logits_bilinear{int(bilinear)}']
    assert out.shape == want.shape
    assert np.abs(out - want).max() <= 2e-5, float(np.abs(out - want).max())
    # the decision the drop step takes on these logits (class 1 > class 0) agrees except within the bar of a tie
    flip = (out[0, 1] > out[0, 0]) != (want[0, 1] > want[0, 0])
    assert np.all(np.abs(want[0, 1] - want[0, 0])[flip] <= 4e-5)


@pytest.mark.gpu
def test_unet_batches_shapes_and_errors():
    import make_unet_golden as mg
    net = mg.seeded(raydrop.UNet, True).cuda()
    x = torch.randn(2, 6, 32, 64, device='cuda', generator=torch.Generator(device='cuda').manual_seed(3))
    both = net(x)
    assert both.shape == (2, 2, 32, 64)
    assert torch.allclose(both[1:], net(x[1:].contiguous()), atol=1e-6)     # images of a batch are independent
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 6, 24, 64, device='cuda'))                       # not a multiple of 16
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 5, 32, 64, device='cuda'))
    with pytest.raises(NotImplementedError):
        net.train()(x)


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
