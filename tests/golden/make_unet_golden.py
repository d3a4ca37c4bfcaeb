"""Generates tests/golden/unet_ref.npz: logits of the reference's OWN ray-drop U-Net (NeRF_Lidar_code/src/unet/,
imported unmodified) in eval mode on CPU, fp32, for both up-sampling modes, on a seeded 6-channel 32 x 1024
range-image feature map.  The weights are the module's torch-seeded initialisation plus seeded BatchNorm statistics;
nerf_lidar_b200.raydrop.UNet constructs its layers in the same order, so the same seed gives the same 17-31 M
parameters on both sides (checked through a checksum stored in the fixture).
  python tests/golden/make_unet_golden.py [--check]"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/NeRF_LiDAR/NeRF_Lidar_code/src'
OUT = os.path.join(HERE, 'unet_ref.npz')
H, W, CIN = 32, 1024, 6


def seeded(cls, bilinear):
    """UNet(6, 2, bilinear) under torch.manual_seed(0), then non-trivial BatchNorm statistics / affine terms."""
    torch.manual_seed(0)
    net = cls(n_channels=CIN, n_classes=2, bilinear=bilinear)
    g = torch.Generator().manual_seed(1)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return net.eval()


def image():
    g = torch.Generator().manual_seed(2)
    return torch.randn(1, CIN, H, W, generator=g)


def checksum(net):
    return np.array([float(sum(p.double().abs().sum() for p in net.state_dict().values() if p.dtype.is_floating_point))])


def run_reference():
    sys.path.insert(0, REF)
    UNet = importlib.import_module('unet').UNet if hasattr(importlib.import_module('unet'), 'UNet') else \
        importlib.import_module('unet.unet_model').UNet
    out = {}
    for bil in (True, False):
        net = seeded(UNet, bil)
        with torch.no_grad():
            out[f'logits_bilinear{int(bil)}'] = net(image()).numpy()
        out[f'checksum_bilinear{int(bil)}'] = checksum(net)
    return out


if __name__ == '__main__':
    got = run_reference()
    if '--check' in sys.argv:
        gold = np.load(OUT)
        for k in gold.files:
            assert np.allclose(gold[k], got[k], rtol=0, atol=1e-6), k      # (CPU conv thread count may reorder sums)
        print('ok')
    else:
        np.savez_compressed(OUT, **got)
        print('wrote', OUT, os.path.getsize(OUT) // 1024, 'KiB', {k: v.shape for k, v in got.items()})
        print({k: float(np.abs(v).max()) for k, v in got.items()})
