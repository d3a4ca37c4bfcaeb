"""One warm and one measured U-Net inference on a 32 x 1024 sweep (for `ncu --metrics gpu__time_duration.sum`)."""
import sys
import torch
sys.path.insert(0, 'tests/golden'); sys.path.insert(0, '.')
import make_unet_golden as mg
from nerf_lidar_b200 import raydrop
bil = len(sys.argv) < 2 or sys.argv[1] != 'transpose'
net = mg.seeded(raydrop.UNet, bil).cuda()
x = mg.image().cuda()
for _ in range(2):
    net(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    net(x)
e1.record()
torch.cuda.synchronize()
print('unet bilinear=%s: %.3f ms per sweep' % (bil, e0.elapsed_time(e1) / 10))
