"""The ray-drop U-Net of csrc/unet.cu against the SAME layers evaluated by torch (cuDNN) on the same device -- what the
reference's `UNet.forward` (R/src/unet/unet_model.py:34-47) runs on a GPU -- in strict fp32 and with cuDNN's default
TF32 convolutions.  Prints time per 32 x 1024 sweep and the largest logit difference against the strict-fp32 run."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, 'tests/golden'); sys.path.insert(0, '.')
import make_unet_golden as mg
from nerf_lidar_b200 import raydrop


def torch_forward(net, x):
    """Unet.forward of the reference, on the product's parameter containers."""
    dc = lambda m, t: m.double_conv(t)
    x1 = dc(net.inc, x)
    xs = [x1]
    for d in (net.down1, net.down2, net.down3, net.down4):
        xs.append(dc(d.maxpool_conv[1], d.maxpool_conv[0](xs[-1])))
    y = xs[-1]
    for u, skip in zip((net.up1, net.up2, net.up3, net.up4), xs[-2::-1]):
        y = u.up(y)
        dy, dx = skip.shape[2] - y.shape[2], skip.shape[3] - y.shape[3]
        y = F.pad(y, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        y = dc(u.conv, torch.cat([skip, y], 1))
    return net.outc.conv(y)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for bil in (True, False):
    net = mg.seeded(raydrop.UNet, bil).cuda()
    x = mg.image().cuda()
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        ms32, ref = timeit(lambda: torch_forward(net, x))
        torch.backends.cudnn.allow_tf32 = True
        ms_tf, out_tf = timeit(lambda: torch_forward(net, x))
        torch.backends.cudnn.benchmark = True
        ms_tfb, _ = timeit(lambda: torch_forward(net, x))
        torch.backends.cudnn.benchmark = False
        ms_nlb, out = timeit(lambda: net(x))
    print(f'bilinear={bil}: csrc/unet.cu {ms_nlb:.3f} ms (max |d logit| vs torch fp32 {float((out - ref).abs().max()):.2e}) | '
          f'torch cuDNN fp32 {ms32:.3f} ms | torch cuDNN TF32 (torch default) {ms_tf:.3f} ms '
          f'(max |d logit| {float((out_tf - ref).abs().max()):.2e}) | TF32 + cudnn.benchmark {ms_tfb:.3f} ms', flush=True)
