"""Dev tool (GPU): compositing / resampling kernels at >= 1 M rays (bandwidth- not launch-bound),
algorithmic GB/s against the measured HBM peak."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6650.0) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else 6650.0


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

g = torch.Generator(device='cuda').manual_seed(0)
dirs = torch.nn.functional.normalize(torch.randn(N, 3, device='cuda', generator=g), dim=-1)
far = torch.full((N,), 8.33, device='cuda')
near = torch.full((N,), 0.033, device='cuda')
for S, K, name, alg in ((32, 19, 'nerf level (rgb + 19 classes + intensity)', 3464), (64, 0, 'proposal level', 820)):
    t = torch.sort(torch.rand(N, S + 1, device='cuda', generator=g) * 8 + 0.05, -1).values
    dens = torch.rand(N, S, device='cuda', generator=g) * 5
    rgb = torch.rand(N, S, 3, device='cuda', generator=g) if K else None
    sem = torch.softmax(torch.randn(N, S, K, device='cuda', generator=g), -1) if K else None
    inten = torch.rand(N, S, device='cuda', generator=g) if K else None
    f = lambda: ops.composite(dens, t, dirs, far, rgb, sem, inten, 1.0, True, True)
    with torch.no_grad():
        ms = timeit(f)
    print(f'composite fwd {name}: N={N} {ms:.3f} ms  {alg * N / ms / 1e6:.0f} GB/s algorithmic = {alg * N / ms / 1e6 / peak * 100:.0f} % of {peak:.0f}')
# resampling: levels 1, 2 (in: sdist 260 + weights 256 + near/far 8 B; out: sdist + tdist)
for S_in, S_out in ((64, 64), (64, 32)):
    s = torch.sort(torch.rand(N, S_in + 1, device='cuda', generator=g), -1).values
    s[:, 0], s[:, -1] = 0, 1
    w = torch.rand(N, S_in, device='cuda', generator=g)
    w = w / w.sum(-1, keepdim=True)
    jit = torch.rand(N, 1, device='cuda', generator=g)
    f = lambda: ops.resample_level(s, w, near, far, S_out, True, 0.0103125, 0.7, jit, True)
    ms = timeit(f)
    alg = 4 * (S_in + 1) + 4 * S_in + 8 + 8 * (S_out + 1)
    print(f'resample {S_in}->{S_out}: N={N} {ms:.3f} ms  {alg * N / ms / 1e6:.0f} GB/s algorithmic = {alg * N / ms / 1e6 / peak * 100:.0f} %')
