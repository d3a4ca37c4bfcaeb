"""Dev tool (GPU): NerfMLP forward alone at the bench size -- time per launch (training mode with saved
activations, and inference mode), cycles per tile and SM clock of block 0 (nlb_debug_set_timeline slots 120-124).
NLB_MLP_FWD_LEGACY=1 selects the one-tile kernel for A/B."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import _lib, configs, models, ops
model = models.Model(configs.nuscenes_single()); mlp = model.nerf_mlp.cuda()
N, S = 10240, 32
feat = torch.randn(N * S, 40, device='cuda') * 0.5
vd = torch.nn.functional.normalize(torch.randn(N, 3, device='cuda'), dim=-1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


t_train = timed(lambda: ops._mlp_forward_raw(mlp, feat, vd, S, True))
t_inf = timed(lambda: ops._mlp_forward_raw(mlp, feat, vd, S, False))
flops = 528384 * N * S
print(f'legacy={os.environ.get("NLB_MLP_FWD_LEGACY", "0")} train {t_train:.4f} ms = {flops / t_train / 1e9:.1f} TFLOP/s   '
      f'inference {t_inf:.4f} ms = {flops / t_inf / 1e9:.1f} TFLOP/s')
buf = torch.zeros(128, dtype=torch.int64, device='cuda')
_lib.check(_lib.load().nlb_debug_set_timeline(buf.data_ptr()))
ops._mlp_forward_raw(mlp, feat, vd, S, True); torch.cuda.synchronize()
t = buf.cpu().tolist()
if t[124] > 0:
    cyc, ns = t[121] - t[120], t[123] - t[122]
    print(f'block 0: {t[124]} tiles, {cyc} cycles = {cyc / t[124]:.0f} per tile, {ns} ns -> SM clock {cyc / ns * 1000:.0f} MHz')

if t[124] > 0 and t[0] > 0:
    t0 = min(x for x in t[:120] if x > 0)
    rel = lambda i: (t[i] - t0) if t[i] > 0 else None
    mm = ['x_rdy', 'HS0 iss', 'V0 iss', 'g_rdy', 'HS1 iss', 'h1_rdy', 'hs1_free', 'V1 iss', 'f_rdy', 'L0 iss', 'h0_rdy', 'L1 iss', 'h2_rdy', 'RGB iss']
    fr = ['top', 'inputs', 'V1 done', 'f sig', 'L0 acc', 'h0 sig', 'L1 acc', 'x sig', 'x saved', 'HS1 acc', 'end']
    ma = ['HS0 acc', 'g sig', 'HS1 acc', 'V0 acc', 'h1 sig', 'h1 saved', 'V1 acc', 'h2 sig', 'h2 saved', 'RGB acc', 'end']
    for j in range(2):
        print('MMA  it', 2 + j, [(n, rel(j * 32 + i)) for i, n in enumerate(mm)])
        print('front k', 2 + j, [(n, rel(64 + j * 16 + i)) for i, n in enumerate(fr)])
        print('main it', 2 + j, [(n, rel(96 + j * 12 + i)) for i, n in enumerate(ma)])
_lib.check(_lib.load().nlb_debug_set_timeline(0))
