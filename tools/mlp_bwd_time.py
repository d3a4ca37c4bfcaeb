"""Dev tool (GPU): NerfMLP data-gradient kernel alone at the bench size (median of 10, L2 flushed between launches)."""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import configs, models, ops
from nerf_lidar_b200._lib import NlbNerfMlpSaved, NlbNerfMlpGradIn, NlbNerfMlpGradOut, ptr, load, check, stream
model = models.Model(configs.nuscenes_single()); mlp = model.nerf_mlp.cuda()
N, S = 10240, 32
M = N * S
feat = torch.randn(M, 40, device='cuda') * 0.3
vd = torch.nn.functional.normalize(torch.randn(N, 3, device='cuda'), dim=-1)
density, rgb, sem, inten, saved = ops._mlp_forward_raw(mlp, feat, vd, S, True)
blob_t = ops.nerf_mlp_pack(mlp, transposed=True)
bf = lambda cols: torch.empty(M, cols, device='cuda', dtype=torch.bfloat16)
d_rgb, d_v1, d_v0, d_hs1, d_g, d_x, d_h0 = bf(16), bf(256), bf(256), bf(32), bf(128), bf(256), bf(64)
g_feat = torch.empty(M, 40, device='cuda')
gd, grgb, gsem, gint = torch.randn_like(density), torch.randn_like(rgb), torch.randn_like(sem), torch.randn_like(inten)
gin = NlbNerfMlpGradIn(ptr(gd), ptr(grgb), ptr(gsem), ptr(gint), ptr(density), ptr(rgb), ptr(sem))
sv = NlbNerfMlpSaved(*[ptr(saved[k]) for k in ('h0', 'x', 'g', 'h1', 'h2')])
gout = NlbNerfMlpGradOut(ptr(d_rgb), ptr(d_v1), ptr(d_v0), ptr(d_hs1), ptr(d_g), ptr(d_x), ptr(d_h0))
run = lambda: check(load().nlb_nerf_mlp_backward(C.byref(gin), C.byref(sv), M, ptr(blob_t), ptr(g_feat), C.byref(gout), stream()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3): run()
ts = []
for _ in range(10):
    flush.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
t = ts[len(ts) // 2]
print(f'nerf_mlp_backward {t:.4f} ms = {2 * 257000 * M / t / 1e9:.1f} TFLOP/s (257 KMAC per row)')
