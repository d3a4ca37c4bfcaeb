"""Dev tool (GPU): back-to-back CUDA-event times of the encode / proposal kernels on
the bench workload's real sample distributions (20 iterations each)."""
import os, sys, math, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import configs, models, synthetic, ops, _lib
from nerf_lidar_b200._lib import ptr, load, check, stream

cfg = configs.nuscenes_single()
model = models.Model(cfg, training=True).cuda()
model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=0).items()}, strict=False)
batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(8192, seed=1)).items()}
with torch.no_grad():
    rend, hist = model(True, batch, 0.25, True)
rays = ops.RayBundle(batch)


def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out = []
for li, mlp in enumerate([model.prop_mlp_0, model.prop_mlp_1]):
    enc = mlp.encoder
    tdist = hist[li]['tdist'].contiguous()
    S = tdist.shape[1] - 1
    deg = torch.rand(rays.N, S, 7, device='cuda')
    rows = rays.N * S
    L = enc.num_levels
    tab = ops._table_desc(enc, enc.embeddings)
    rd = rays.desc(tdist, deg, 0.35)
    l0, l2 = mlp.density_layer[0], mlp.density_layer[2]
    W0, b0, W1, b1 = l0.weight.detach().contiguous(), l0.bias.detach().contiguous(), l2.weight.detach().reshape(-1).contiguous(), l2.bias.detach().contiguous()
    dens = torch.empty(rays.N, S, device='cuda')
    feats = torch.empty(rows, L, device='cuda')
    f = timeit(lambda: check(load().nlb_prop_forward(C.byref(rd), C.byref(tab), ptr(W0), ptr(b0), ptr(W1), ptr(b1), ptr(dens), ptr(feats), stream())))
    gd = torch.randn(rays.N, S, device='cuda') * 1e-3
    gt = torch.zeros_like(enc.embeddings)
    gW0, gb0, gW1, gb1 = torch.zeros_like(W0), torch.zeros_like(b0), torch.zeros_like(W1), torch.zeros_like(b1)
    ws = torch.empty(load().nlb_prop_backward_workspace_bytes(rays.N, S, C.byref(tab)) // 4, device='cuda')
    b = timeit(lambda: check(load().nlb_prop_backward(C.byref(rd), C.byref(tab), ptr(W0), ptr(b0), ptr(W1), ptr(b1), ptr(feats), ptr(gd), ptr(gt), ptr(gW0), ptr(gb0), ptr(gW1), ptr(gb1), ptr(ws), stream())))
    out.append(f'prop{L}: fwd {f:.3f} bwd {b:.3f}')
enc = model.nerf_mlp.encoder
tdist = hist[2]['tdist'].contiguous()
S = tdist.shape[1] - 1
deg = torch.rand(rays.N, S, 7, device='cuda')
rows = rays.N * S
tab = ops._table_desc(enc, enc.embeddings)
rd = rays.desc(tdist, deg, 0.35)
feats = torch.empty(rows, 40, device='cuda')
g = torch.randn(rows, 40, device='cuda')
gt = torch.zeros_like(enc.embeddings)
f = timeit(lambda: check(load().nlb_encode_forward(C.byref(rd), C.byref(tab), ptr(feats), stream())))
wsn = torch.empty(max(load().nlb_encode_backward_workspace_bytes(C.byref(tab)) // 4, 1), device='cuda')
b = timeit(lambda: check(load().nlb_encode_backward(C.byref(rd), C.byref(tab), ptr(g), ptr(gt), ptr(wsn), stream())))
out.append(f'nerf: fwd {f:.3f} bwd {b:.3f}')
print(os.environ.get('NLB_LIB', 'default'), ' | '.join(out))

# ---- NerfMLP fused kernels (training forward with saved activations, data-gradient backward)
mlp = model.nerf_mlp
S = 32
M = rays.N * S
feat = (torch.randn(M, 40, device='cuda') * 0.3)
vd = ops.f32(batch['viewdirs'])
import ctypes as C2
from nerf_lidar_b200._lib import NlbNerfMlpSaved, NlbNerfMlpGradIn, NlbNerfMlpGradOut
density, rgb, sem, inten, saved = ops._mlp_forward_raw(mlp, feat, vd, S, True)
tf = timeit(lambda: ops._mlp_forward_raw(mlp, feat, vd, S, True))
ti = timeit(lambda: ops._mlp_forward_raw(mlp, feat, vd, S, False))
blob_t = ops.nerf_mlp_pack(mlp, transposed=True)
bf = lambda cols: torch.empty(M, cols, device='cuda', dtype=torch.bfloat16)
d_rgb, d_v1, d_v0, d_hs1, d_g, d_x, d_h0 = bf(16), bf(256), bf(256), bf(32), bf(128), bf(256), bf(64)
g_feat = torch.empty(M, 40, device='cuda')
gd, grgb, gsem, gint = torch.randn_like(density), torch.randn_like(rgb), torch.randn_like(sem), torch.randn_like(inten)
gin = NlbNerfMlpGradIn(ptr(gd), ptr(grgb), ptr(gsem), ptr(gint), ptr(density), ptr(rgb), ptr(sem))
sv = NlbNerfMlpSaved(*[ptr(saved[k]) for k in ('h0', 'x', 'g', 'h1', 'h2')])
gout = NlbNerfMlpGradOut(ptr(d_rgb), ptr(d_v1), ptr(d_v0), ptr(d_hs1), ptr(d_g), ptr(d_x), ptr(d_h0))
tb = timeit(lambda: check(load().nlb_nerf_mlp_backward(C.byref(gin), C.byref(sv), M, ptr(blob_t), ptr(g_feat), C.byref(gout), stream())))
flop = 2 * 264192 * M
print(f'mlp: fwd(train) {tf:.3f} ms = {flop / tf / 1e9:.0f} TFLOP/s | fwd(infer) {ti:.3f} ms = {flop / ti / 1e9:.0f} TFLOP/s | bwd {tb:.3f} ms = {2 * flop / tb / 1e9:.0f} TFLOP/s (data-gradient chain, 2x fwd FLOPs nominal)')
