"""Dev tool (GPU): cumulative per-level cost of the fused encode forward / backward
kernels on the bench workload's real sample distributions (levels 0..L'-1)."""
import os, sys, math, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import configs, models, synthetic, ops, _lib
from nerf_lidar_b200._lib import NlbTable, ptr, load, check, stream

cfg = configs.nuscenes_single()
model = models.Model(cfg, training=True).cuda()
model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=0).items()}, strict=False)
batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(8192, seed=1)).items()}
with torch.no_grad():
    rend, hist = model(True, batch, 0.25, True)
rays = ops.RayBundle(batch)
encs = [model.prop_mlp_0.encoder, model.prop_mlp_1.encoder, model.nerf_mlp.encoder]


def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for li, enc in enumerate(encs):
    tdist = hist[li]['tdist'].contiguous()
    S = tdist.shape[1] - 1
    deg = torch.rand(rays.N, S, 7, device='cuda')
    rows = rays.N * S
    Cc = enc.level_dim
    grad = torch.zeros_like(enc.embeddings)
    print(f'== table {li}: L={enc.num_levels} C={Cc} S={S} rows={rows} offsets={enc.offsets.tolist()}')
    prev_f = prev_b = 0.
    only_full = len(sys.argv) > 1 and sys.argv[1] == 'full'
    for L in ([enc.num_levels] if only_full else range(1, enc.num_levels + 1)):
        tab = NlbTable(ptr(enc.embeddings), ptr(enc.offsets), ptr(enc.grid_sizes), L, Cc, int(enc.base_resolution),
                       float(math.log2(enc.per_level_scale)), ops.host_offsets(enc))
        feats = torch.empty(rows, L * Cc, device='cuda')
        g = torch.randn(rows, L * Cc, device='cuda')
        rd = rays.desc(tdist, deg, 0.35)
        f = timeit(lambda: check(load().nlb_encode_forward(C.byref(rd), C.byref(tab), ptr(feats), stream())))
        wsn = torch.empty(max(load().nlb_encode_backward_workspace_bytes(C.byref(tab)) // 4, 1), device='cuda')
        b = timeit(lambda: check(load().nlb_encode_backward(C.byref(rd), C.byref(tab), ptr(g), ptr(grad), ptr(wsn), stream())))
        print(f'  L={L:2d} fwd {f:7.3f} ms (+{f - prev_f:6.3f})   bwd {b:7.3f} ms (+{b - prev_b:6.3f})')
        prev_f, prev_b = f, b
