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
    ws = torch.empty(load().nlb_prop_backward_workspace_bytes(rays.N, S, L) // 4, device='cuda')
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
b = timeit(lambda: check(load().nlb_encode_backward(C.byref(rd), C.byref(tab), ptr(g), ptr(gt), stream())))
out.append(f'nerf: fwd {f:.3f} bwd {b:.3f}')
print(os.environ.get('NLB_LIB', 'default'), ' | '.join(out))
