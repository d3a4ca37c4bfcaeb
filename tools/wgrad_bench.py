"""Dev tool (GPU): the NerfMLP weight-gradient kernel alone on bench-sized operands (327 680 rows), CUDA-event
timed.  NLB_WGRAD_SPLIT=a,b,c,d,e,f changes the CTA shares of the six roles."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import _lib
from nerf_lidar_b200._lib import NlbNerfMlpSaved, NlbNerfMlpGradOut, NlbNerfMlpWeights, ptr, check, load, stream
M = int(sys.argv[1]) if len(sys.argv) > 1 else 327680
bf = lambda c: (torch.randn(M, c, device='cuda') * 0.1).to(torch.bfloat16)
sv = dict(h0=bf(64), x=bf(256), g=bf(128), h1=bf(256), h2=bf(256), f0=bf(64))
dcat, d_rgb, d_hs1, d_x, d_h0 = bf(640), bf(16), bf(32), bf(256), bf(64)
shapes = [(64, 40), (64,), (256, 64), (256,), (64, 256), (64,), (19, 64), (19,), (64, 256), (64,), (1, 64), (1,),
          (256, 283), (256,), (256, 539), (256,), (3, 256), (3,)]
grads = [torch.zeros(*s, device='cuda') for s in shapes]
svs = NlbNerfMlpSaved(*[ptr(sv[k]) for k in ('h0', 'x', 'g', 'h1', 'h2', 'f0')])
go = NlbNerfMlpGradOut(ptr(d_rgb), dcat[:, 384:].data_ptr(), dcat[:, 128:384].data_ptr(), ptr(d_hs1), dcat.data_ptr(), ptr(d_x), ptr(d_h0), 640, 640, 640)
wg = NlbNerfMlpWeights(*[ptr(t) for t in grads])
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(load().nlb_nerf_mlp_wgrad(C.byref(svs), C.byref(go), M, C.byref(wg), stream()))
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(os.environ.get('NLB_WGRAD_SPLIT', 'default'), 'ms:', ' '.join(f'{t:.3f}' for t in ts))
# check one product against torch
want = (dcat[:, 128:384].float().t() @ sv['x'].float()) * len(ts)
print('W_v0 rel err', float((grads[12][:, :256] - want).norm() / want.norm()))
