import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import _lib, configs, models, ops
model = models.Model(configs.nuscenes_single()); mlp = model.nerf_mlp.cuda()
N, S = 10240, 32
feat = torch.randn(N * S, 40, device='cuda') * 0.5
vd = torch.nn.functional.normalize(torch.randn(N, 3, device='cuda'), dim=-1)
for _ in range(3): ops.nerf_mlp_forward(mlp, feat, vd, S)
buf = torch.zeros(128, dtype=torch.int64, device='cuda')
_lib.check(_lib.load().nlb_debug_set_timeline(buf.data_ptr()))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); ops.nerf_mlp_forward(mlp, feat, vd, S); e1.record(); torch.cuda.synchronize()
print('kernel ms', e0.elapsed_time(e1))
t = buf.cpu().tolist(); t0 = min(x for x in t[:100] if x > 0)
names_m = ['start','F rdy','L0 iss','H0 rdy','L1 iss','X rdy','HS0 iss','V0 iss','G rdy','HS1 iss','H1 rdy','V1 iss','H2 rdy','RGB iss']
for tile in range(2):
    print('tile', tile, 'MMA thread:', [(n, t[tile*16+i]-t0) for i, n in enumerate(names_m)])
    en = ['L0','L1','HS0','HS1','V0','V1','RGB']
    print('tile', tile, 'EPI thread:', [(en[i], t[64+tile*16+2*i]-t0, t[64+tile*16+2*i+1]-t0) for i in range(7)], 'F staged', t[64+tile*16+14]-t0)
print('tile 1 per layer (wait for weights, issue+commit) cycles:', [(n, t[100+2*i], t[101+2*i]) for i, n in enumerate(['L0','L1','HS0','V0','HS1','V1','RGB'])])
