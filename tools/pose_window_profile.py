"""Where the pose-refinement window's extra time goes: per-ABI-call device time of the eager step in both
regimes, the graphed step time of both, and the count of torch kernels (profiler) in the window step."""
import sys
import torch
sys.path.insert(0, '.')
from nerf_lidar_b200 import _lib, configs, models, synthetic, train, posenet

dev = torch.device('cuda')
cfg = configs.nuscenes_single(use_intensity=True)
model = models.Model(cfg, training=True).to(dev)
model.load_state_dict({k: v.to(dev) for k, v in synthetic.init_state_dict(seed=0, table_std=1e-4).items()}, strict=False)
tr = train.Trainer(model, cfg)
B = 8192
b = synthetic.make_train_batch(B, seed=1)
b['glo_idx'] = synthetic.sensor_index(b)
cb = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
num_patch = (B // 4) // 1024
net, opt, lr_fn = posenet.create_posenet(1, cfg, num_lidars=1, device=dev)


def run(fn, step, n):
    for i in range(3):
        fn(cb, step + i, num_patch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(cb, step + 10 + i, num_patch)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
for regime, step in (('steady', 6000), ('window', 1000)):
    tr.attach_posenet(*((net, opt, lr_fn) if regime == 'window' else (None, None, None)))
    res[regime, 'graph'] = run(tr.train_step_graphed, step, 10)
    _lib.TIMER = _lib.KernelTimer()
    res[regime, 'eager'] = run(tr.train_step, step + 100, 3)
    kt = _lib.TIMER.summary()
    _lib.TIMER = None
    res[regime, 'abi_ms'] = sum(v[1] for v in kt.values()) / 6     # 3 warm-up + 3 timed steps were all recorded
    res[regime, 'kt'] = {k: round(v[1] / 6, 4) for k, v in kt.items()}
    print(regime, 'graph %.3f ms  eager %.3f ms  sum of ABI calls %.3f ms' % (res[regime, 'graph'], res[regime, 'eager'], res[regime, 'abi_ms']), flush=True)
diff = {k: round(res['window', 'kt'].get(k, 0) - res['steady', 'kt'].get(k, 0), 4) for k in res['window', 'kt']}
print('per-call difference (ms per step):', {k: v for k, v in diff.items() if abs(v) > 0.01})
tr.attach_posenet(net, opt, lr_fn)
from torch.profiler import profile, ProfilerActivity
tr.train_step(cb, 1200, num_patch)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_step(cb, 1201, num_patch)
    torch.cuda.synchronize()
rows = sorted(((e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0 and e.cpu_time_total == 0),
              key=lambda r: -r[2])
print('device kernels in one window step: %d launches, %.3f ms' % (sum(r[1] for r in rows), sum(r[2] for r in rows)))
for r in rows[:28]:
    print('  %-90s x%-3d %.3f ms' % (r[0][:90], r[1], r[2]))
