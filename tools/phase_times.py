"""Dev tool (GPU): CUDA-event time of the phases of one training step."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import configs, models, synthetic, train
cfg = configs.nuscenes_single()
model = models.Model(cfg, training=True).cuda()
model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=0).items()}, strict=False)
tr = train.Trainer(model, cfg)
batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(8192, seed=1)).items()}
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
acc = {}
for i in range(13):
    step = 6000 + i
    train_frac = float(np.clip((step - 1) / (cfg.max_steps - 1), 0, 1))
    e0 = ev()
    rend, hist = model(True, batch, train_frac, True)
    e1 = ev()
    losses = train.compute_losses(batch, rend, hist, cfg, step, 2)
    loss = sum(v for k, v in losses.items() if k != 'hash_decay')
    e2 = ev()
    loss.backward()
    e3 = ev()
    tr.optimizer_step(step)
    e4 = ev()
    torch.cuda.synchronize()
    if i >= 3:
        for k, a, b in (('forward', e0, e1), ('losses', e1, e2), ('backward', e2, e3), ('optimizer', e3, e4)):
            acc[k] = acc.get(k, 0) + a.elapsed_time(b) / 10
print({k: round(v, 3) for k, v in acc.items()}, 'total', round(sum(acc.values()), 3))
