"""Dev tool (GPU): N training steps of the bench workload, for ncu captures."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import configs, models, synthetic, train
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = configs.nuscenes_single()
model = models.Model(cfg, training=True).cuda()
model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=0).items()}, strict=False)
tr = train.Trainer(model, cfg)
batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(8192, seed=1)).items()}
for i in range(steps):
    out = tr.train_step(batch, 6000 + i, 2)
torch.cuda.synchronize()
print('ok', float(out['loss']))
