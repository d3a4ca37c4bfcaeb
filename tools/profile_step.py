"""Dev tool (GPU): torch.profiler breakdown of one training step."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import configs, models, synthetic, train
from torch.profiler import profile, ProfilerActivity
cfg = configs.nuscenes_single()
model = models.Model(cfg, training=True).cuda()
model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=0).items()}, strict=False)
tr = train.Trainer(model, cfg)
B = 8192
batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=1)).items()}
for i in range(3):
    tr.train_step(batch, 6000 + i, 2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        tr.train_step(batch, 6010 + i, 2)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70))
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=25, max_name_column_width=70))
