"""Dev tool (GPU): the three-graph data-parallel step on ONE GPU without a process group (the collectives are
no-ops), eager step first -- isolates capture problems of Trainer.train_step_graphed from NCCL."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import configs, models, synthetic, train
cfg = configs.nuscenes_single()
dev = torch.device('cuda', 0)
sd = {k: v.to(dev) for k, v in synthetic.init_state_dict(seed=5, table_std=0.1).items()}
m = models.Model(cfg, training=True).to(dev); m.load_state_dict(sd, strict=False)
world = int(os.environ.get('PROBE_WORLD', '2'))
tr = train.Trainer(m, cfg, world=world, rank=0)
batch = {k: v.to(dev) for k, v in synthetic.to_torch(synthetic.make_train_batch(1024, seed=80)).items()}
n = batch['origins'].shape[0]
rin = [{k: torch.from_numpy(v).to(dev) for k, v in x.items()} for x in synthetic.make_rand_inputs(n, seed=90)]
if os.environ.get('PROBE_EAGER_FIRST', '1') == '1':
    print('eager', float(tr.train_step(batch, 6000, 0, rin)['loss']))
for i in range(3):
    out = tr.train_step_graphed(batch, 6500 + i, 0, rin)
    torch.cuda.synchronize()
    print('graphed', i, float(out['loss']))
print('probe ok')
