"""Two small eager training steps (and one render chunk) for compute-sanitizer runs:
  PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool initcheck python tools/sanitize_step.py
  compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import sys

import torch

sys.path.insert(0, '.')
from nerf_lidar_b200 import configs, models, synthetic, train  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = configs.nuscenes_single()
model = models.Model(cfg, training=True).cuda()
model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=5, table_std=0.1).items()}, strict=False)
tr = train.Trainer(model, cfg)
batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=50)).items()}
for i in range(2):
    out = tr.train_step(batch, 6000 + i, 0)
torch.cuda.synchronize()
print({k: float(v) for k, v in out.items()})
with torch.no_grad():
    model.eval(); model.training = False
    r, _ = model(False, batch, 1.0, True)
torch.cuda.synchronize()
print('ok', float(r[-1]['depth'].mean()))
