"""Dev tool (GPU, one device): gradients of two half batches accumulated in one trainer against the sum of the
two halves computed separately, and run-to-run repeatability of one half -- at the initial state and after
optimizer steps (isolates compute nondeterminism from the data-parallel exchange)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import configs, models, synthetic, train
cfg = configs.nuscenes_single()
dev = torch.device('cuda', 0)
sd = {k: v.to(dev) for k, v in synthetic.init_state_dict(seed=5, table_std=0.1).items()}
m = models.Model(cfg, training=True).to(dev); m.load_state_dict(sd, strict=False)
tr = train.Trainer(m, cfg)
B = int(os.environ.get('GC_B', '1024'))
halves = [{k: v.to(dev) for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=80 + r)).items()} for r in range(2)]
rins = [[{k: torch.from_numpy(v).to(dev) for k, v in x.items()} for x in synthetic.make_rand_inputs(halves[r]['origins'].shape[0], seed=90 + r)] for r in range(2)]
rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))


def grads(which, step):
    tr.flat_grad.zero_()
    for t in tr.tables: t['grad'].zero_()
    for r in which:
        _, ml, pl = tr.forward_losses(halves[r], step, 0, rins[r])
        (ml + pl).backward()
    torch.cuda.synchronize()
    return [tr.flat_grad.clone()] + [t['grad'].clone() for t in tr.tables]


names = ['dense'] + [t['name'].split('.')[0] for t in tr.tables]
for i in range(3):
    step = 6000 + 500 * i
    both = grads([0, 1], step)
    g0, g0b, g1 = grads([0], step), grads([0], step), grads([1], step)
    for n, a, x, xb, y in zip(names, both, g0, g0b, g1):
        print(f'step {i} {n}: accumulated vs separate {rel(x + y, a):.2e}   repeat of half 0 {rel(xb, x):.2e}')
    # one optimizer step from the accumulated gradients
    for dst, src in zip([tr.flat_grad] + [t['grad'] for t in tr.tables], both):
        dst.copy_(src * 0.5)
    tr.optimizer_step(step)
