"""Analysis script (CPU; lives under tests/ because it runs the oracle, which only test infrastructure may
import): how often do consecutive multisamples of an interval fall into the same grid cell?
Runs the oracle forward on the bench's synthetic rays and reports, per table and level, the share of
(interval, j >= 1) samples whose cell equals that of sample j - 1 (lane level) and the share of
(32-interval warp, j >= 1) steps where that holds for every lane (warp level) -- the upper bound of what
re-using the 8 gathered corner rows across samples can save in the fused forward kernels.
  python tests/analysis_cell_sharing.py [rays]"""
import os, sys
os.environ.setdefault('TORCHDYNAMO_DISABLE', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nerf_lidar_b200 import synthetic
from oracle import zipnerf_oracle as zo

n = int(sys.argv[1]) if len(sys.argv) > 1 else 640
sd = synthetic.init_state_dict(seed=0, table_std=0.05)
batch = synthetic.to_torch(synthetic.make_train_batch(n, seed=0))
with torch.no_grad():
    _, hist = zo.model_forward(sd, batch, None, 0.5)
for lvl, pre in enumerate(('prop_mlp_0.', 'prop_mlp_1.', 'nerf_mlp.')):
    t = hist[lvl]['tdist']
    means, stds = zo.cast_rays(t, batch['origins'], batch['directions'], batch['radii'], batch['base_x'], batch['base_y'], None)
    N, S = means.shape[:2]
    z, _ = zo.contract_mean_std(means.reshape(-1, 3), stds.reshape(-1))
    x01 = ((z / 2 + 1) / 2).reshape(N * S, 7, 3).numpy()
    gs = sd[pre + 'encoder.grid_sizes'].numpy()
    print(pre, 'rows', N * S)
    for l, g in enumerate(gs):
        scale = float(g) - 2.0          # grid_sizes = resolution + 1 = ceil(scale) + 2; scale = 2^l*H - 1
        cell = np.floor(x01 * scale + 0.5).astype(np.int64)
        same = (cell[:, 1:] == cell[:, :-1]).all(-1)            # [rows, 6]
        rows = same.shape[0] // 32 * 32
        warp = same[:rows].reshape(-1, 32, 6).all(1)
        uniq = 1 + (~same).sum(1)
        print(f'  level {l} res {int(g) - 1:5d}: lane-level same-cell {same.mean():.3f}  warp-level {warp.mean():.3f}  '
              f'cells per interval {uniq.mean():.2f}')
