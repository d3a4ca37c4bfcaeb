"""Generates tests/golden/*.npz by executing the REFERENCE's own Python
(/root/reference/NeRF_LiDAR/zipnerf/internal/models.py Model.forward, unmodified)
on CPU in the build container.  The reference tree does not travel to the GPU
box, so the outputs are committed as small fixtures; inputs and weights are
regenerated from seeds by nerf_lidar_b200.synthetic on both sides.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from nerf_lidar_b200 import synthetic  # noqa: E402

CASES = {
    # name: (batch_size, seed, table_std, rand, train_frac)
    'eval_init': dict(batch_size=64, seed=11, table_std=1e-4, rand=False, train_frac=1.0),
    'train_visible': dict(batch_size=64, seed=12, table_std=0.5, rand=True, train_frac=0.5),
}

HIST_KEYS = ('density', 'rgb', 'semantic', 'intensity', 'sdist', 'weights', 'tdist')
REND_KEYS = ('rgb', 'depth', 'semantic', 'intensity', 'acc', 'distance_mean', 'distance_median',
             'distance_percentile_5', 'distance_percentile_95')


def run_reference(case):
    models = ref_shims.import_reference()
    ref_shims.apply_gin_bindings(models)
    torch.manual_seed(0)
    model = models.Model(config=ref_shims.RefConfig())
    sd = synthetic.init_state_dict(seed=case['seed'], table_std=case['table_std'])
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith('encoder.idx') for k in missing), missing
    batch = synthetic.to_torch(synthetic.make_train_batch(case['batch_size'], seed=case['seed']))
    n = batch['origins'].shape[0]
    model.eval()
    model.training = False
    if case['rand']:
        rin = synthetic.make_rand_inputs(n, seed=case['seed'])
        queue = []
        for r in rin:
            queue += [torch.from_numpy(r['jitter']), torch.from_numpy(r['deg'])]
        with ref_shims.injected_rand(queue) as q:
            with torch.no_grad():
                rend, hist = model(True, batch, case['train_frac'], True)
        assert not q, 'reference drew fewer random tensors than expected'
    else:
        with torch.no_grad():
            rend, hist = model(False, batch, case['train_frac'], True)
    return rend, hist


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, case in CASES.items():
        rend, hist = run_reference(case)
        blob = {}
        for i, h in enumerate(hist):
            for k in HIST_KEYS:
                if h.get(k) is not None:
                    blob[f'hist{i}_{k}'] = h[k].numpy().astype(np.float32)
        for i, r in enumerate(rend):
            for k in REND_KEYS:
                if k in r:
                    blob[f'rend{i}_{k}'] = r[k].numpy().astype(np.float32)
        path = os.path.join(out_dir, f'{name}.npz')
        np.savez_compressed(path, **blob)
        print(name, 'wrote', path, os.path.getsize(path) // 1024, 'KiB', len(blob), 'arrays')


if __name__ == '__main__':
    main()
