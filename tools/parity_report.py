"""Per-key error of the free-running Model.forward against the reference golden outputs (max |d| / max |ref|),
for the fp32-head and the bf16 tensor-core paths: the numbers the bars of tests/test_gpu_model.py are set from."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from tests.helpers import CASES, load_case, use_torch_heads  # noqa: E402
from nerf_lidar_b200 import configs, models  # noqa: E402

for name in CASES:
    for dtype in (torch.float32, torch.bfloat16):
        case, golden, sd, batch, rin = load_case(name, 'cuda')
        model = models.Model(configs.nuscenes_single()).cuda()
        model.load_state_dict(sd, strict=False)
        model.eval(); model.training = False
        if dtype == torch.float32:
            use_torch_heads(model)
        with torch.no_grad():
            rend, hist = model(case['rand'], batch, case['train_frac'], True, rand_inputs=rin)
        print(f'== {name} {dtype}')
        for key, ref in golden.items():
            kind, k = key.split('_', 1)
            i = int(kind[-1])
            src = hist[i] if kind.startswith('hist') else rend[i]
            got = src[k].float().cpu().numpy().reshape(ref.shape)
            scale = np.abs(ref).max() + 1e-30
            d = np.abs(got - ref)
            print(f'  {key:32s} max {d.max() / scale:.3e}  median {np.median(d) / scale:.3e}  scale {scale:.3e}')
