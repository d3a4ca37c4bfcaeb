"""Generates tests/golden/obj_tracks.npz: the reference's OWN Model.forward with the dynamic-object branch
(Config.instance_obj=True, Z/internal/models.py:306-315,401-477; obj_utils.get_pose / box_pts; per-class ObjMLP
with split shape / texture latents as configs/nuscenes_single.gin binds it) on seeded rays, two synthetic tracks
(a car and a truck, moving, yawed) and random-init weights.  The ObjMLP weights are the reference modules' own
torch-seeded initialisation and are stored in the fixture together with the outputs.
  TORCHDYNAMO_DISABLE=1 python tests/golden/make_obj_golden.py [--check]"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from nerf_lidar_b200 import synthetic  # noqa: E402

OUT = os.path.join(HERE, 'obj_tracks.npz')
SEED, B = 21, 192


def tracks():
    """[n_obj][T, 9] = centre(3), yaw about z, wlh(3), timestamp, track id (datasets.py:1442-1452)."""
    T = 6
    ts = np.linspace(-1.0, 1.0, T)
    car = np.stack([np.stack([0.15 + 0.25 * ts, 0.02 * ts, np.zeros(T)], -1)[:, 0], 0.02 * ts, np.zeros(T),
                    0.3 + 0.2 * ts, np.full(T, 0.45), np.full(T, 0.30), np.full(T, 0.30), ts, np.zeros(T)], -1)
    truck = np.stack([-0.45 + 0.1 * ts, 0.05 + 0.0 * ts, np.full(T, 0.01), -0.4 + 0.1 * ts, np.full(T, 0.55),
                      np.full(T, 0.35), np.full(T, 0.40), ts, np.ones(T)], -1)
    return {0: car.astype(np.float64), 1: truck.astype(np.float64)}, {0: 'vehicle.car', 1: 'vehicle.truck'}


def make_batch():
    b = synthetic.make_train_batch(B, seed=SEED)
    rng = np.random.default_rng(SEED + 1)
    b['timestamp'] = rng.uniform(-0.9, 0.9, b['timestamp'].shape).astype(np.float32)
    return b


class ObjConfig(ref_shims.RefConfig):
    instance_obj = True
    latent_size = 128
    use_intensity = False      # ObjMLP has no intensity head: the reference's merge loop cannot overwrite one


def apply_obj_bindings(models):
    """configs/nuscenes_single.gin:36-44."""
    O = models.ObjMLP
    O.disable_rgb = False
    O.grid_disired_resolution = 1024
    O.density_init = True
    O.disable_density_normals = True
    O.obj_mode = False
    O.bottleneck_width = 64
    O.grid_level_dim = 2
    O.net_width_viewdirs = 32
    O.split_latent = True


def run_reference():
    models = ref_shims.import_reference()
    ref_shims.apply_gin_bindings(models)
    apply_obj_bindings(models)
    torch.manual_seed(0)
    obj_info, obj_type = tracks()
    g = torch.Generator().manual_seed(3)
    latents = {f'obj_latent_{i}': torch.nn.Parameter(torch.randn(128, generator=g)) for i in obj_info}
    model = models.Model(config=ObjConfig(), bboxes=(obj_info, obj_type), latent_vector_dict=latents)
    sd = synthetic.init_state_dict(seed=SEED, table_std=0.3, use_intensity=False)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # visible object tables (the reference initialises them at 1e-4)
    osd = {}
    for name, p in model.named_parameters():
        if name.startswith('obj_mlp') or name.startswith('latent_vector_dict'):
            if name.endswith('encoder.embeddings'):
                with torch.no_grad():
                    p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.5)
            osd[name] = p.detach().clone()
    batch = synthetic.to_torch(make_batch())
    model.eval()
    model.training = False
    with torch.no_grad():
        rend, hist = model(False, batch, 1.0, True)
    out = {}
    for i, h in enumerate(hist):
        for k in ('density', 'rgb', 'semantic', 'weights', 'tdist', 'obj_mask'):
            if h.get(k) is not None:
                out[f'hist{i}_{k}'] = h[k].numpy().astype(np.float32)
    for i, r in enumerate(rend):
        for k in ('rgb', 'depth', 'semantic', 'acc', 'obj_mask', 'instance_mask'):
            if k in r:
                out[f'rend{i}_{k}'] = r[k].numpy().astype(np.float32)
    out['pose'] = models.obj_utils.get_pose(batch['timestamp'], model.tracks).numpy().astype(np.float32)
    for k, v in osd.items():
        if not k.endswith('encoder.embeddings'):       # the tables are regenerated from the seed (60 MB each)
            out['sd::' + k] = v.numpy()
    return out, model


if __name__ == '__main__':
    got, model = run_reference()
    if '--check' in sys.argv:
        gold = np.load(OUT)
        for k in gold.files:
            assert np.array_equal(gold[k], got[k]), k
        print('ok')
    else:
        np.savez_compressed(OUT, **got)
        print('wrote', OUT, os.path.getsize(OUT) // 1024, 'KiB')
        for i in range(3):
            print('level', i, 'object samples', int(got[f'hist{i}_obj_mask'].sum()), 'of', got[f'hist{i}_obj_mask'].size)
        print([k for k in got if k.startswith('sd::')])
