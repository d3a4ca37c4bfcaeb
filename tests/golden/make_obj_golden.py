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
    car = np.stack([0.15 + 0.25 * ts, 0.02 * ts, np.zeros(T), 0.3 + 0.2 * ts, np.full(T, 0.45), np.full(T, 0.30),
                    np.full(T, 0.30), ts, np.zeros(T)], -1)
    truck = np.stack([-0.45 + 0.1 * ts, 0.05 + 0.0 * ts, np.full(T, 0.01), -0.4 + 0.1 * ts, np.full(T, 0.55),
                      np.full(T, 0.35), np.full(T, 0.40), ts, np.ones(T)], -1)
    return {0: car.astype(np.float64), 1: truck.astype(np.float64)}, {0: 'vehicle.car', 1: 'vehicle.truck'}


def object_table(shape, gen):
    """Visible object tables (the reference initialises them at 1e-4): U(-0.5, 0.5), drawn in module order
    (obj_mlp_13, obj_mlp_14) from ONE generator seeded with 4."""
    return (torch.rand(shape, generator=gen) * 2 - 1) * 0.5


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


def loss_coefficients(N, S):
    g = torch.Generator().manual_seed(7)
    return torch.randn(N, S, generator=g), torch.randn(N, S, 3, generator=g)


def branch_gradients(models, model, batch, tdist):
    """Gradients of the final-level object branch: the per-track loop of Z/internal/models.py:401-472 executed with
    the reference's own obj_utils.box_pts and ObjMLP modules on the golden sample distances, a fixed linear
    functional of the merged density / rgb as the loss.  Dense parameters are stored whole, the two 70 MB table
    gradients as their non-zero rows."""
    ou = models.obj_utils
    # track refinement (Z/train.py:244-257): the refined track is a function of Track_opt's corrections, so the
    # gradient reaches get_pose's blend and box_pts' translation / yaw -- kept here w.r.t. the interpolated pose
    # and w.r.t. the track table itself
    tracks_leaf = model.tracks.detach().clone().requires_grad_(True)
    obj_pose = ou.get_pose(batch['timestamp'], tracks_leaf)
    obj_pose.retain_grad()
    t_mids = 0.5 * (tdist[..., :-1] + tdist[..., 1:])
    pts_w = t_mids[..., None] * batch['directions'][:, None, :] + batch['origins'][:, None, :]
    pts_o, viewdirs_o, imap = ou.box_pts(pts=pts_w, viewdirs=batch['viewdirs'], obj_pose=obj_pose, sym=False)
    N, S = t_mids.shape
    dens, rgbm = torch.zeros(N, S), torch.zeros(N, S, 3)
    model.zero_grad()
    for track_id in range(len(model.bboxes[0].keys())):
        idx = imap[:, :, track_id]
        if idx.sum() == 0:
            continue
        pts_k = pts_o[idx][:, track_id, :]
        stds = torch.zeros_like(pts_k)[..., 0]
        vd_k = viewdirs_o[idx][:, track_id, :]
        class_id = ou.query_class(model.obj_type_info[track_id])
        obj_mlp = model.get_submodule(f'obj_mlp_{class_id}')
        latent = model.latent_vector_dict[f'obj_latent_{track_id}'][None, ...].repeat(pts_k.shape[0], 1)
        r = obj_mlp(False, pts_k, stds, viewdirs=vd_k, latent=latent, glo_vec=None, exposure=None)
        for key, cur in (('density', dens), ('rgb', rgbm)):
            tmp = torch.zeros_like(cur)
            tmp[idx] = r[key]
            m = idx if idx.shape == cur.shape else idx[..., None].expand(cur.shape)
            if key == 'density':
                dens = torch.where(m, tmp, cur)
            else:
                rgbm = torch.where(m, tmp, cur)
    a, b = loss_coefficients(N, S)
    ((dens * a).sum() + (rgbm * b).sum()).backward()
    out = {'grad_pose': obj_pose.grad.numpy().copy(), 'grad_tracks': tracks_leaf.grad.numpy().copy()}
    for name, p in model.named_parameters():
        if not (name.startswith('obj_mlp') or name.startswith('latent_vector_dict')) or p.grad is None:
            continue
        if name.endswith('encoder.embeddings'):
            rows = torch.nonzero(p.grad.abs().sum(-1) > 0).reshape(-1)
            out['grad_rows::' + name] = rows.numpy().astype(np.int64)
            out['grad_vals::' + name] = p.grad[rows].numpy()
        else:
            out['grad::' + name] = p.grad.numpy().copy()
    model.zero_grad()
    return out


def run_reference():
    models = ref_shims.import_reference()
    ref_shims.apply_gin_bindings(models)
    apply_obj_bindings(models)
    torch.manual_seed(0)
    obj_info, obj_type = tracks()
    g = torch.Generator().manual_seed(3)
    latents = {f'obj_latent_{i}': torch.nn.Parameter(torch.randn(128, generator=g)) for i in obj_info}
    tg = torch.Generator().manual_seed(4)      # object tables: regenerated by the tests from this seed
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
                    p.copy_(object_table(tuple(p.shape), tg))
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
    out.update(branch_gradients(models, model, batch, hist[2]['tdist']))
    for k, v in osd.items():
        if not k.endswith('encoder.embeddings'):       # the tables are regenerated from the seed (30 MB each)
            out['sd::' + k] = v.numpy()
        else:
            out['tabsum::' + k] = np.array([float(v.double().sum()), float(v.double().abs().sum())])
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
