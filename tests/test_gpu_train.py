"""Training step around the hot path: loss values and gradients of the B200 path
against autograd through the oracle, and the fused hash-decay + Adam kernel
against torch.optim.Adam."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import train_oracle as to
from oracle import zipnerf_oracle as zo
from nerf_lidar_b200 import synthetic
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def test_losses_and_gradients_vs_oracle():
    from nerf_lidar_b200 import configs, models, train
    B = 2048
    sd = synthetic.init_state_dict(seed=21, table_std=0.2)
    batch = synthetic.to_torch(synthetic.make_train_batch(B, seed=21))
    n = batch['origins'].shape[0]
    rin = [{k: torch.from_numpy(v) for k, v in r.items()} for r in synthetic.make_rand_inputs(n, seed=21)]
    step, num_patch = 6000, (B // 4) // 1024
    # oracle: autograd on CPU, without the hash-decay term (applied inside the fused Adam here)
    ref = to.RefTrainer(sd)
    train_frac = float(np.clip((step - 1) / (25000 - 1), 0, 1))
    rend, hist = zo.model_forward(ref.p, batch, rin, train_frac, True, training=False)
    ls_ref = to.losses(batch, rend, hist, step, num_patch)
    sum(ls_ref.values()).backward()
    # B200 path
    cfg = configs.nuscenes_single()
    model = models.Model(cfg, training=True).cuda()
    model.load_state_dict(sd, strict=False)
    model.nerf_mlp.mlp_dtype = torch.float32
    tr = train.Trainer(model, cfg)
    cb = {k: v.cuda() for k, v in batch.items()}
    crin = [{k: v.cuda() for k, v in r.items()} for r in rin]
    r2, h2 = model(True, cb, train_frac, True, rand_inputs=crin)
    ls = train.compute_losses(cb, r2, h2, cfg, step, num_patch)
    for k, v in ls_ref.items():
        assert abs(float(ls[k]) - float(v)) <= 2e-4 * max(abs(float(v)), 1e-3), (k, float(ls[k]), float(v))
    sum(ls.values()).backward()
    for t in tr.tables:
        want = ref.p[t['name']].grad
        got = t['grad'].cpu()
        # sparse, sign-mixed sums: compare against the gradient's own scale
        assert_close(got, want, 2e-2, 'grad ' + t['name'])
        rel_l2 = float((got - want).norm() / want.norm())
        assert rel_l2 < 2e-2, (t['name'], rel_l2)
    for name, p in model.named_parameters():
        if name.endswith('embeddings'):
            continue
        want = ref.p[name].grad
        got = p.grad.cpu()
        rel_l2 = float((got - want).norm() / (want.norm() + 1e-20))
        assert rel_l2 < 2e-2, (name, rel_l2)


def test_fused_decay_adam_vs_torch():
    from nerf_lidar_b200 import _lib
    offs = np.array([0, 4920, 40864, 60000], np.int32)
    L, Cc = 3, 4
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(int(offs[-1]), Cc, generator=g) * 0.1
    mult, lr = 0.1, 0.01
    # reference: Adam on grad + d/dp [mult * sum_tables mean_levels mean_rows p^2], NaN scrubbed
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=lr, betas=(0.9, 0.99), eps=1e-15)
    pd = p0.clone().cuda()
    gd = torch.zeros_like(pd)
    m, v = torch.zeros_like(pd), torch.zeros_like(pd)
    offs_c = (C.c_int32 * 4)(*offs.tolist())
    lib = _lib.load()
    for step in range(1, 4):
        grad = torch.randn(p0.shape, generator=g) * 1e-3
        grad[5, 1] = float('nan')
        grad[6, 2] = float('inf')
        opt.zero_grad()
        per = torch.stack([(pr[offs[l]:offs[l + 1]] ** 2).mean(0) for l in range(L)])
        (mult * per.mean()).backward()
        pr.grad += grad
        pr.grad.nan_to_num_()
        opt.step()
        gd.copy_(grad.cuda() * 2.0)  # grad_scale 0.5 below (data-parallel mean)
        sumsq = torch.zeros(L, device='cuda')
        _lib.check(lib.nlb_adam_table_step(pd.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), offs_c, L, Cc,
                                           mult, lr, 0.9, 0.99, 1e-15, step, 0.5, sumsq.data_ptr(), _lib.stream()))
        assert float(gd.abs().sum()) == 0.0  # gradient buffer cleared for the next step
        want_sq = torch.stack([(pd[offs[l]:offs[l + 1]].double() ** 2).sum() for l in range(L)]).float()
        assert_close(sumsq, want_sq, 1e-4, 'per-level sum of squares of the updated table')
        ok = torch.isfinite(grad)
        assert_close(pd.cpu()[ok], pr.detach()[ok], 1e-5, f'params after step {step}')


def test_trainer_runs_and_learns():
    """Three optimisation steps on one batch reduce the data loss and keep parameters finite."""
    from nerf_lidar_b200 import configs, models, train
    B = 2048
    cfg = configs.nuscenes_single()
    model = models.Model(cfg, training=True).cuda()
    model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=2, table_std=1e-4).items()}, strict=False)
    tr = train.Trainer(model, cfg)
    batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=2)).items()}
    first = None
    for i in range(8):
        out = tr.train_step(batch, 6000 + i, (B // 4) // 1024)
        if first is None:
            first = float(out['data'])
        # the hash-decay value handed over by the fused optimizer pass equals the direct evaluation
        assert abs(float(model._hash_decay_value) - float(model.hash_decay_loss())) <= 1e-4 * float(model.hash_decay_loss())
    assert float(out['data']) < first
    for p in model.parameters():
        assert torch.isfinite(p).all()


def test_fused_mlp_sees_updated_weights():
    """The optimizer updates the dense parameters through raw pointers; the packed
    tensor-core operand images must be refreshed for the next forward."""
    from nerf_lidar_b200 import configs, models, train, ops
    B = 1024
    cfg = configs.nuscenes_single()
    model = models.Model(cfg, training=True).cuda()
    model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=4, table_std=0.1).items()}, strict=False)
    tr = train.Trainer(model, cfg)
    batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=4)).items()}
    for i in range(3):
        tr.train_step(batch, 6000 + i, 0)
    mlp = model.nerf_mlp
    feat = torch.randn(4096, 40, device='cuda') * 0.1
    vd = torch.nn.functional.normalize(torch.randn(128, 3, device='cuda'), dim=-1)
    with torch.no_grad():
        got = ops.nerf_mlp_forward(mlp, feat, vd, 32)
        mlp.fused_mlp = False
        want = mlp.heads(feat, vd, 32)
        mlp.fused_mlp = True
    for k in ('density', 'rgb', 'semantic', 'intensity'):
        assert_close(got[k].reshape(-1), want[k].reshape(-1).float(), 2e-2, 'fused vs torch after optimizer steps: ' + k)


def test_graphed_step_matches_eager():
    """Trainer.train_step_graphed (one CUDA graph per step) against the eager step on
    the same batches and injected random draws, over a learning-rate / anneal change.
    Two runs of the SAME schedule already differ: the order of the scatter's atomics moves table gradients by
    ulps, the bf16 rounding of the MLP operands turns that into 2^-9 jumps of single activations, and Adam's
    normalised steps turn near-zero gradients of either sign into +-lr.  So the graphed run is held to the
    noise floor measured between two eager runs (plus loose absolute caps), not to tuned constants."""
    from nerf_lidar_b200 import configs, models, train
    B = 1024
    cfg = configs.nuscenes_single()
    sd = {k: v.cuda() for k, v in synthetic.init_state_dict(seed=5, table_std=0.1).items()}
    trainers = []
    for _ in range(3):   # eager, graphed, eager again (the noise floor)
        model = models.Model(cfg, training=True).cuda()
        model.load_state_dict(sd, strict=False)
        trainers.append(train.Trainer(model, cfg))
    batches = [{k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=50 + i)).items()} for i in range(2)]
    n = batches[0]['origins'].shape[0]
    rins = [[{k: torch.from_numpy(v).cuda() for k, v in r.items()} for r in synthetic.make_rand_inputs(n, seed=60 + i)] for i in range(2)]
    for i in range(4):
        step = 6000 + 700 * i   # anneal, learning rate and bias corrections all move between steps
        a = trainers[0].train_step(batches[i % 2], step, 0, rins[i % 2])
        b = trainers[1].train_step_graphed(batches[i % 2], step, 0, rins[i % 2])
        trainers[2].train_step(batches[i % 2], step, 0, rins[i % 2])
        for k in a:
            assert abs(float(a[k]) - float(b[k])) <= 2e-3 * max(abs(float(a[k])), 1e-3), (i, k, float(a[k]), float(b[k]))
    params = [dict(t.model.named_parameters()) for t in trainers]
    for name, pa in params[0].items():
        graphed = (pa - params[1][name]).abs().reshape(-1)
        floor = (pa - params[2][name]).abs().reshape(-1)
        mean_g, mean_f = float(graphed.mean()), float(floor.mean())
        far_g, far_f = float((graphed > 5e-3).float().mean()), float((floor > 5e-3).float().mean())
        # (a step that is skipped or applied twice moves every entry by ~lr = 1e-2: 25x these bounds; observed
        # noise: mean 2e-4, 1 % of the entries beyond 5e-3, with run-to-run jumps of the same size)
        assert mean_g <= 3 * mean_f + 1e-4 and mean_g <= 1e-3, (name, mean_g, mean_f, float(graphed.max()))
        assert far_g <= 3 * far_f + 1e-2 and far_g < 5e-2, (name, far_g, far_f)
