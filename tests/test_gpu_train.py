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
from tests.helpers import assert_close, torch_heads, use_torch_heads

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
    use_torch_heads(model, torch.float32)
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
        want = torch_heads(mlp, feat, vd, 32, torch.bfloat16)
    for k in ('density', 'rgb', 'semantic', 'intensity'):
        assert_close(got[k].reshape(-1), want[k].reshape(-1).float(), 2e-2, 'fused vs torch after optimizer steps: ' + k)


def _snapshot_gradients(tr):
    """Copies of the pre-optimizer gradient buffers, taken by the step itself (so a captured step takes them at
    replay time): the optimizer pass zeroes the buffers it consumes."""
    orig = tr.optimizer_step
    tr.snap = dict(flat=torch.zeros_like(tr.flat_grad), **{t['name']: torch.zeros_like(t['grad']) for t in tr.tables})

    def step_with_snapshot(step, reduce=True):
        tr.snap['flat'].copy_(tr.flat_grad)
        for t in tr.tables:
            tr.snap[t['name']].copy_(t['grad'])
        orig(step, reduce)
    tr.optimizer_step = step_with_snapshot


def _copy_state(src, dst):
    dst.flat.copy_(src.flat); dst.flat_m.copy_(src.flat_m); dst.flat_v.copy_(src.flat_v)
    dst.hash_decay_value.copy_(src.hash_decay_value)
    for a, b in zip(src.tables, dst.tables):
        b['param'].data.copy_(a['param'].data); b['m'].copy_(a['m']); b['v'].copy_(a['v'])
    dst._mark_packed_stale()


def _rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _dense_segments(tr):
    out, off = [], 0
    for name, p in tr.model.named_parameters():
        if not name.endswith('embeddings'):
            out.append((name, off, p.numel()))
            off += (p.numel() + 3) // 4 * 4
    return out


def test_graphed_step_matches_eager():
    """ONE replay of the captured step (Trainer.train_step_graphed, the path bench.py times) against ONE eager
    step from IDENTICAL state, over a learning-rate / anneal change and alternating batches: loss dictionary,
    pre-optimizer gradients (per dense parameter and per table) and the Adam moments.

    Well-conditioned on purpose.  Free-running trainers drift apart chaotically -- with adam_eps = 1e-15 a
    near-zero gradient entry of either sign is a +-lr step, so ONE flipped entry after the first step is a 1e-2
    parameter difference that feeds the next forward (tools/graph_vs_eager.py --free: two EAGER trainers diverge
    exactly as fast as graph vs eager, 1e-5 .. 2e-4 mean after four steps depending on which entries flip).
    Parameters are therefore compared through m / v, and on the entries whose gradient is clear of zero."""
    from nerf_lidar_b200 import configs, models, train
    B = 1024
    cfg = configs.nuscenes_single()
    sd = {k: v.cuda() for k, v in synthetic.init_state_dict(seed=5, table_std=0.1).items()}
    eager, graphed = [], []
    trainers = []
    for _ in range(2):
        model = models.Model(cfg, training=True).cuda()
        model.load_state_dict(sd, strict=False)
        tr = train.Trainer(model, cfg)
        _snapshot_gradients(tr)
        trainers.append(tr)
    eager, graphed = trainers
    batches = [{k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=50 + i)).items()} for i in range(2)]
    n = batches[0]['origins'].shape[0]
    rins = [[{k: torch.from_numpy(v).cuda() for k, v in r.items()} for r in synthetic.make_rand_inputs(n, seed=60 + i)] for i in range(2)]
    segments = _dense_segments(eager)
    for i in range(4):
        step = 6000 + 700 * i   # anneal, learning rate and bias corrections all move between steps
        _copy_state(eager, graphed)
        before = eager.flat.clone()
        a = eager.train_step(batches[i % 2], step, 0, rins[i % 2])
        b = graphed.train_step_graphed(batches[i % 2], step, 0, rins[i % 2])   # i = 0: eager warm-up + capture
        assert set(a) == set(b)
        for k in a:
            assert abs(float(a[k]) - float(b[k])) <= 1e-5 * max(abs(float(a[k])), 1e-3), (i, k, float(a[k]), float(b[k]))
        # pre-optimizer gradients: atomics order (tables, proposal MLPs) and fp32 summation order only
        for t in eager.tables:
            r = _rel_l2(graphed.snap[t['name']], eager.snap[t['name']])
            assert r <= 1e-5, (i, t['name'], r)
        for name, o, k in segments:
            ga, gb = eager.snap['flat'][o:o + k], graphed.snap['flat'][o:o + k]
            assert float(ga.abs().max()) > 0, name
            r = _rel_l2(gb, ga)
            assert r <= 1e-4, (i, name, r)
        # gradient buffers consumed and cleared, moments and parameters updated identically
        assert float(graphed.flat_grad.abs().max()) == 0.0 and all(float(t['grad'].abs().max()) == 0.0 for t in graphed.tables)
        assert _rel_l2(graphed.flat_m, eager.flat_m) <= 1e-4 and _rel_l2(graphed.flat_v, eager.flat_v) <= 1e-4
        for ta, tb in zip(eager.tables, graphed.tables):
            assert _rel_l2(tb['m'], ta['m']) <= 1e-5 and _rel_l2(tb['v'], ta['v']) <= 1e-5, (i, ta['name'])
            assert float((tb['param'] - ta['param']).abs().mean()) <= 1e-8, (i, ta['name'])
        moved = (eager.flat - before).abs()
        assert float(moved.mean()) > 1e-4                      # the step did move the parameters (~lr per entry)
        clear = eager.snap['flat'].abs() > 1e-6 * eager.snap['flat'].abs().max()
        assert float((graphed.flat - eager.flat).abs()[clear].max()) <= 1e-5, i
        assert float((graphed.flat - eager.flat).abs().mean()) <= 1e-6, i


def test_no_uninitialised_reads_in_the_step():
    """Every buffer the operators allocate with torch.empty is fully written before it is read: a training step and
    a render with those allocations poisoned (NaN) give the same losses / gradients / outputs as the plain run."""
    from nerf_lidar_b200 import configs, models, train
    from tests.helpers import poisoned_empty
    B = 256
    cfg = configs.nuscenes_single()
    sd = {k: v.cuda() for k, v in synthetic.init_state_dict(seed=7, table_std=0.1).items()}
    batch = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=70)).items()}
    n = batch['origins'].shape[0]
    rin = [{k: torch.from_numpy(v).cuda() for k, v in r.items()} for r in synthetic.make_rand_inputs(n, seed=71)]
    results = []
    for poison in (False, True):
        model = models.Model(cfg, training=True).cuda()
        model.load_state_dict(sd, strict=False)
        tr = train.Trainer(model, cfg)
        _snapshot_gradients(tr)
        with poisoned_empty(poison):
            model.eval(); model.training = False
            with torch.no_grad():   # (before the step: the scatter's atomics order would move the tables by ulps)
                rend, _ = model(False, batch, 1.0, True)
            model.train(); model.training = True
            out = tr.train_step(batch, 6000, 0, rin)
        results.append((out, tr.snap, rend[-1]))
    (la, ga, ra), (lb, gb, rb) = results
    for k in la:
        assert np.isfinite(float(lb[k])) and abs(float(la[k]) - float(lb[k])) <= 1e-5 * max(abs(float(la[k])), 1e-3), k
    for k in ga:
        assert torch.isfinite(gb[k]).all(), k
        assert _rel_l2(gb[k], ga[k]) <= 1e-4, (k, _rel_l2(gb[k], ga[k]))
    for k in ('rgb', 'depth', 'semantic', 'intensity', 'acc', 'distance_median'):
        assert torch.isfinite(rb[k]).all(), k
        assert torch.equal(rb[k], ra[k]), k      # the forward has no atomics: bit-identical
