"""Diagnostic: one replayed (CUDA-graph) training step against one eager step from IDENTICAL state.
Pre-optimizer gradients (flat dense-layer gradient per parameter, the three table gradients), the loss
dictionary and the Adam moments are compared; a second eager trainer gives the run-to-run noise floor."""
import sys

import torch

sys.path.insert(0, '.')
from nerf_lidar_b200 import configs, models, synthetic, train  # noqa: E402


def snap_hook(tr):
    orig = tr.optimizer_step
    tr.snap = dict(flat=torch.zeros_like(tr.flat_grad), **{t['name']: torch.zeros_like(t['grad']) for t in tr.tables})

    def f(step, reduce=True):
        tr.snap['flat'].copy_(tr.flat_grad)
        for t in tr.tables:
            tr.snap[t['name']].copy_(t['grad'])
        orig(step, reduce)
    tr.optimizer_step = f


def sync_state(src, dst):
    dst.flat.copy_(src.flat); dst.flat_m.copy_(src.flat_m); dst.flat_v.copy_(src.flat_v)
    dst.hash_decay_value.copy_(src.hash_decay_value)
    for a, b in zip(src.tables, dst.tables):
        b['param'].data.copy_(a['param'].data); b['m'].copy_(a['m']); b['v'].copy_(a['v'])
    dst._mark_packed_stale()


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    B = 1024
    cfg = configs.nuscenes_single()
    sd = {k: v.cuda() for k, v in synthetic.init_state_dict(seed=5, table_std=0.1).items()}
    trs = []
    for _ in range(3):
        m = models.Model(cfg, training=True).cuda()
        m.load_state_dict(sd, strict=False)
        t = train.Trainer(m, cfg)
        snap_hook(t)
        trs.append(t)
    batches = [{k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=50 + i)).items()} for i in range(2)]
    n = batches[0]['origins'].shape[0]
    rins = [[{k: torch.from_numpy(v).cuda() for k, v in r.items()} for r in synthetic.make_rand_inputs(n, seed=60 + i)] for i in range(2)]
    names = []
    off = 0
    for name, p in trs[0].model.named_parameters():
        if name.endswith('embeddings'):
            continue
        names.append((name, off, p.numel()))
        off += (p.numel() + 3) // 4 * 4
    for i in range(5):
        step = 6000 + 700 * i
        if i > 0 and '--free' not in sys.argv:
            sync_state(trs[0], trs[1]); sync_state(trs[0], trs[2])
        a = trs[0].train_step(batches[i % 2], step, 0, rins[i % 2])
        b = trs[1].train_step_graphed(batches[i % 2], step, 0, rins[i % 2])
        c = trs[2].train_step(batches[i % 2], step, 0, rins[i % 2])
        torch.cuda.synchronize()
        print(f'--- step {i} ({"eager warm-up + capture" if i == 0 else "replay"})')
        for k in a:
            print(f'  loss {k:12s} eager {float(a[k]):.7e} graph {float(b[k]):.7e} eager2 {float(c[k]):.7e}')
        for t in trs[0].tables:
            k = t['name']
            print(f'  grad {k:34s} graph-vs-eager rel {rel(trs[1].snap[k], trs[0].snap[k]):.3e}   eager2-vs-eager {rel(trs[2].snap[k], trs[0].snap[k]):.3e}')
        for name, o, k in names:
            ga, gb, gc = (t.snap['flat'][o:o + k] for t in trs)
            r1, r2 = rel(gb, ga), rel(gc, ga)
            flag = '  <<<' if r1 > 10 * r2 + 1e-5 else ''
            print(f'  grad {name:34s} graph-vs-eager rel {r1:.3e} max {float((gb - ga).abs().max()):.2e}   eager2-vs-eager {r2:.3e}   |g| {float(ga.norm()):.2e}{flag}')
        for name, o, k in names:
            pa, pb, pc = (t.flat[o:o + k] for t in trs)
            print(f'  param {name:34s} mean|d| graph {float((pb - pa).abs().mean()):.3e} max {float((pb - pa).abs().max()):.2e}   eager2 {float((pc - pa).abs().mean()):.3e} max {float((pc - pa).abs().max()):.2e}')
        print(f'  flat_m graph-vs-eager {rel(trs[1].flat_m, trs[0].flat_m):.3e} eager2 {rel(trs[2].flat_m, trs[0].flat_m):.3e}')
        print(f'  flat_v graph-vs-eager {rel(trs[1].flat_v, trs[0].flat_v):.3e} eager2 {rel(trs[2].flat_v, trs[0].flat_v):.3e}')
        print(f'  flat   graph-vs-eager mean|d| {float((trs[1].flat - trs[0].flat).abs().mean()):.3e} eager2 {float((trs[2].flat - trs[0].flat).abs().mean()):.3e}')
        for ta, tb, tc in zip(trs[0].tables, trs[1].tables, trs[2].tables):
            print(f'  table {ta["name"]:30s} param mean|d| graph {float((tb["param"] - ta["param"]).abs().mean()):.3e} eager2 {float((tc["param"] - ta["param"]).abs().mean()):.3e}')


if __name__ == '__main__':
    main()
