"""Worker of tests/test_gpu_dp.py (launched by torch.distributed.run, one rank per GPU): a data-parallel training
step -- eager and CUDA-graph -- against the same step done on ONE GPU: the per-rank gradients averaged (DDP's mean,
Z/train.py:459) and the same fused hash-decay + Adam pass.  Every rank checks its own slice of the sharded moments
and the all-gathered parameters."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerf_lidar_b200 import configs, models, synthetic, train  # noqa: E402


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    B = 1024
    cfg = configs.nuscenes_single()
    sd = {k: v.to(dev) for k, v in synthetic.init_state_dict(seed=5, table_std=0.1).items()}

    def fresh(w, r):
        m = models.Model(cfg, training=True).to(dev)
        m.load_state_dict(sd, strict=False)
        return train.Trainer(m, cfg, world=w, rank=r)

    dp, ref = fresh(world, rank), fresh(1, 0)
    # every rank's rays and random draws (all ranks build all of them: the single-GPU reference needs every half)
    halves, rins = [], []
    for r in range(world):
        halves.append({k: v.to(dev) for k, v in synthetic.to_torch(synthetic.make_train_batch(B, seed=80 + r)).items()})
        n = halves[-1]['origins'].shape[0]
        rins.append([{k: torch.from_numpy(v).to(dev) for k, v in x.items()} for x in synthetic.make_rand_inputs(n, seed=90 + r)])
    seq = {'e': dp.train_step, 'g': dp.train_step_graphed}
    for i, fn in enumerate([seq[c] for c in os.environ.get('DPW_SEQ', 'egg')]):   # eager, capture, replay
        step = 6000 + 500 * i
        # identical state on both sides (free-running trainers drift apart chaotically, see test_gpu_train.py)
        if i > 0:
            # ONE source for the state: the single-GPU references of the two ranks are independent computations
            # whose gradients differ in the last bits (atomics order), and Adam with eps = 1e-15 turns a
            # near-zero gradient of either sign into a +-lr step -- copied rank-locally, the "replicated"
            # parameters of the data-parallel trainer would differ between the ranks (measured: 1e-4-level
            # moment differences at the next step).  Rank 0's reference state goes to every rank.
            with torch.no_grad():
                for t in ref.tables:
                    for x in (t['param'].data, t['m'], t['v']):
                        dist.broadcast(x, 0)
                for x in (ref.flat, ref.flat_m, ref.flat_v, ref.hash_decay_value):
                    dist.broadcast(x, 0)
            ref._mark_packed_stale()
            for a, b in zip(ref.tables, dp.tables):
                b['param'].data.copy_(a['param'].data)
                lo, cnt = b['arena']['lo'], b['arena']['cnt']
                b['m'][:cnt].copy_(a['m'][lo:lo + cnt]); b['v'][:cnt].copy_(a['v'][lo:lo + cnt])
            fa = dp.flat_arena
            dp.flat.copy_(ref.flat)
            dp.flat_m[:fa['cnt']].copy_(ref.flat_m[fa['lo']:fa['lo'] + fa['cnt']])
            dp.flat_v[:fa['cnt']].copy_(ref.flat_v[fa['lo']:fa['lo'] + fa['cnt']])
            dp.hash_decay_value.copy_(ref.hash_decay_value)
            dp._mark_packed_stale()
        out = fn(halves[rank], step, 0, rins[rank])
        dp.sync()   # a replayed step leaves the NeRF table's all-gather in flight (it hides under the next step)
        # one GPU: both halves' gradients accumulate in the buffers, then their mean drives the same optimizer pass
        ref_losses = []
        for r in range(world):
            ls, main_l, prop_l = ref.forward_losses(halves[r], step, 0, rins[r])
            (main_l + prop_l).backward()
            ref_losses.append(ls)
        ref.flat_grad.mul_(1.0 / world)
        for t in ref.tables:
            t['grad'].mul_(1.0 / world)
        ref.optimizer_step(step)
        torch.cuda.synchronize()
        # Verification under no_grad: `t['param']` are Parameters, and an autograd graph built over them HERE (on the
        # default stream) that stays alive -- `chk` survives into the next iteration -- keeps their AccumulateGrad nodes
        # bound to the default stream; the next captured backward then joins that stream into the capture
        # (cudaErrorStreamCaptureIsolation in GraphTask::exec_post_processing).
        with torch.no_grad():
            for k in ref_losses[rank]:
                if k in ('loss', 'hash_decay'):
                    continue
                a, b = float(out[k]), float(ref_losses[rank][k])
                assert abs(a - b) <= 1e-5 * max(abs(b), 1e-3), (i, rank, k, a, b)
            for a, b in zip(ref.tables, dp.tables):
                lo, cnt = b['arena']['lo'], b['arena']['cnt']
                rm, rv = rel(b['m'][:cnt], a['m'][lo:lo + cnt]), rel(b['v'][:cnt], a['v'][lo:lo + cnt])
                print(f'step {i} rank {rank} {a["name"]}: rel m {rm:.2e} v {rv:.2e}', flush=True)
                assert rm <= 1e-5 and rv <= 1e-5, (i, rank, a['name'], rm, rv)
                assert float((b['param'] - a['param']).abs().mean()) <= 1e-8, (i, rank, a['name'], 'param')
                assert float(b['grad'].abs().max()) == 0.0 and float(b['arena']['grad'].abs().max()) == 0.0
            fa = dp.flat_arena
            assert rel(dp.flat_m[:fa['cnt']], ref.flat_m[fa['lo']:fa['lo'] + fa['cnt']]) <= 1e-4, (i, rank, 'flat m')
            assert rel(dp.flat_v[:fa['cnt']], ref.flat_v[fa['lo']:fa['lo'] + fa['cnt']]) <= 1e-4, (i, rank, 'flat v')
            assert float((dp.flat - ref.flat).abs().mean()) <= 1e-6, (i, rank, 'flat params')
            assert float(fa['grad'].abs().max()) == 0.0
            assert abs(float(dp.hash_decay_value) - float(ref.hash_decay_value)) <= 1e-5 * float(ref.hash_decay_value)
            # parameters identical on every rank after the all-gather
            chk = torch.stack([dp.flat.double().sum()] + [t['param'].double().sum() for t in dp.tables])
            lst = [torch.empty_like(chk) for _ in range(world)]
            dist.all_gather(lst, chk)
            assert all(torch.equal(lst[0], x) for x in lst), (i, rank, 'ranks disagree after the all-gather')
    # pose-refinement window under data parallelism (Z/train.py:97,200-221,464-466: the posenet is DDP-wrapped):
    # every rank refines its own rays, the corrections' gradients are averaged over the ranks before their Adam
    # step, so the corrections stay identical everywhere
    if os.environ.get('DPW_POSE', '1') == '1':
        from nerf_lidar_b200 import posenet
        for h in halves:
            h['glo_idx'] = torch.from_numpy(synthetic.sensor_index({'lidar_mask': h['lidar_mask'].cpu().numpy()})).to(dev)
        net, opt, lr_fn = posenet.create_posenet(1, cfg, num_lidars=1, device=dev)
        dp.attach_posenet(net, opt, lr_fn)
        for i in range(2):
            step = 1000 + i
            out = dp.train_step_graphed(halves[rank], step, 0, rins[rank])   # falls back to the eager window step
            dp.sync()
            assert torch.isfinite(out['loss']).all()
        lst = [torch.empty_like(net.r.data) for _ in range(world)]
        dist.all_gather(lst, net.r.data.contiguous())
        assert all(torch.equal(lst[0], x) for x in lst), 'pose corrections differ between the ranks'
        assert float(net.r.detach().abs().max()) > 0 and float(net.t.detach().abs().max()) == 0.0
        print(f'rank {rank} pose corrections after 2 window steps: max |r| {float(net.r.detach().abs().max()):.2e}', flush=True)
        dp.attach_posenet(None, None, None)
    dist.barrier()
    if rank == 0:
        print('dp ok')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
