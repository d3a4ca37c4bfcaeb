#!/usr/bin/env python
"""bench.py -- headline benchmark of the zipnerf hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (this repo, CUDA)
  python bench.py --impl reference --gpus N --steps K ...   (reference CPU path)
  torchrun ... bench.py --gpus N ...                        (N > 1, one rank per GPU)

Workload (BASELINE.json configs[1]): one full training step of the
nuscenes_single.gin camera+LiDAR run -- 8192 camera rays + 2048 extra LiDAR rays
per GPU (Z/internal/datasets.py:352-403), forward through the three sampling
levels, all loss terms, backward, hash-decay + Adam over 77.66 M parameters --
on synthetic nuScenes-shaped rays and random-init weights.  Metric: training
rays/s counted on the nominal 8192-ray batch like the reference's
train_rays_per_sec (Z/train.py:485); weak scaling (per-GPU batch fixed).

One JSON line on stdout (rank 0).  `value` has the batch resident in HBM; `e2e`
copies every step's batch from pinned host memory and reads the loss back.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 8192              # nominal rays per GPU per step (BASELINE configs[1])
SAMPLES = (64, 64, 32)
# algorithmic bytes per sample point, SURVEY.md 8(d): 12 (xyz) + L*C*4 (out) + L*8*C*4 (gathers)
ALG_BYTES = {
    'nerf_encode_fwd': (12 + 10 * 4 * 4 + 10 * 8 * 4 * 4, 32 * 7),        # 1452 B/point, 224 points/ray
    'nerf_encode_bwd': (12 + 10 * 4 * 4 + 2 * 10 * 8 * 4 * 4, 32 * 7),    # 2732
    'prop6_fwd': (12 + 6 * 4 + 6 * 8 * 4, 64 * 7),                        # 228
    'prop8_fwd': (12 + 8 * 4 + 8 * 8 * 4, 64 * 7),                        # 300
    'prop6_bwd': (12 + 6 * 4 + 2 * 6 * 8 * 4, 64 * 7),                    # 420
    'prop8_bwd': (12 + 8 * 4 + 2 * 8 * 8 * 4, 64 * 7),                    # 556
}
ADAM_BYTES_PER_PARAM = 32  # p,g,m,v read + p,m,v,g written
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per ABI call from the committed `ncu --set full`
# captures of this workload (profiles/r1_s14_kernels_full.txt, r1_s14_fwd_kernels_full.txt; the 6-level proposal
# forward and the Adam pass from r1_s12_kernels_full.txt).  Far below the algorithmic bytes wherever a table (or
# one level group of it) stays L2-resident: the algorithmic figure counts every corner gather /
# read-modify-write as HBM traffic.
NCU_DRAM_BYTES = {
    'nerf_encode_fwd': (95.7 + 198.1 + 234.7) * 1e6 + (61.0 + 14.6 + 18.4) * 1e6,                 # three level groups
    'nerf_encode_bwd': (120.3 + 156.9 + 156.9 + 112.2) * 1e6 + (6.3 + 61.6 + 158.3 + 4.2) * 1e6,  # four level groups
    'prop6_fwd': 40.3e6 + 42.8e6,
    'prop8_fwd': 58.1e6 + 62.6e6,
    'prop6_bwd': 107.3e6 + 4.6e6 + 18.4e6,                      # scatter + proposal-MLP backward
    'prop8_bwd': (95.5 + 4.6) * 1e6 + (128.6 + 4.0) * 1e6 + 23.6e6,  # dense launch + hashed pair launch + MLP backward
    'adam_table': (0.9604 + 0.1730 + 0.1057 + 0.9026 + 0.1170 + 0.0482) * 1e9 / 3.0,
}


def _env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(',')])

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        """Median SM clock / throttle reasons of the samples taken in [t0, t1] (the timed
        region; the sampler runs from before the warm-up so short regions still get samples;
        falls back to the samples under load nearest to it)."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r[1:] for r in self.rows if t0 is None or (t0 - 0.05 <= r[0] <= t1 + 0.15)]
        if not rows:
            rows = [r[1:] for r in self.rows if t0 is None or r[0] >= t0 - 1.0] or [r[1:] for r in self.rows]
        sm = sorted(int(r[1]) for r in rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------- reference / CPU arm
def cpu_reference_steps(steps, warmup, sample_rays):
    """Times the oracle port of the reference training step (oracle/train_oracle.py)
    on the host cores with all threads, on a bounded sample of the workload."""
    import torch
    from nerf_lidar_b200 import synthetic
    from oracle import train_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic.init_state_dict(seed=0, table_std=1e-4)
    tr = train_oracle.RefTrainer(sd)
    batches = []
    for i in range(2):
        b = synthetic.to_torch(synthetic.make_train_batch(sample_rays, seed=100 + i))
        n = b['origins'].shape[0]
        rin = [{k: torch.from_numpy(v) for k, v in r.items()} for r in synthetic.make_rand_inputs(n, seed=i)]
        batches.append((b, rin))
    num_patch = (sample_rays // 4) // 1024
    times = []
    for i in range(warmup + steps):
        b, rin = batches[i % 2]
        t0 = time.perf_counter()
        tr.step(b, rin, 6000 + i, num_patch)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return sample_rays / med, med, cores, torch.get_num_threads()


def run_reference_arm(args):
    rank = _env_int('RANK', 0)
    if rank != 0:
        return
    sample = 1024
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    rps, med, cores, threads = cpu_reference_steps(steps, warmup, sample)
    line = {
        'impl': 'reference', 'metric': 'train_rays_per_sec', 'value': rps, 'unit': 'rays/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': med * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'zipnerf nuscenes_single.gin camera+LiDAR training step (BASELINE configs[1]); '
                               'reference path = oracle port on host cores', 'rays_per_step_nominal': sample,
                   'rays_through_model': sample + sample // 4, 'samples': list(SAMPLES)},
        'cpu_baseline': {'value': rps, 'unit': 'rays/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{sample}-ray (+{sample // 4} LiDAR) training steps, median of {steps}, '
                                   f'{cores} host cores'},
        'e2e': {'value': rps, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------- CUDA arm
def run_cuda_arm(args):
    import torch
    import torch.distributed as dist
    from nerf_lidar_b200 import _lib, configs, models, synthetic, train

    world, rank, local = _env_int('WORLD_SIZE', 1), _env_int('RANK', 0), _env_int('LOCAL_RANK', 0)
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (this repository has no CPU fallback; use --impl reference)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL prints its version banner on STDOUT (at NCCL_DEBUG=VERSION and WARN); stdout carries the one
        # JSON line, so NCCL's log goes to stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()
    cfg = configs.nuscenes_single(use_intensity=True, instance_obj=False)
    cfg.batch_size = BATCH * world
    model = models.Model(cfg, training=True).to(dev)
    model.load_state_dict({k: v.to(dev) for k, v in synthetic.init_state_dict(seed=0, table_std=1e-4).items()},
                          strict=False)
    trainer = train.Trainer(model, cfg, world=world, rank=rank)
    num_patch = (BATCH // 4) // (cfg.patch_size ** 2)

    # a pool of distinct batches: pinned on the host (e2e) and resident on the device (value)
    n_pool = 4
    host, resident = [], []
    for i in range(n_pool):
        b = synthetic.make_train_batch(BATCH, seed=1000 + 17 * rank + i)
        hb = {k: torch.from_numpy(v).pin_memory() for k, v in b.items()}
        host.append(hb)
        resident.append({k: v.to(dev) for k, v in hb.items()})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    loss_pin = torch.zeros(1).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_fn = trainer.train_step if args.eager else trainer.train_step_graphed

    def timed_loop(n_steps, first_step, e2e, fn=None):
        fn = fn or step_fn
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            if e2e and fn == trainer.train_step:
                b = {k: v.to(dev, non_blocking=True) for k, v in host[i % n_pool].items()}
            elif e2e:
                b = host[i % n_pool]       # pinned host tensors: the graphed step copies them into its static buffers
            else:
                b = resident[i % n_pool]
            out = fn(b, first_step + i, num_patch)
            if e2e:
                loss_pin.copy_(out['loss'].reshape(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()  # the user reads the loss every step
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    step0 = 6000  # steady state: past the pose-refinement window (Config.end_step = 5000)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timed_loop(max(3, args.warmup), step0, False)
    t0 = sampler.mark()
    ms = timed_loop(args.steps, step0 + 100, False)
    t1 = sampler.mark()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    timed_loop(2, step0 + 200, True)
    ms_e2e = timed_loop(args.steps, step0 + 300, True)
    # per-kernel device times (roofline): the same steps issued eagerly with CUDA events around every
    # ABI call on the launching stream (events cannot be read back from inside a replayed graph)
    timed_loop(2, step0 + 400, False, trainer.train_step)
    _lib.TIMER = _lib.KernelTimer() if rank == 0 else None
    launches0 = _lib.LAUNCHES
    n_prof = min(args.steps, 5)
    ms_eager = timed_loop(n_prof, step0 + 500, False, trainer.train_step)
    launches = (_lib.LAUNCHES - launches0) // n_prof * args.steps   # ABI calls per step x timed steps
    kt = _lib.TIMER.summary() if _lib.TIMER is not None else {}
    _lib.TIMER = None

    # ---- rendering (BASELINE configs[2] / [4]): LiDAR sweep and one 1600x900 camera frame through
    # models.render_image, rays sharded contiguously over the ranks, one packed gather at the end
    render = None
    if not args.no_render:
        class _Acc:
            process_index, num_processes, is_main_process = rank, world, rank == 0
        render = {}
        for name, make, reps in (('lidar_sweep_32x1084', synthetic.make_lidar_sweep, 5),
                                 ('camera_frame_1600x900', synthetic.make_camera_frame, 2)):
            rb = {k: torch.from_numpy(v).to(dev) for k, v in make(seed=0).items()}
            n_rays = rb['origins'].shape[0]
            with torch.no_grad():
                models.render_image(model, _Acc, rb, False, cfg, image=False, verbose=False)  # warm-up
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    out = models.render_image(model, _Acc, rb, False, cfg, image=False, verbose=False)
                e1.record()
                barrier()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            render[name] = {'rays': n_rays, 'ms': float(t.item()), 'rays_per_s': n_rays / float(t.item()) * 1e3,
                            'outputs': sorted(k for k in out if not k.startswith('ray_'))}
            del rb, out
        model.train()
        model.training = True

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'MEASURED_PEAKS.json hbm_gbs (measured)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    rays_model = BATCH + BATCH // cfg.lidar_batch_ratio
    # dominant kernel of the step by measured device time
    per_kernel = {}
    for name, (count, total_ms) in kt.items():
        avg = total_ms / max(count, 1)
        if name in ALG_BYTES:
            per_point, pts = ALG_BYTES[name]
            alg = per_point * pts * rays_model
        elif name == 'adam_table':
            alg = ADAM_BYTES_PER_PARAM * 77346760 / 3.0  # three tables per step, averaged per launch
        else:
            alg = None
        per_kernel[name] = {'launches': count, 'avg_ms': avg, 'share_of_step': total_ms / n_prof / (ms / args.steps),
                            'alg_gbs': (alg / (avg * 1e-3) / 1e9) if alg else None,
                            'dram_bytes_ncu': NCU_DRAM_BYTES.get(name)}
    dom = max((k for k in per_kernel if per_kernel[k]['alg_gbs']), key=lambda k: per_kernel[k]['share_of_step'],
              default=None)
    roofline = None
    if dom:
        a = per_kernel[dom]['alg_gbs']
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': a, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': a / hbm_peak, 'traffic': NCU_DRAM_BYTES.get(dom), 'peak_source': peak_src,
                    'avg_launch_ms': per_kernel[dom]['avg_ms'],
                    'algorithmic_bytes_per_launch': ALG_BYTES[dom][0] * ALG_BYTES[dom][1] * rays_model
                    if dom in ALG_BYTES else None,
                    'note': 'achieved = SURVEY 8(d) algorithmic bytes / measured launch time; it can exceed the HBM '
                            'peak because merged same-cell corners and L2-resident level groups never reach DRAM '
                            '(traffic = DRAM bytes per launch from the committed ncu --set full capture)'}
    # the only dense contraction on the path: fused NerfMLP kernels against the measured sustained bf16 rate
    roofline_mlp = None
    tf_peak = float(peaks.get('bf16_tflops_sustained', 1356.8))
    rows_mlp = rays_model * SAMPLES[-1]
    for name, macs in (('nerf_mlp_fwd', 264192), ('nerf_mlp_bwd', 257024)):
        if name in per_kernel:
            t = 2.0 * macs * rows_mlp / (per_kernel[name]['avg_ms'] * 1e-3) / 1e12
            per_kernel[name]['tflops'] = t
            if name == 'nerf_mlp_fwd':
                roofline_mlp = {'bound': 'tensor', 'kernel': name, 'achieved': t, 'peak': tf_peak, 'unit': 'TFLOP/s',
                                'frac': t / tf_peak, 'traffic': None,
                                'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if 'bf16_tflops_sustained'
                                in peaks else 'fallback 1356.8 TFLOP/s', 'avg_launch_ms': per_kernel[name]['avg_ms'],
                                'flops_per_row': 2 * macs, 'rows': rows_mlp}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rps, med, cores, threads = cpu_reference_steps(3, 1, 512)
        cpu = {'value': rps, 'unit': 'rays/s', 'cores': threads, 'kind': 'port',
               'sample': f'512-ray (+128 LiDAR) training steps of the oracle port, median of 3, {cores} host cores'}
    value = BATCH * world * args.steps / (ms * 1e-3)
    e2e = BATCH * world * args.steps / (ms_e2e * 1e-3)
    line = {
        'metric': 'train_rays_per_sec', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32 grid/compositing, bf16 MLP operands (fp32 accumulate)',
        'data': 'synthetic',
        'config': {'workload': 'zipnerf nuscenes_single.gin camera+LiDAR training step (BASELINE configs[1])',
                   'rays_per_step_nominal_per_gpu': BATCH, 'rays_through_model_per_gpu': rays_model,
                   'global_batch': BATCH * world, 'samples': list(SAMPLES), 'multisamples': 7,
                   'params': 77656777, 'parallelism': f'dp{world}',
                   'launch': 'eager' if args.eager else 'whole training step replayed as one CUDA graph',
                   'eager_ms_per_step': ms_eager / n_prof,
                   'l2_policy': 'inputs larger than L2: 1.24 GB of table + optimizer state streamed every step, '
                                'batches rotate over a pool of 4'},
        'clocks': clocks,
        'e2e': {'value': e2e, 'unit': 'rays/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4},
        'gpu_launches': launches,
        'roofline': roofline,
        'roofline_mlp': roofline_mlp,
        'render': render,
        'kernels': per_kernel,
        'cpu_baseline': cpu,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _emit(line: dict):
    """The one JSON line goes to the REAL stdout; while the run is in progress file descriptor 1 points at
    stderr so that nothing a library prints (NCCL's version banner, ...) can precede or follow it."""
    data = (json.dumps(line) + '\n').encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    sys.stdout.flush()
    os.write(fd, data)


_REAL_STDOUT = None


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-render', action='store_true', help='skip the rendering throughput measurement')
    ap.add_argument('--eager', action='store_true', help='issue the step eagerly instead of replaying a CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == '__main__':
    main()
