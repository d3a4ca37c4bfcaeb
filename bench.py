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
def workload_config(world):
    """`config` of BOTH arms: the workload is the same, the CPU arm times a bounded sample of it."""
    return {'workload': 'zipnerf nuscenes_single.gin camera+LiDAR training step (BASELINE configs[1])',
            'rays_per_step_nominal_per_gpu': BATCH, 'rays_through_model_per_gpu': BATCH + BATCH // 4,
            'global_batch': BATCH * world, 'samples': list(SAMPLES), 'multisamples': 7, 'params': 77656777,
            'parallelism': f'dp{world}'}


def cpu_reference_steps(steps, warmup, sample_rays, forward_only=False):
    """Times the oracle port of the reference path (oracle/train_oracle.py, oracle/zipnerf_oracle.py: the
    reference's Python restated on torch-CPU with a torch-gather grid) on the host cores
    with all threads, on a bounded sample of the workload: a full training step, or the forward only
    (rand=False: rendering)."""
    import torch
    from nerf_lidar_b200 import synthetic
    from oracle import train_oracle, zipnerf_oracle
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
        if forward_only:
            with torch.no_grad():
                zipnerf_oracle.model_forward(sd, b, None, 1.0)
        else:
            tr.step(b, rin, 6000 + i, num_patch)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return sample_rays / med, med, cores, torch.get_num_threads()


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU path (kind "port": the reference's grid encoder is CUDA-only and its
    Python cannot travel to the GPU box) on this arm's config / metric / unit, every step a
    bounded 1024-ray sample of the 8192-ray workload, --steps / --warmup as given."""
    rank = _env_int('RANK', 0)
    if rank != 0:
        return
    sample = 1024
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    rps, med, cores, threads = cpu_reference_steps(steps, warmup, sample)
    line = {
        'impl': 'reference', 'metric': 'train_rays_per_sec', 'value': rps, 'unit': 'rays/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': med * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(max(1, args.gpus)),
        'cpu_baseline': {'value': rps, 'unit': 'rays/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{sample}-ray (+{sample // 4} LiDAR) sample of the {BATCH}-ray training step per '
                                   f'timed step, median of {steps}, {cores} host cores'},
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

    def timed_loop(n_steps, first_step, e2e, fn=None, pool=None):
        fn = fn or step_fn
        resident_ = pool or resident
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            if e2e and fn == trainer.train_step:
                b = {k: v.to(dev, non_blocking=True) for k, v in host[i % n_pool].items()}
            elif e2e:
                b = host[i % n_pool]       # pinned host tensors: the graphed step copies them into its static buffers
            else:
                b = resident_[i % n_pool]
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

    # ---- secondary regime: the pose-refinement window 0 < step < 5000 -- LearnPose corrections
    # applied to the batch with gradients, ray-geometry gradients out of the three levels, Adam on the corrections;
    # the whole step replayed as one CUDA graph like the steady state
    pose = None
    if world == 1 and not args.no_pose_window:
        from nerf_lidar_b200 import posenet
        net, opt, lr_fn = posenet.create_posenet(1, cfg, num_lidars=1, device=dev)
        trainer.attach_posenet(net, opt, lr_fn)
        pool = [dict(b, glo_idx=torch.from_numpy(synthetic.sensor_index({'lidar_mask': b['lidar_mask'].cpu().numpy()})).to(dev))
                for b in resident]
        timed_loop(3, 1000, False, pool=pool)
        ms_pose = timed_loop(args.steps, 1100, False, pool=pool)
        timed_loop(2, 1300, False, trainer.train_step, pool=pool)
        _lib.TIMER = _lib.KernelTimer()
        ms_pose_eager = timed_loop(3, 1400, False, trainer.train_step, pool=pool)
        kt_pose = {k: v[1] / max(v[0], 1) for k, v in _lib.TIMER.summary().items() if 'input_bwd' in k}
        _lib.TIMER = None
        pose = {'steps': '0 < step < Config.end_step = 5000 (gin: learn_R, not learn_t)',
                'ms_per_step': ms_pose / args.steps, 'rays_per_s': BATCH * args.steps / (ms_pose * 1e-3),
                'eager_ms_per_step': ms_pose_eager / 3, 'input_gradient_kernels_ms': kt_pose,
                'correction_moved': float(net.r.detach().abs().max())}
        trainer.attach_posenet(None, None, None)
        del pool

    # ---- rendering (BASELINE configs[2] / [4]) through models.render_image: rays sharded contiguously over the
    # ranks, local chunks replayed as CUDA graphs, one packed gather at the end.  configs[4] = 4 cameras at
    # 1600 x 900 with the video sampling num_prop_samples = (256, 64) (Z/render_video.py:130) + one LiDAR sweep.
    render = None
    if not args.no_render:
        import copy

        class _Acc:
            process_index, num_processes, is_main_process = rank, world, rank == 0
        rcfg = copy.copy(cfg)
        rcfg.render_chunk_size = 65536      # the 34 688-ray sweep is one chunk per rank (Config default: 16384)
        enc_bytes = lambda samples: sum(b * 7 * s_ for b, s_ in zip((228, 300, 1452), samples))

        def timed_render(batches, reps, samples):
            model.num_prop_samples = tuple(samples[:2])
            with torch.no_grad():
                for rb in batches[:1] + batches[-1:]:                                   # warm-up: graph capture
                    models.render_image(model, _Acc, rb, False, rcfg, image=False, verbose=False)
                barrier()
                ts = []
                for _ in range(reps):       # median of the repetitions: one pass is tens of host-driven graph replays
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for rb in batches:
                        out = models.render_image(model, _Acc, rb, False, rcfg, image=False, verbose=False)
                    e1.record()
                    barrier()
                    ts.append(e0.elapsed_time(e1))
                ts.sort()
            t = torch.tensor([ts[len(ts) // 2]], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            model.num_prop_samples = (SAMPLES[0], SAMPLES[1])
            return float(t.item()), sorted(k for k in out if not k.startswith('ray_'))

        to_dev = lambda b: {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
        sweep = to_dev(synthetic.make_lidar_sweep(seed=0))
        frames = [to_dev(synthetic.make_camera_frame(seed=s_)) for s_ in range(4)]
        n_sweep, n_frame = sweep['origins'].shape[0], frames[0]['origins'].shape[0]
        render = {}
        for name, batches, reps, samples in (
                ('lidar_sweep_32x1084', [sweep], 5, SAMPLES),
                ('camera_frame_1600x900', frames[:1], 3, SAMPLES),
                ('sensor_fusion_4x1600x900_plus_sweep', frames + [sweep], 1, (256, 64, 32))):
            n_rays = sum(b['origins'].shape[0] for b in batches)
            ms_r, outs = timed_render(batches, reps, samples)
            render[name] = {'rays': n_rays, 'ms': ms_r, 'rays_per_s': n_rays / ms_r * 1e3, 'samples': list(samples),
                            'chunk_rays': rcfg.render_chunk_size, 'outputs': outs,
                            'encode_alg_bytes_per_ray': enc_bytes(samples)}
        del sweep, frames
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
                            '(traffic = DRAM bytes per launch from the committed ncu --set full capture).  NOT an HBM '
                            'efficiency: the scatter kernels are bound by the L2 reduction (RED) request rate -- the '
                            'hashed levels issue the 0.72 lane-reductions per clock per SM the probe measured '
                            '(profiles/r1_s15_gather_probe.txt, DESIGN.md section 8)'}
        if roofline['traffic']:
            roofline['dram_gbs'] = roofline['traffic'] / (per_kernel[dom]['avg_ms'] * 1e-3) / 1e9
            roofline['dram_frac_of_peak'] = roofline['dram_gbs'] / hbm_peak
    # the only dense contraction on the path: fused NerfMLP kernels against the measured sustained bf16 rate
    roofline_mlp = None
    tf_peak = float(peaks.get('bf16_tflops_sustained', 1356.8))
    rows_mlp = rays_model * SAMPLES[-1]
    if 'nerf_mlp_wgrad' in per_kernel:   # HBM-bound: 5.6 KB of bf16 operands per row (csrc/nerf_wgrad.cu)
        per_kernel['nerf_mlp_wgrad']['alg_gbs'] = 5600.0 * rows_mlp / (per_kernel['nerf_mlp_wgrad']['avg_ms'] * 1e-3) / 1e9
    if 'mlp_grad_sums' in per_kernel:    # one pass over the 1008 bf16 columns of pre-activation gradients (csrc/reduce.cu)
        per_kernel['mlp_grad_sums']['alg_gbs'] = 2016.0 * rows_mlp / (per_kernel['mlp_grad_sums']['avg_ms'] * 1e-3) / 1e9
    for name, macs in (('nerf_mlp_fwd', 264192), ('nerf_mlp_bwd', 257024), ('nerf_mlp_wgrad', 253184)):
        if name in per_kernel:
            t = 2.0 * macs * rows_mlp / (per_kernel[name]['avg_ms'] * 1e-3) / 1e12
            per_kernel[name]['tflops'] = t
            if name == 'nerf_mlp_fwd':
                roofline_mlp = {'bound': 'tensor', 'kernel': name, 'achieved': t, 'peak': tf_peak, 'unit': 'TFLOP/s',
                                'frac': t / tf_peak, 'traffic': None,
                                'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if 'bf16_tflops_sustained'
                                in peaks else 'fallback 1356.8 TFLOP/s', 'avg_launch_ms': per_kernel[name]['avg_ms'],
                                'flops_per_row': 2 * macs, 'rows': rows_mlp}
    # forward-only roofline of the renders: hash-grid algorithmic bytes per ray against the measured HBM rate
    if render:
        for r in render.values():
            roof = hbm_peak * 1e9 / r['encode_alg_bytes_per_ray'] * world
            r['encode_roofline_rays_per_s'] = roof
            r['frac_of_encode_roofline'] = r['rays_per_s'] / roof
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # 4096 rays, training step (forward + backward + Adam) and forward only (render),
        # the oracle port on all host cores; one timed iteration each after a small warm-up (~15-20 s of CPU work)
        cpu_reference_steps(1, 0, 256)
        rps, med, cores, threads = cpu_reference_steps(1, 0, 4096)
        rps_f, med_f, _, _ = cpu_reference_steps(1, 0, 4096, forward_only=True)
        cpu = {'value': rps, 'unit': 'rays/s', 'cores': threads, 'kind': 'port',
               'sample': f'one 4096-ray (+1024 LiDAR) training step of the oracle port ({med:.1f} s) on {cores} host '
                         f'cores; forward only (render): {rps_f:.0f} rays/s ({med_f:.1f} s)',
               'render_value': rps_f}
    ref_kernel = None
    if world == 1 and not args.no_reference_kernel:
        ref_kernel = reference_kernel_leg(model, resident[0], per_kernel)
    extras = stream_kernel_leg(dev, hbm_peak) if world == 1 else None
    stage3 = stage3_leg(dev) if world == 1 and not args.no_render else None
    value = BATCH * world * args.steps / (ms * 1e-3)
    e2e = BATCH * world * args.steps / (ms_e2e * 1e-3)
    if roofline is not None:
        # (nested so that the driver, which keeps `roofline` and `config`, records them)
        roofline['mlp'] = roofline_mlp
        roofline['reference_kernel'] = ref_kernel
        roofline['kernels'] = {k: {'avg_ms': round(v['avg_ms'], 4), 'launches_per_step': v['launches'] // n_prof,
                                   'share_of_step': round(v['share_of_step'], 4),
                                   **({'alg_gbs': round(v['alg_gbs'], 1)} if v.get('alg_gbs') else {}),
                                   **({'tflops': round(v['tflops'], 1)} if v.get('tflops') else {})}
                               for k, v in per_kernel.items()}
    config = workload_config(world)
    config.update({'launch': 'eager' if args.eager else 'whole training step replayed as one CUDA graph',
                   'eager_ms_per_step': ms_eager / n_prof,
                   'l2_policy': 'inputs larger than L2: 1.24 GB of table + optimizer state streamed every step, '
                                'batches rotate over a pool of 4',
                   'render': render,
                   'pose_refine_window': pose,
                   'stage3_raydrop': stage3,
                   # bandwidth-bound kernels at 1 M rays (launch-bound at the 10 240-ray batch) and the
                   # device-resident data layer
                   'stream_kernels_1M_rays': extras})
    line = {
        'metric': 'train_rays_per_sec', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32 grid/compositing, bf16 MLP operands (fp32 accumulate)',
        'data': 'synthetic',
        'config': config,
        'clocks': clocks,
        'e2e': {'value': e2e, 'unit': 'rays/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4},
        'gpu_launches': launches,
        'roofline': roofline,
        'cpu_baseline': cpu,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def stream_kernel_leg(dev, hbm_peak, n=1 << 20):
    """Compositing (3464 B/ray at the NeRF level, 820 B/ray at a proposal level), resampling and the
    on-GPU ray generation at 1 M rays, where they are bandwidth- rather than launch-bound: algorithmic GB/s against
    the measured HBM rate.  Inputs (0.5 GB per case) exceed L2."""
    import torch
    from nerf_lidar_b200 import ops, raygen, synthetic as sy

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    g = torch.Generator(device=dev).manual_seed(0)
    out = {}
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, device=dev, generator=g), dim=-1)
    far = torch.full((n,), 8.33, device=dev)
    with torch.no_grad():
        for S, K, name, alg in ((32, 19, 'composite_nerf_level', 3464), (64, 0, 'composite_proposal_level', 820)):
            t = torch.sort(torch.rand(n, S + 1, device=dev, generator=g) * 8 + 0.05, -1).values
            dens = torch.rand(n, S, device=dev, generator=g) * 5
            rgb = torch.rand(n, S, 3, device=dev, generator=g) if K else None
            sem = torch.softmax(torch.randn(n, S, K, device=dev, generator=g), -1) if K else None
            inten = torch.rand(n, S, device=dev, generator=g) if K else None
            ms = timeit(lambda: ops.composite(dens, t, dirs, far, rgb, sem, inten, 1.0, True, True))
            out[name] = {'ms': round(ms, 4), 'alg_bytes_per_ray': alg, 'alg_gbs': round(alg * n / ms / 1e6, 1),
                         'frac_of_hbm': round(alg * n / ms / 1e6 / hbm_peak, 3)}
            del t, dens, rgb, sem, inten
        # camera rays: 12 B of pixel / camera indices in, 76 B of ray fields out (float64 arithmetic inside)
        px = torch.randint(0, sy.IMG_W, (n,), device=dev, generator=g)
        py = torch.randint(0, sy.IMG_H, (n,), device=dev, generator=g)
        cam = torch.randint(0, 64, (n,), device=dev, generator=g)
        K3 = torch.tensor([[sy.FOCAL, 0, sy.IMG_W / 2], [0, sy.FOCAL, sy.IMG_H / 2], [0, 0, 1.]], dtype=torch.float64)
        p2c = torch.linalg.inv(K3).to(dev)
        c2w = torch.eye(3, 4, dtype=torch.float64, device=dev).repeat(64, 1, 1)
        ms = timeit(lambda: raygen.pixels_to_rays(px, py, p2c, c2w, cam_idx=cam))
        out['camera_rays'] = {'ms': round(ms, 4), 'rays_per_s': round(n / ms * 1e3), 'alg_gbs': round(88 * n / ms / 1e6, 1)}
    torch.cuda.empty_cache()
    return out


def stage3_leg(dev):
    """Stage 3 on one synthetic 32 x 1024 sweep: depth filter -> range projection -> ray-drop U-Net ->
    drop selection, all on the device; the U-Net alone in both precisions, and the SAME layers evaluated by torch
    (cuDNN) the way the reference's `UNet.forward` (R/src/unet/unet_model.py:34-47) runs on a GPU."""
    import torch
    import torch.nn.functional as F
    from nerf_lidar_b200 import raydrop
    H, W = 32, 1024
    g = torch.Generator(device=dev).manual_seed(0)
    az = torch.linspace(-3.14159, 3.14159, W, device=dev).repeat(H)
    el = torch.linspace(0.186, -0.535, H, device=dev).repeat_interleave(W)
    depth = 3 + 60 * torch.rand(H * W, device=dev, generator=g)
    pts = torch.stack([torch.cos(el) * torch.cos(az), torch.cos(el) * torch.sin(az), torch.sin(el)], -1) * depth[:, None]
    sem = torch.randint(0, 19, (H * W,), device=dev, generator=g).float()
    rgb = torch.rand(H * W, 3, device=dev, generator=g)
    torch.manual_seed(0)
    net = raydrop.UNet(6, 2, bilinear=True).to(dev).eval()

    def features():
        fm = raydrop.depth_filter(pts, sem, return_mask=True, width=1, threshold=1)
        scan = raydrop.LaserScan(H=H, W=W, fov_up=10.67, fov_down=-30.67)
        scan.set_points(pts, semantic=sem, rgb=rgb)
        scan.do_range_projection()
        x = torch.cat([scan.proj_range[None], scan.proj_semantic[None], scan.proj_mask[None], scan.proj_rgb.permute(2, 0, 1)], 0)
        return fm, scan, x[None].contiguous()

    def whole():
        fm, scan, x = features()
        return raydrop.drop_rays(net(x)[0], scan, pts, sem, fm, mask_thre=0.5)

    def torch_forward(x):
        dc = lambda m, t: m.double_conv(t)
        xs = [dc(net.inc, x)]
        for d in (net.down1, net.down2, net.down3, net.down4):
            xs.append(dc(d.maxpool_conv[1], d.maxpool_conv[0](xs[-1])))
        y = xs[-1]
        for u, skip in zip((net.up1, net.up2, net.up3, net.up4), xs[-2::-1]):
            y = dc(u.conv, torch.cat([skip, u.up(y)], 1))
        return net.outc.conv(y)

    def timeit(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    with torch.no_grad():
        _, _, x = features()
        out = {'sweep': f'{H} x {W} points, 6-channel features, UNet(6, 2, bilinear=True)',
               'whole_stage_ms': round(timeit(whole), 4)}
        net.tf32 = True
        out['unet_tf32_ms'] = round(timeit(lambda: net(x)), 4)
        a = net(x)
        net.tf32 = False
        out['unet_fp32_ms'] = round(timeit(lambda: net(x)), 4)
        out['tf32_vs_fp32_max_abs_logit'] = float((a - net(x)).abs().max())
        net.tf32 = None
        old = torch.backends.cudnn.allow_tf32
        try:
            torch.backends.cudnn.allow_tf32 = True
            out['torch_cudnn_tf32_ms'] = round(timeit(lambda: torch_forward(x)), 4)
            torch.backends.cudnn.allow_tf32 = False
            out['torch_cudnn_fp32_ms'] = round(timeit(lambda: torch_forward(x)), 4)
        finally:
            torch.backends.cudnn.allow_tf32 = old
        out['points_per_s'] = round(H * W / out['whole_stage_ms'] * 1e3)
    torch.cuda.empty_cache()
    return out


def reference_kernel_leg(model, batch, per_kernel):
    """The kernel to beat for subsystem (1): the reference's own gridencoder.cu, compiled
    unmodified into oracle/_ref (oracle/build_ref.py), timed on this workload's sample points as the chain the
    fused kernels replace -- forward: kernel_grid `[L,B,C]` + the permute copy (Z/gridencoder/grid.py:54-57) + erf
    re-weighting and multisample mean (Z/internal/models.py:974-977); backward: autograd of that mean, the permute
    copy, zeros_like(embeddings) and kernel_grid_backward (Z/gridencoder/grid.py:65-89).  The fused kernels'
    times beside it also contain cast_rays + contract (and the PropMLP on the proposal levels).  A baseline leg:
    nothing of it is on the product path."""
    import torch
    try:
        from oracle import ref_grid
        if not ref_grid.available():
            return {'unavailable': 'oracle/_ref/_gridencoder_ref.so not built (python oracle/build_ref.py)'}
        ref_grid.backend()
    except Exception as e:   # noqa: BLE001
        return {'unavailable': f'{type(e).__name__}: {e}'[:200]}
    from nerf_lidar_b200 import ops
    rays = ops.RayBundle(batch)
    out = {}
    with torch.no_grad():
        for name, mlp, S in (('prop6', model.prop_mlp_0, SAMPLES[0]), ('prop8', model.prop_mlp_1, SAMPLES[1]),
                             ('nerf_encode', model.nerf_mlp, SAMPLES[2])):
            enc = mlp.encoder
            L, C = enc.num_levels, enc.level_dim
            _, tdist = ops.resample_level(None, None, batch['near'], batch['far'], S, False, 0.5, 1.0, None, False)
            pts = ops.sample_points(tdist, None, rays)
            x = pts[..., :3].reshape(-1, 3).contiguous()
            stds = pts[..., 3].reshape(-1, 7).contiguous()
            del pts
            Mrows = stds.shape[0]
            g = torch.randn(Mrows, L * C, device=x.device)
            emb = enc.embeddings.detach()

            def fwd():
                f, _ = ref_grid.encode_forward(x, emb, enc.offsets, enc.per_level_scale, enc.base_resolution)
                return ref_grid.erf_mean(f.reshape(Mrows, 7, L * C), stds, enc.grid_sizes, L)

            def bwd():
                w = torch.erf(1 / torch.clamp(torch.sqrt(8 * stds[..., None] ** 2 * enc.grid_sizes ** 2), min=1e-10))
                g7 = (g.view(Mrows, 1, L, C) * w[..., None] / 7).reshape(Mrows * 7, L * C)
                return ref_grid.encode_backward(g7, x, emb, enc.offsets, enc.per_level_scale, enc.base_resolution)[0]

            res = {'points': int(x.shape[0])}
            for tag, fn in (('fwd', fwd), ('bwd', bwd)):
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ref_ms = e0.elapsed_time(e1) / 3
                ours = per_kernel.get(f'{name}_{tag}', {}).get('avg_ms')
                res[f'reference_{tag}_ms'] = round(ref_ms, 4)
                res[f'fused_{tag}_ms'] = round(ours, 4) if ours else None
                res[f'speedup_{tag}'] = round(ref_ms / ours, 2) if ours else None
            out[name] = res
            del x, stds, g
            torch.cuda.empty_cache()
    return out


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
    ap.add_argument('--no-pose-window', action='store_true', help='skip the pose-refinement-window step timing')
    ap.add_argument('--no-reference-kernel', action='store_true', help='skip timing the reference grid kernel (oracle/_ref)')
    ap.add_argument('--eager', action='store_true', help='issue the step eagerly instead of replaying a CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == '__main__':
    main()
