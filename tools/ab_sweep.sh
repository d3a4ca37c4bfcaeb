#!/bin/bash
# Dev tool (GPU box): A/B timing of the development switches in one gpurun call.
#   tools/ab_sweep.sh > gpurun_out/ab_sweep.txt
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu-baseline --no-render --no-pose-window --no-reference-kernel --steps 20 --warmup 5"
line() { python - "$1" <<'P'
import json, sys
for l in open(sys.argv[1]):
    l = l.strip()
    if l.startswith('{') and '"metric"' in l:
        d = json.loads(l)
        k = d['roofline']['kernels']
        pick = {n: k[n]['avg_ms'] for n in ('adam_table', 'mlp_grad_sums', 'mlp_bias_grad', 'mlp_ray_sum', 'nerf_encode_bwd', 'prop8_bwd', 'prop6_bwd', 'nerf_encode_fwd', 'loss_sums', 'loss_seed') if n in k}
        print('   ms_per_step %.4f  e2e %.4f  eager %.3f  launches %d  %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['eager_ms_per_step'], d['gpu_launches'], pick))
P
}
run() { echo "== $1"; shift; env "$@" $B > /tmp/ab.log 2>&1 || tail -5 /tmp/ab.log; line /tmp/ab.log; }
run default X=1
run default-again X=1
run loss-unfused NLB_LOSS_UNFUSED=1
run adam-8-per-sm NLB_ADAM_BLOCKS_PER_SM=8
run adam-4-per-sm NLB_ADAM_BLOCKS_PER_SM=4
run scatter-4-per-sm NLB_SCATTER_BLOCKS_PER_SM=4
run scatter-3-per-sm NLB_SCATTER_BLOCKS_PER_SM=3
run pair-8-per-sm NLB_SCATTER_PAIR_BLOCKS_PER_SM=8
run scatter-l2-40 NLB_SCATTER_L2_MB=40
run gather-l2-100 NLB_GATHER_L2_MB=100
