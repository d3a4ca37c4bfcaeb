#!/bin/bash
# Dev tool (GPU box): A/B timing of development switches in one gpurun call.
#   tools/ab_sweep.sh "name ENV=1 ..." "name2 ENV2=..." > gpurun_out/ab_sweep.txt      (no arguments: the default build twice)
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu-baseline --no-render --no-pose-window --no-reference-kernel --steps 20 --warmup 5"
line() { python - "$1" <<'P'
import json, sys
for l in open(sys.argv[1]):
    l = l.strip()
    if l.startswith('{') and '"metric"' in l:
        d = json.loads(l)
        k = d['roofline']['kernels']
        pick = {n: k[n]['avg_ms'] for n in ('adam_table', 'mlp_grad_sums', 'nerf_encode_bwd', 'prop8_bwd', 'prop6_bwd', 'nerf_encode_fwd', 'prop8_fwd', 'prop6_fwd', 'nerf_mlp_fwd', 'nerf_mlp_bwd', 'nerf_mlp_wgrad') if n in k}
        print('   ms_per_step %.4f  e2e %.4f  eager %.3f  launches %d  %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['eager_ms_per_step'], d['gpu_launches'], pick))
P
}
run() { echo "== $1"; shift; env "$@" $B > /tmp/ab.log 2>&1 || tail -5 /tmp/ab.log; line /tmp/ab.log; }
if [ $# -eq 0 ]; then set -- "default X=1" "default-again X=1"; fi
for spec in "$@"; do run $spec; done
