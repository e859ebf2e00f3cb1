#!/bin/bash
# On the GPU box: alternates the bench of .ab_prev (A) and of the working tree (B), `reps` times each.
reps=${1:-2}
root=$(cd "$(dirname "$0")/.." && pwd)
for i in $(seq $reps); do
  for side in A B; do
    dir=$root; [ $side = A ] && dir=$root/.ab_prev
    (cd $dir && python bench.py --no-cpu --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$side', 'ms/step', d['ms_per_step'], 'views/s', d['value'], 'e2e', d['e2e']['ms_per_step'], 'fine', r['kernels_ms_per_launch']['render_fine_kernel'], 'bwd', r['kernels_ms_per_launch']['render_backward_kernel'], 'calls', r['calls_ms'])")
  done
done
