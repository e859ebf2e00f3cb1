#!/bin/bash
# On the GPU box: the bench of the working tree with each library build in .ab_prev/variants/*.so (and the default).
root=$(cd "$(dirname "$0")/.." && pwd); cd $root
show='import json,sys
d=json.loads(sys.stdin.read()); r=d["roofline"]
print(sys.argv[1], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "fine", r["kernels_ms_per_launch"]["render_fine_kernel"], "bwd", r["kernels_ms_per_launch"]["render_backward_kernel"])'
for rep in 1 2; do
  python bench.py --no-cpu --steps 50 2>/dev/null | python -c "$show" default
  for so in .ab_prev/variants/*.so; do
    TRB_LIB_PATH=$root/$so python bench.py --no-cpu --steps 50 2>/dev/null | python -c "$show" $(basename $so)
  done
done
