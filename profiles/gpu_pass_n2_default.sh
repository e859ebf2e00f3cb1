#!/bin/bash
# The driver's N=2 command, verbatim flags (C5 at spec included), plus the reference arm under torchrun.
tag=${1:-n2d}
out=gpurun_out
t0=$SECONDS; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${2:-2} --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus ${2:-2} --steps 20 --warmup 5 > $out/bench_n2_$tag.json 2> $out/bench_n2_$tag.err; echo "bench rc=$?"
echo "wall $((SECONDS - t0)) s"
python - <<PY
import json
d = json.loads(open('$out/bench_n2_$tag.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d.get('collective_check'))
print('c5', d['c5']['views_per_s'], d['c5']['seconds_per_pass'])
print((d.get('collective_timing') or {}).get('rank_skew'))
print(len(open('$out/bench_n2_$tag.json').read().strip().splitlines()), 'line(s) on stdout')
PY
