#!/bin/bash
# N-GPU bench line with the rank-skew report (collective_timing.rank_skew).
tag=${1:-skew}; n=${2:-2}
out=gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $n --no-c5 > $out/bench_n${n}_$tag.json 2> $out/bench_n${n}_$tag.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('$out/bench_n${n}_$tag.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['config']['launch_mode'])
    print(' check', d.get('collective_check'))
    t = d.get('collective_timing') or {}
    print(' timing', t.get('push_us'), t.get('wait_and_sum_us'))
    print(' skew', t.get('rank_skew'), t.get('cuda_contexts'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 $out/bench_n${n}_$tag.err | grep -v "^\*\*\*\|OMP"
