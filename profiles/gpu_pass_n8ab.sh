#!/bin/bash
# N-GPU A/B of the all-reduce variants inside one process (bench.py's collective_ab): fused push vs separate kernel.
tag=${1:-n8ab}; n=${2:-8}
out=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $n --no-c5 > $out/bench_n${n}_$tag.json 2> $out/bench_n${n}_$tag.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('$out/bench_n${n}_$tag.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['config']['launch_mode'])
    print(' check', d.get('collective_check'), '|', d.get('collective_check_fused'))
    print(' ab', d.get('collective_ab'))
    t = d.get('collective_timing') or {}
    print(' timing', t.get('push_us'), t.get('wait_and_sum_us'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 $out/bench_n${n}_$tag.err
timeout 300 python bench.py --no-c5 --no-configs --no-cpu > $out/bench_n1_$tag.json 2>> $out/bench_n${n}_$tag.err
python -c "
import json; d=json.loads(open('$out/bench_n1_$tag.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
