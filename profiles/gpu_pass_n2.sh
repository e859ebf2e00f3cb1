#!/bin/bash
# Two-GPU check at HEAD: the 2-GPU legs of tests/test_parallel.py and the bench line at N=2 (fused and separate all-reduce).
tag=${1:-n2}
out=gpurun_out
timeout 900 python -m pytest tests/test_parallel.py -m gpu -x -q > $out/tests_parallel_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_parallel_$tag.log
tail -5 $out/tests_parallel_$tag.log
for mode in fused separate; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --collective $mode --no-c5 > $out/bench_n2_${mode}_$tag.json 2> $out/bench_n2_${mode}_$tag.err; echo "bench $mode rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('$out/bench_n2_${mode}_$tag.json').read().strip().splitlines()[-1])
    print('$mode', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['config']['launch_mode'])
    print(' check', d.get('collective_check'), '|', d.get('collective_check_fused'))
    print(' ab', d.get('collective_ab'))
    t = d.get('collective_timing') or {}
    print(' timing', t.get('push_us'), t.get('wait_and_sum_us'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 $out/bench_n2_${mode}_$tag.err
done
