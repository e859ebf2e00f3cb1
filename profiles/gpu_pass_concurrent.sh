#!/bin/bash
# Two INDEPENDENT single-GPU bench processes at the same time on two GPUs of one box (no torch.distributed, no peer
# memory): does a GPU's own step slow down just because its neighbour is busy?
tag=${1:-cc}
out=gpurun_out
timeout 300 python bench.py --no-c5 --no-configs --no-cpu > $out/bench_alone_$tag.json 2> $out/bench_cc_$tag.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-c5 --no-configs --no-cpu > $out/bench_cc0_$tag.json 2>> $out/bench_cc_$tag.err &
p0=$!
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --no-c5 --no-configs --no-cpu > $out/bench_cc1_$tag.json 2>> $out/bench_cc_$tag.err &
p1=$!
wait $p0; wait $p1
python -c "
import json
for f in ('bench_alone', 'bench_cc0', 'bench_cc1'):
    d = json.loads(open('$out/' + f + '_$tag.json').read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['timing']['ms_per_step_repeats'], d['roofline']['kernels_ms_per_launch'])
"
