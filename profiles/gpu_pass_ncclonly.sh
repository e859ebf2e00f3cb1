#!/bin/bash
# Is a rank's OWN step (no all-reduce in it) slower under torchrun because of NCCL's set-up or because of the
# symmetric-memory inbox?  Same box: N=1 alone; N=2 with the peer kernel; N=2 with TRB_NCCL_ALLREDUCE=1 (no inbox).
tag=${1:-no}
out=gpurun_out
timeout 300 python bench.py --no-c5 --no-configs --no-cpu > $out/bench_no_n1_$tag.json 2> $out/bench_no_$tag.err
for mode in peer nccl; do
  [ $mode = nccl ] && export TRB_NCCL_ALLREDUCE=1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --no-c5 > $out/bench_no_${mode}_$tag.json 2>> $out/bench_no_$tag.err; echo "$mode rc=$?"
done
python -c "
import json
for f in ('n1', 'peer', 'nccl'):
    d = json.loads(open('$out/bench_no_' + f + '_$tag.json').read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['config'].get('collective'), d['config']['launch_mode'], (d.get('collective_timing') or {}).get('rank_skew'), d['roofline']['kernels_ms_per_launch'])
"
tail -5 $out/bench_no_$tag.err | grep -v "^\*\*\*\|OMP"
