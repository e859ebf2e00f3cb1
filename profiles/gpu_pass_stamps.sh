#!/bin/bash
# Step timeline from %globaltimer stamps (diagnostic build libtrb_stamps.so): N=1 and N=2 on the same box.
tag=${1:-st}; n=${2:-2}
out=gpurun_out
export TRB_LIB_PATH=$PWD/torch_renderer_b200/libtrb_stamps.so
export TRB_STEP_STAMPS_OUT=$out/stamps_$tag
timeout 300 python bench.py --no-c5 --no-configs --no-cpu > $out/bench_stamps_n1_$tag.json 2> $out/bench_stamps_$tag.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $n --no-c5 > $out/bench_stamps_n${n}_$tag.json 2>> $out/bench_stamps_$tag.err; echo "n$n rc=$?"
for f in $out/stamps_${tag}_*.json; do echo $f; cat $f; echo; done
python -c "
import json
for f in ('$out/bench_stamps_n1_$tag.json', '$out/bench_stamps_n${n}_$tag.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], (d.get('collective_timing') or {}).get('rank_skew'))
"
