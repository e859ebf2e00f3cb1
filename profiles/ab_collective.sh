#!/bin/bash
# On a 2+ GPU box: the bench with the NCCL all-reduce (TRB_NCCL_ALLREDUCE=1) vs the peer-memory kernel, alternating.
n=${1:-2}
root=$(cd "$(dirname "$0")/.." && pwd); cd $root
show='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], "ms/step", d["ms_per_step"], "views/s", d["value"], "e2e", d["e2e"]["ms_per_step"], d["config"].get("collective"))'
for i in 1 2; do
  TRB_NCCL_ALLREDUCE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$i bench.py --gpus $n --steps 100 --no-cpu 2>/dev/null | python -c "$show" nccl
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$i bench.py --gpus $n --steps 100 --no-cpu 2>/dev/null | python -c "$show" peer
done
python bench.py --no-cpu --steps 100 2>/dev/null | python -c "$show" n1
