#!/bin/bash
# r02g: full GPU tests (binned point rasteriser), points bench, bench (with the clipped-route config).
tag=r02g
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
tail -5 $out/tests_$tag.log
timeout 300 python profiles/points_bench.py > $out/points_bench_$tag.json 2> $out/points_bench_$tag.err; cat $out/points_bench_$tag.json
timeout 900 python bench.py --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["kernels_ms_per_launch"])
for k, v in d["other_configs"].items():
    print(k, v.get("ms_per_step"), v.get("fine_kernel_ms"), v.get("backward_kernel_ms"), (v.get("captured") or {}).get("ms_per_step"), v.get("error"))
print(d["c5"]["views_per_s"], d["c5"]["seconds_per_pass"])
PY
