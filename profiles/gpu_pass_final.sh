#!/bin/bash
# Round-end style pass: build check, full GPU tests, smoke(), bench (all blocks + cpu baseline), reference arm, points.
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
tail -3 $out/tests_$tag.log
timeout 300 python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_$tag.log
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err; echo "ref rc=$?"
timeout 300 python profiles/points_bench.py > $out/points_bench_$tag.json 2> $out/points_bench_$tag.err
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "eager", d["eager_exact"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["moved"]["frac"])
for k, v in d["other_configs"].items():
    print(k, v.get("ms_per_step"), v.get("fine_kernel_ms"), v.get("backward_kernel_ms"), (v.get("captured") or {}).get("ms_per_step"), v.get("error"))
print("c5", d["c5"]["views_per_s"], d["c5"]["seconds_per_pass"], "cpu", d.get("cpu_baseline"))
PY
cut -c1-300 $out/points_bench_$tag.json
