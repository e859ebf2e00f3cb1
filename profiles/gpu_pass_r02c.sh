#!/bin/bash
# r02c: full GPU tests (sparse Fragments, capture_step, STRIP=1), same-box A/B of K>1 kernel variants, bench,
# ncu of the C2 fine kernel in sparse mode.
tag=r02c
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
tail -5 $out/tests_$tag.log
for lib in base kg4 kg4c4 bwd4 bwd5; do
  p=$PWD/torch_renderer_b200/libtrb_$lib.so; [ $lib = base ] && p=$PWD/torch_renderer_b200/libtrb.so
  for c in C5 C3; do
    TRB_LIB_PATH=$p timeout 300 python profiles/run_config.py $c 10 > $out/ab_${lib}_${c}_$tag.json 2>> $out/ab_$tag.err
    python - <<PY
import json
try:
    d = json.load(open("$out/ab_${lib}_${c}_$tag.json"))
    print("$lib $c", "step", d["ms_per_step_device"], "fine", d["fine_kernel_ms"], "bwd", d["backward_kernel_ms"])
except Exception as e:
    print("$lib $c failed", e)
PY
  done
done
timeout 900 python bench.py --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-configs --no-c5 > $out/plain_$tag.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_fine_k1_kernel|render_backward_kernel' -s 6 -c 2 \
  -o $out/prof_c2_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-configs --no-c5 > $out/ncu_full_$tag.log 2>&1
cut -c1-300 $out/bench_$tag.json
