#!/bin/bash
# One GPU-box pass: tests, bench, per-config timings, ncu launch list + full captures of the two fused kernels.
# Usage (from the repo root, under gpurun): bash profiles/gpu_collect.sh <tag>
tag=${1:-run}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/smi_$tag.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
timeout 900 python profiles/bench_configs.py > $out/configs_$tag.json 2> $out/configs_$tag.err; echo "configs rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $out/ncu_launch_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_fine_k1_kernel|render_backward_kernel' -s 6 -c 2 \
  -o $out/prof_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > $out/ncu_full_$tag.log 2>&1
tail -3 $out/tests_$tag.log; cat $out/bench_$tag.json; cat $out/configs_$tag.json
