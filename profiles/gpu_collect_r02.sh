#!/bin/bash
# Round-2 one-box pass: tests, bench (headline + other_configs + C5 at spec), reference arm, K>1 walk counters
# (diagnostic build), ncu launch lists and full captures of the K>1 fine + backward kernels on C5 / C3.
# Usage (repo root, under gpurun): bash profiles/gpu_collect_r02.sh <tag> [skip-tests]
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/smi_$tag.log 2>&1
if [ -z "$2" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
  tail -3 $out/tests_$tag.log
fi
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
if [ -f torch_renderer_b200/libtrb_stats.so ]; then
  for c in C3 C3cow C5; do
    TRB_LIB_PATH=$PWD/torch_renderer_b200/libtrb_stats.so timeout 300 python profiles/kn_stats.py $c > $out/kn_stats_${c}_$tag.json 2>> $out/kn_stats_$tag.err
  done
fi
for c in C5 C3; do
  timeout 600 python profiles/run_config.py $c 2 > $out/plain_${c}_$tag.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_${c}_$tag.csv \
    python profiles/run_config.py $c 2 > $out/ncu_launch_${c}_$tag.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_fine_kn_kernel|render_backward_kernel' -s 6 -c 2 \
    -o $out/prof_${c}_$tag -f python profiles/run_config.py $c 2 > $out/ncu_full_${c}_$tag.log 2>&1
done
cat $out/bench_$tag.json
