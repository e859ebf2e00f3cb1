#!/bin/bash
# ncu evidence at HEAD: launch lists (share of the step per kernel) and full captures of the dominant kernels, C2 and C5.
# Every ncu command runs only after the same command exited 0 without ncu.
tag=${1:-r04n}
out=gpurun_out
mkdir -p $out
C2="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-configs --no-c5"
timeout 600 $C2 > $out/plain_c2_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_c2_$tag.csv \
  $C2 > $out/ncu_launch_c2_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_fine_k1_kernel|render_backward_kernel' -s 6 -c 2 \
  -o $out/prof_c2_$tag -f $C2 > $out/ncu_full_c2_$tag.log 2>&1
for c in C5; do
  timeout 600 python profiles/run_config.py $c 2 > $out/plain_${c}_$tag.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${c}_$tag.csv \
    python profiles/run_config.py $c 2 > $out/ncu_launch_${c}_$tag.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_fine_kn_kernel|render_backward_kernel' -s 6 -c 2 \
    -o $out/prof_${c}_$tag -f python profiles/run_config.py $c 2 > $out/ncu_full_${c}_$tag.log 2>&1
done
ls -la $out/*$tag*
