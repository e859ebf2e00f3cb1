#!/bin/bash
# r02h: tests, points bench, bench (moved bytes), then the round's ncu evidence: launch lists (warm-up skipped) and
# full captures of the K>1 fine + backward kernels on C5 / C3 and of the C2 kernels.
tag=r02h
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
tail -3 $out/tests_$tag.log
timeout 300 python profiles/points_bench.py > $out/points_bench_$tag.json 2> $out/points_bench_$tag.err; cut -c1-400 $out/points_bench_$tag.json
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
for c in C5 C3; do
  timeout 600 python profiles/run_config.py $c 2 > $out/plain_${c}_$tag.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${c}_$tag.csv \
    python profiles/run_config.py $c 2 > $out/ncu_launch_${c}_$tag.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_fine_kn_kernel|render_backward_kernel|silhouette_backward' -s 6 -c 2 \
    -o $out/prof_${c}_$tag -f python profiles/run_config.py $c 2 > $out/ncu_full_${c}_$tag.log 2>&1
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-configs --no-c5 > $out/plain_c2_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_c2_$tag.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-configs --no-c5 > $out/ncu_launch_c2_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_fine_k1_kernel|render_backward_kernel' -s 6 -c 2 \
  -o $out/prof_c2_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu --no-graph --no-configs --no-c5 > $out/ncu_full_c2_$tag.log 2>&1
cut -c1-300 $out/bench_$tag.json
