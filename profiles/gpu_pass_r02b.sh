#!/bin/bash
# r02b: tests of the new kernels (two-level list ordering, sparse Fragments), bench, host profiles, in-process vs
# stand-alone timing of the other configs, walk counters.
tag=r02b
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
tail -5 $out/tests_$tag.log
timeout 900 python bench.py --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --no-cpu --no-c5 --no-graph > $out/bench_${tag}_nograph.json 2>> $out/bench_$tag.err; echo "bench nograph rc=$?"
for c in C1 C3 C4 pose_step C5; do
  timeout 300 python profiles/run_config.py $c 20 > $out/standalone_${c}_$tag.json 2>> $out/standalone_$tag.err
done
for c in C1 C3 pose_step; do
  timeout 300 python profiles/host_profile.py $c > $out/host_${c}_$tag.txt 2>&1
done
for c in C3 C5; do
  TRB_LIB_PATH=$PWD/torch_renderer_b200/libtrb_stats.so timeout 300 python profiles/kn_stats.py $c > $out/kn_stats_${c}_$tag.json 2>> $out/kn_stats_$tag.err
done
cat $out/bench_$tag.json | cut -c1-400
