#!/bin/bash
# r02e: full GPU tests; same-box A/B (K=1 CTAs per SM on C2 / C1, 8x8-tile parking depth on C3); bench.
tag=r02e
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
tail -5 $out/tests_$tag.log
for lib in base kg8old kg8_3; do
  p=$PWD/torch_renderer_b200/libtrb_$lib.so; [ $lib = base ] && p=$PWD/torch_renderer_b200/libtrb.so
  for c in C3 C3cow; do
    TRB_LIB_PATH=$p timeout 300 python profiles/run_config.py $c 20 > $out/ab_${lib}_${c}_$tag.json 2>> $out/ab_$tag.err
    python -c "
import json
try:
    d = json.load(open('$out/ab_${lib}_${c}_$tag.json')); print('$lib $c', 'step', d['ms_per_step_device'], 'fine', d['fine_kernel_ms'], 'bwd', d['backward_kernel_ms'])
except Exception as e: print('$lib $c failed', e)"
  done
done
for lib in base k1c5 k1c6; do
  p=$PWD/torch_renderer_b200/libtrb_$lib.so; [ $lib = base ] && p=$PWD/torch_renderer_b200/libtrb.so
  TRB_LIB_PATH=$p timeout 300 python bench.py --no-cpu --no-configs --no-c5 > $out/ab_${lib}_bench_$tag.json 2>> $out/ab_$tag.err
  python -c "
import json
try:
    d = json.load(open('$out/ab_${lib}_bench_$tag.json')); print('$lib C2', d['ms_per_step'], d['roofline']['kernels_ms_per_launch'], 'eager', d['eager_exact']['ms_per_step'])
except Exception as e: print('$lib bench failed', e)"
done
timeout 900 python bench.py --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
cut -c1-300 $out/bench_$tag.json
