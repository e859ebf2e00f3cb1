#!/bin/bash
# r02x: tests; A/B/C of the K=1 item loop: prev (flat), base (+ per-face constants hoisted), twophase (+ dense exact evaluation).
tag=r02x
out=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
tail -3 $out/tests_$tag.log
for lib in base prev twophase base prev twophase; do
  p=$PWD/torch_renderer_b200/libtrb_$lib.so; [ $lib = base ] && p=$PWD/torch_renderer_b200/libtrb.so
  TRB_LIB_PATH=$p timeout 300 python bench.py --no-cpu --no-configs --no-c5 > $out/ab_${lib}_bench_$tag.json 2>> $out/ab_$tag.err
  python -c "
import json
try:
    d = json.load(open('$out/ab_${lib}_bench_$tag.json')); print('$lib C2', d['ms_per_step'], d['roofline']['kernels_ms_per_launch'], 'e2e', d['e2e']['ms_per_step'])
except Exception as e: print('$lib bench failed', e)"
done
for lib in base prev twophase; do
  p=$PWD/torch_renderer_b200/libtrb_$lib.so; [ $lib = base ] && p=$PWD/torch_renderer_b200/libtrb.so
  for c in C1 C4; do
    TRB_LIB_PATH=$p timeout 300 python profiles/run_config.py $c 50 > $out/ab_${lib}_${c}_$tag.json 2>> $out/ab_$tag.err
    python -c "
import json
try:
    d = json.load(open('$out/ab_${lib}_${c}_$tag.json')); print('$lib $c', 'step', d['ms_per_step_device'], 'fine', d['fine_kernel_ms'], 'bwd', d['backward_kernel_ms'])
except Exception as e: print('$lib $c failed', e)"
  done
done
