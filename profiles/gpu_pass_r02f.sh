#!/bin/bash
# r02f: full GPU tests; A/B of two threads per pixel on 8x8 tiles (C3); bench.
tag=r02f
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
tail -5 $out/tests_$tag.log
for lib in base split1; do
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
TRB_LIB_PATH=$PWD/torch_renderer_b200/libtrb_stats.so timeout 200 python profiles/kn_stats.py C3 | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ('busy_tiles','max_list_entries','max_walk_cycles','full_evals','insertions')})"
timeout 900 python bench.py --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
cut -c1-300 $out/bench_$tag.json
