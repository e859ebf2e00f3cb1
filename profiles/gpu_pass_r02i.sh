#!/bin/bash
# r02i: A/B of the out-of-line scatter in the K>1 backward (C5, C3cow phong n/a); first-render time of C5 (capacity
# estimate); quick K>1 parity subset.
tag=r02i
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log
tail -3 $out/tests_$tag.log
for lib in base inl; do
  p=$PWD/torch_renderer_b200/libtrb_$lib.so; [ $lib = base ] && p=$PWD/torch_renderer_b200/libtrb.so
  TRB_LIB_PATH=$p timeout 300 python profiles/run_config.py C5 10 > $out/ab_${lib}_C5_$tag.json 2>> $out/ab_$tag.err
  python -c "
import json
try:
    d = json.load(open('$out/ab_${lib}_C5_$tag.json')); print('$lib C5', 'step', d['ms_per_step_device'], 'fine', d['fine_kernel_ms'], 'bwd', d['backward_kernel_ms'])
except Exception as e: print('$lib failed', e)"
done
python - <<PY
import sys, time, torch
sys.path.insert(0, "profiles")
import configs
dev = torch.device("cuda:0")
step, info = configs.BUILDERS["C5"](dev)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize()
    print("C5 render", i, round((time.perf_counter() - t0) * 1e3, 1), "ms")
PY
