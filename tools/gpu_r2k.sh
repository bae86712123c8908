#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/gemm_trace_ln.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -s -k "layernorm_folded or linear" > gpurun_out/pytest_fold.log 2>&1; echo "pytest fold exit=$?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/pytest_fold.log | tail -8
for v in "LN=0" "LN=1" "LN=1 PE=1"; do
  env $v timeout -s KILL 120 python tools/gemm_trace.py 73728 320 960 0 2>&1 | head -12 >> gpurun_out/gemm_trace_ln.txt
done
cat gpurun_out/gemm_trace_ln.txt | cut -c1-175
echo skip-fhadd
for s in 0 1; do
  IVV_LN_FOLD=$s timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_fold_$s.json 2> gpurun_out/bench_fold_$s.err
  python -c "
import json
j=json.load(open('gpurun_out/bench_fold_$s.json')); print('LN_FOLD=$s', j['value'], j['ms_per_step'], j['gpu_launches'], j['clocks'])"
done
timeout -s KILL 600 python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/pytest_models.log 2>&1; echo "pytest models exit=$?"; tail -3 gpurun_out/pytest_models.log
