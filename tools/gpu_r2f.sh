#!/bin/bash
# weight-stationary pair160 bring-up: GEMM + frame I/O parity, A/B (IVV_NO_WS=1 vs default), traces, bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "linear or conv or gemm or frame_io" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit=$?"; tail -5 gpurun_out/pytest_gemm.log
rm -f gpurun_out/gemm_ab_ws.txt
GEMM_AB_ONLY=linear timeout -s KILL 300 python tools/gemm_ab.py IVV_NO_WS=1 IVV_X=1 > gpurun_out/gemm_ab_ws.txt 2>&1; cat gpurun_out/gemm_ab_ws.txt
rm -f gpurun_out/gemm_trace_ws.txt
for args in "73728 320 320 1" "73728 320 960 0"; do
  timeout -s KILL 120 python tools/gemm_trace.py $args >> gpurun_out/gemm_trace_ws.txt 2>&1
done
head -24 gpurun_out/gemm_trace_ws.txt
timeout -s KILL 600 python -m pytest tests/test_models_gpu.py -m gpu -q -x > gpurun_out/pytest_models.log 2>&1; echo "pytest models exit=$?"; tail -3 gpurun_out/pytest_models.log
timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_ws.json 2> gpurun_out/bench_ws.err
python -c "
import json
j=json.load(open('gpurun_out/bench_ws.json')); print('bench', j['value'], j['ms_per_step'], j['clocks'])"
