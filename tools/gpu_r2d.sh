#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
for args in "73728 320 320 1" "73728 320 320 0" "73728 320 960 0" "18432 640 640 1" "4608 1280 1280 1"; do
  timeout -s KILL 120 python tools/gemm_trace.py $args >> gpurun_out/gemm_trace_p160.txt 2>&1
done
cat gpurun_out/gemm_trace_p160.txt
for skip in 0 1 3 4 5; do
  IVV_DEBUG_SKIP=$skip GEMM_AB_ONLY=linear timeout -s KILL 200 python tools/gemm_ab.py IVV_EPI2=1 2>&1 | sed "s/^/skip=$skip /" >> gpurun_out/gemm_knock_p160.txt
done
cat gpurun_out/gemm_knock_p160.txt
