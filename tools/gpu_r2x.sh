#!/bin/bash
# clock64 traces of the short-K pair kernel on the QKV shape: plain ring vs activation-stationary
mkdir -p gpurun_out
T="timeout -s KILL"
for st in "IVV_AS=0" "IVV_X=0"; do
  echo "## $st"; env $st $T 200 python tools/gemm_trace.py 73728 320 960 0 2>&1 | grep -v Warn | head -30
done > gpurun_out/as_trace.txt
cat gpurun_out/as_trace.txt
