#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/gemm_trace_ln.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for v in "LN=0" "LN=1" "LN=1 PE=1"; do
  env $v timeout -s KILL 120 python tools/gemm_trace.py 73728 320 960 0 >> gpurun_out/gemm_trace_ln.txt 2>&1
done
cat gpurun_out/gemm_trace_ln.txt | cut -c1-175
