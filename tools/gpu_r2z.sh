#!/bin/bash
# 320-wide tiles, second version (register residual + two slabs per group; competes with the short-K pair kernel from K = 1280):
# correctness, per-shape A/B against the K >= 2560 rule, in-graph timeline A/B, whole-model parity.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 500 python -m pytest tests/test_kernels_gpu.py -q -k "wide or conv3x3 or linear" > gpurun_out/t_z.log 2>&1; echo "tests exit=$?"; tail -8 gpurun_out/t_z.log
GEMM_AB_ONLY=wide $T 300 python tools/gemm_ab.py IVV_WIDE=0 IVV_X=0 > gpurun_out/wide2_ab.txt 2>&1; grep -v Warn gpurun_out/wide2_ab.txt
$T 300 python tools/linear_ab.py IVV_WIDE_K=2560 IVV_X=0 > gpurun_out/wide2_lin.txt 2>&1; grep -E "res2|ffout|sum" gpurun_out/wide2_lin.txt
for st in "IVV_WIDE_K=2560" "IVV_X=0"; do
  env $st $T 300 python tools/graph_timeline.py gpurun_out/tl7_$st.json > gpurun_out/tl7_$st.txt 2>&1; echo "$st"; sed -n 4p gpurun_out/tl7_$st.txt
  grep -E "4608, (5760|11520|17280|23040|5120|1280), 1280|18432, 2560, 640|73728, 1280, 320" gpurun_out/tl7_$st.txt | cut -c1-110
done
$T 600 python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -x > gpurun_out/t_z_models.log 2>&1; echo "models exit=$?"; tail -3 gpurun_out/t_z_models.log
