#!/bin/bash
# 320-wide pair tiles (N = 1280 layers of the 8x12 level): correctness, per-shape A/B against IVV_WIDE=0, in-graph timeline A/B.
# Also re-checks GroupNorm after the gn_stats_finish refactor.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 400 python -m pytest tests/test_kernels_gpu.py -q -x -k "wide or conv3x3 or linear or groupnorm" > gpurun_out/t_v.log 2>&1; echo "tests exit=$?"; tail -15 gpurun_out/t_v.log
GEMM_AB_ONLY=wide $T 300 python tools/gemm_ab.py IVV_WIDE=0 IVV_X=0 > gpurun_out/wide_ab.txt 2>&1; grep -v Warn gpurun_out/wide_ab.txt
for st in "IVV_WIDE=0" "IVV_X=0"; do
  env $st $T 300 python tools/graph_timeline.py gpurun_out/tl4_$st.json > gpurun_out/tl4_$st.txt 2>&1; echo "$st"; sed -n 4p gpurun_out/tl4_$st.txt
  grep -E "4608, (5760|11520|17280|23040|5120), 1280|18432, 2560, 640" gpurun_out/tl4_$st.txt | cut -c1-110
done
