#!/bin/bash
# Short-K pair kernel with ONE output slab and a sixth pipeline stage for K >= 640 (IVV_SLAB1=0 disables): correctness, A/B.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 500 python -m pytest tests/test_kernels_gpu.py -q -k "linear or layernorm_folded" > gpurun_out/t_aa.log 2>&1; echo "tests exit=$?"; tail -5 gpurun_out/t_aa.log
$T 300 python tools/linear_ab.py IVV_SLAB1=0 IVV_X=0 > gpurun_out/slab1_ab.txt 2>&1; grep -v Warn gpurun_out/slab1_ab.txt | tail -32
for st in "IVV_SLAB1=0" "IVV_X=0"; do
  env $st $T 300 python tools/graph_timeline.py gpurun_out/tl8_$st.json > gpurun_out/tl8_$st.txt 2>&1; echo "$st"; sed -n 4p gpurun_out/tl8_$st.txt
  grep -E "gemm', (18432, 640|4608, 1280|1152, 1280|73728, 1280), (640|1920|1280|3840|320)," gpurun_out/tl8_$st.txt | cut -c1-110
done
