#!/bin/bash
# Two K blocks per ring slot in the short-K pair kernel (IVV_KB2=1, opt-in): correctness in both modes, A/B, timeline.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 200 python -m pytest tests/test_kernels_gpu.py -q -k "linear or layernorm_folded" > gpurun_out/t_ac0.log 2>&1; echo "tests default exit=$?"; tail -1 gpurun_out/t_ac0.log
IVV_KB2=1 $T 200 python -m pytest tests/test_kernels_gpu.py -q -k "linear or layernorm_folded" > gpurun_out/t_ac1.log 2>&1; echo "tests KB2 exit=$?"; tail -3 gpurun_out/t_ac1.log
$T 200 python tools/linear_ab.py IVV_X=0 IVV_KB2=1 > gpurun_out/kb2_ab.txt 2>&1; grep -E "qkv1|qkv2|res1|res2|res3|ffout0|sum" gpurun_out/kb2_ab.txt
for st in "IVV_X=0" "IVV_KB2=1"; do
  env $st $T 200 python tools/graph_timeline.py gpurun_out/tl9_$st.json > gpurun_out/tl9_$st.txt 2>&1; echo "$st"; sed -n 4p gpurun_out/tl9_$st.txt
  grep -E "gemm', (18432, 640|4608, 1280|1152, 1280|73728, 1280), (640|1920|1280|3840|320)," gpurun_out/tl9_$st.txt | cut -c1-110
done
