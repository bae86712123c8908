#!/bin/bash
# Activation-stationary modes: L2 prefetch of the next M pair (IVV_AS_PF=0 disables), AS from two N tiles (IVV_AS=2).
mkdir -p gpurun_out
T="timeout -s KILL"
$T 300 python -m pytest tests/test_kernels_gpu.py -q -k "activation_stationary or geglu or layernorm_folded" > gpurun_out/t_y.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/t_y.log
$T 300 python tools/linear_ab.py IVV_AS_PF=0 IVV_X=0 IVV_AS=2 > gpurun_out/as_pf_ab.txt 2>&1; grep -E "qkv0|q0|res0|geglu0|sum" gpurun_out/as_pf_ab.txt
for st in "IVV_AS_PF=0" "IVV_X=0" "IVV_AS=2"; do
  env $st $T 300 python tools/graph_timeline.py gpurun_out/tl6_$st.json > gpurun_out/tl6_$st.txt 2>&1; echo "$st"; sed -n 4p gpurun_out/tl6_$st.txt
  grep -E "73728, 320, (960|2560|320)," gpurun_out/tl6_$st.txt | cut -c1-110
done
