#!/bin/bash
# Activation-stationary mode (pair160 kernel: 320 -> 960 QKV; persistent GEGLU kernel: 320 -> 2560): correctness, A/B, timeline.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 500 python -m pytest tests/test_kernels_gpu.py -q -k "linear or layernorm_folded" > gpurun_out/t_w.log 2>&1; echo "tests exit=$?"; tail -12 gpurun_out/t_w.log
$T 300 python tools/linear_ab.py IVV_AS=0 IVV_X=0 > gpurun_out/as_ab.txt 2>&1; grep -v Warn gpurun_out/as_ab.txt | tail -40
for st in "IVV_AS=0" "IVV_X=0"; do
  env $st $T 300 python tools/graph_timeline.py gpurun_out/tl5_$st.json > gpurun_out/tl5_$st.txt 2>&1; echo "$st"; sed -n 4p gpurun_out/tl5_$st.txt
  grep -E "73728, 320, (960|2560|320)," gpurun_out/tl5_$st.txt | cut -c1-110
done
