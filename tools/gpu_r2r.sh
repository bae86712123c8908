#!/bin/bash
# One-kernel GroupNorm v2 (every CTA merges, parallel in-CTA reduction), 8 loads in flight in gn_stats, gn_apply 4 vs 8.
mkdir -p gpurun_out
T="timeout -s KILL"
IVV_GN_FUSED=1 $T 300 python -m pytest tests/test_kernels_gpu.py tests/test_raft_gpu.py -q -x -k "groupnorm or raft or norm" > gpurun_out/t_gn.log 2>&1; echo "gn fused exit=$?"; tail -2 gpurun_out/t_gn.log
$T 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "groupnorm" > gpurun_out/t_gn0.log 2>&1; echo "gn default exit=$?"; tail -1 gpurun_out/t_gn0.log
$T 400 python tools/gn_ab.py "" IVV_GN_FUSED=1 IVV_LIB_PATH=$PWD/insv2v_b200/libivv_b200_gnapply8.so > gpurun_out/gn_ab2.txt 2>&1; grep -v Warn gpurun_out/gn_ab2.txt
IVV_GN_FUSED=1 $T 600 python -m pytest tests/test_models_gpu.py -q -x > gpurun_out/t_models_gn.log 2>&1; echo "models gn fused exit=$?"; tail -1 gpurun_out/t_models_gn.log
for st in "IVV_X=0" "IVV_GN_FUSED=1" "IVV_GN_FUSED=1 IVV_LIB_PATH=$PWD/insv2v_b200/libivv_b200_gnapply8.so"; do
  tag=$(echo "$st" | tr ' /=' '___' | tail -c 30)
  env $st $T 300 python tools/graph_timeline.py gpurun_out/tl_$tag.json > gpurun_out/tl_$tag.txt 2>&1; echo "$st" | tail -c 50; sed -n 4p gpurun_out/tl_$tag.txt
  python - <<PY
import re
t=0
for l in open('gpurun_out/tl_$tag.txt'):
    if 'groupnorm' in l:
        m=re.search(r"\)\s+(\d+)\s+([\d.]+)\s+([\d.]+)",l); t+=float(m.group(2))
print('  groupnorm total ms', round(t,3))
PY
done
for st in "IVV_X=0" "IVV_GN_FUSED=1"; do
  env $st $T 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench2_$st.json 2> gpurun_out/bench2_$st.err
  python -c "
import json
try:
    j=json.load(open('gpurun_out/bench2_$st.json')); print('$st', j['value'], j['ms_per_step'], j['gpu_launches'], j['clocks'])
except Exception as e: print('$st', 'FAILED', e)"
done
