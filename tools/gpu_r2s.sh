#!/bin/bash
# LayerNorm folded into the GEGLU GEMMs (+ one-kernel GroupNorm now default): correctness, model parity, timeline, bench A/B.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 400 python -m pytest tests/test_kernels_gpu.py -q -x -s -k "geglu or layernorm or groupnorm or linear" > gpurun_out/t_k.log 2>&1; echo "kernels exit=$?"; grep "rel-L2 vs fp32" gpurun_out/t_k.log | tail -6; tail -2 gpurun_out/t_k.log
$T 900 python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -x -s > gpurun_out/t_models.log 2>&1; echo "models exit=$?"; grep -i "rel-L2\|rel_l2" gpurun_out/t_models.log | tail -12; tail -2 gpurun_out/t_models.log
for st in "IVV_GEGLU_LN=0" "IVV_X=0"; do
  env $st $T 300 python tools/graph_timeline.py gpurun_out/tl_$st.json > gpurun_out/tl_$st.txt 2>&1; echo "$st"; sed -n 4p gpurun_out/tl_$st.txt; grep "geglu\|layernorm" gpurun_out/tl_$st.txt | cut -c1-100
done
for st in "IVV_GEGLU_LN=0" "IVV_X=0"; do
  env $st $T 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench3_$st.json 2> gpurun_out/bench3_$st.err
  python -c "
import json
try:
    j=json.load(open('gpurun_out/bench3_$st.json')); print('$st', j['value'], j['ms_per_step'], j['gpu_launches'], j['clocks'])
except Exception as e: print('$st', 'FAILED', e)"
done
