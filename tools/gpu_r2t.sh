#!/bin/bash
# Persistent one-tile attention (d = 80 / 160), IVV_ATTN_PERSIST1=1: correctness, per-shape A/B, timeline, bench A/B.
mkdir -p gpurun_out
T="timeout -s KILL"
IVV_ATTN_PERSIST1=1 $T 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > gpurun_out/t_attn_p1.log 2>&1; echo "attn persist1 exit=$?"; tail -3 gpurun_out/t_attn_p1.log
$T 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > gpurun_out/t_attn_def.log 2>&1; echo "attn default exit=$?"; tail -1 gpurun_out/t_attn_def.log
ATTN_BENCH_SMALL=1 $T 400 python tools/attn_bench.py > gpurun_out/attn_small.txt 2>&1; grep -v Warn gpurun_out/attn_small.txt
IVV_ATTN_PERSIST1=1 $T 600 python -m pytest tests/test_models_gpu.py -q -x > gpurun_out/t_models_p1.log 2>&1; echo "models persist1 exit=$?"; tail -1 gpurun_out/t_models_p1.log
for st in "IVV_X=0" "IVV_ATTN_PERSIST1=1"; do
  env $st $T 300 python tools/graph_timeline.py gpurun_out/tl2_$st.json > gpurun_out/tl2_$st.txt 2>&1; echo "$st"; sed -n 4p gpurun_out/tl2_$st.txt; grep "'attention'" gpurun_out/tl2_$st.txt | cut -c1-105
done
for st in "IVV_X=0" "IVV_ATTN_PERSIST1=1"; do
  env $st $T 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench4_$st.json 2> gpurun_out/bench4_$st.err
  python -c "
import json
try:
    j=json.load(open('gpurun_out/bench4_$st.json')); print('$st', j['value'], j['ms_per_step'], j['gpu_launches'], j['clocks'])
except Exception as e: print('$st', 'FAILED', e)"
done
