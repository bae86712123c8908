#!/bin/bash
# Bring-up of the one-kernel GroupNorm (IVV_GN_FUSED=1): correctness, per-shape A/B, clip bench A/B, in-graph timeline.
mkdir -p gpurun_out
T="timeout -s KILL"
IVV_GN_FUSED=1 $T 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "groupnorm" > gpurun_out/t_gn.log 2>&1; echo "gn fused exit=$?"; tail -3 gpurun_out/t_gn.log
$T 120 python -m pytest tests/test_kernels_gpu.py -q -x -k "groupnorm" > gpurun_out/t_gn0.log 2>&1; echo "gn default exit=$?"; tail -1 gpurun_out/t_gn0.log
$T 400 python tools/gn_ab.py "" IVV_GN_FUSED=1 > gpurun_out/gn_ab.txt 2>&1; grep -v Warn gpurun_out/gn_ab.txt
IVV_GN_FUSED=1 $T 600 python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -x > gpurun_out/t_models_gn.log 2>&1; echo "models gn fused exit=$?"; tail -2 gpurun_out/t_models_gn.log
for st in "IVV_X=0" "IVV_GN_FUSED=1"; do
  env $st $T 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_$st.json 2> gpurun_out/bench_$st.err
  python -c "
import json
try:
    j=json.load(open('gpurun_out/bench_$st.json')); print('$st', j['value'], j['ms_per_step'], j['gpu_launches'], j['clocks'])
except Exception as e: print('$st', 'FAILED', e)"
done
$T 300 python tools/graph_timeline.py gpurun_out/graph_timeline_default.json > gpurun_out/graph_timeline_default.txt 2>&1; head -8 gpurun_out/graph_timeline_default.txt | tail -5
IVV_GN_FUSED=1 $T 300 python tools/graph_timeline.py gpurun_out/graph_timeline_gnfused.json > gpurun_out/graph_timeline_gnfused.txt 2>&1; head -8 gpurun_out/graph_timeline_gnfused.txt | tail -5
