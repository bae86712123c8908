#!/bin/bash
# Bring-up of: GEGLU bias staged in smem (+ two output slabs for K <= 320), 4-CTA clusters with activation multicast in the
# short-K pair kernel. Correctness first (each under its own timeout), then A/B timing per shape and on the clip bench.
mkdir -p gpurun_out
T="timeout -s KILL"
python -c "import ctypes; L=ctypes.CDLL('insv2v_b200/libivv_b200.so'); import torch; torch.cuda.init(); torch.zeros(1,device='cuda'); print('cl4 clusters', L.ivv_debug_cl4_clusters())" 2>&1 | tail -1
$T 400 python -m pytest tests/test_kernels_gpu.py -q -x -k "geglu or linear or layernorm_folded" > gpurun_out/t_default.log 2>&1; echo "default exit=$?"; tail -2 gpurun_out/t_default.log
IVV_CL4=1 IVV_CL4_MIN=0 $T 400 python -m pytest tests/test_kernels_gpu.py -q -x -s -k "linear or layernorm_folded" > gpurun_out/t_cl4.log 2>&1; echo "cl4 exit=$?"; tail -2 gpurun_out/t_cl4.log
IVV_GEGLU_DS=2 $T 300 python -m pytest tests/test_kernels_gpu.py -q -x -s -k "geglu" > gpurun_out/t_geglu_ds.log 2>&1; echo "geglu_ds exit=$?"; tail -2 gpurun_out/t_geglu_ds.log
$T 900 python tools/linear_ab.py IVV_LIB_PATH=$PWD/insv2v_b200/libivv_b200_base.so "" IVV_CL4=1 IVV_GEGLU_DS=1 IVV_GEGLU_DS=2 > gpurun_out/linear_ab.txt 2>&1; cat gpurun_out/linear_ab.txt | grep -v Warning
for st in "IVV_LIB_PATH=$PWD/insv2v_b200/libivv_b200_base.so" "IVV_X=0" "IVV_CL4=1" "IVV_GEGLU_DS=1" "IVV_CL4=1 IVV_GEGLU_DS=1"; do
  tag=$(echo "$st" | tr ' /=' '___' | tail -c 40)
  env $st $T 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python -c "
import json,sys
try:
    j=json.load(open('gpurun_out/bench_$tag.json')); print('$st'[-40:], j['value'], j['ms_per_step'], j['gpu_launches'], j['clocks'])
except Exception as e: print('$st', 'FAILED', e)"
done
