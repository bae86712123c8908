#!/bin/bash
# pair160 (v3 epilogue) bring-up: parity tests of every GEMM path, A/B against the v2 kernels, in-graph timeline, bench.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "linear or conv or gemm" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit=$?"; tail -5 gpurun_out/pytest_gemm.log
GEMM_AB_ONLY=linear timeout -s KILL 300 python tools/gemm_ab.py IVV_EPI2=0 IVV_EPI2=1 > gpurun_out/gemm_ab_epi2.txt 2>&1; cat gpurun_out/gemm_ab_epi2.txt
timeout -s KILL 600 python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/pytest_models.log 2>&1; echo "pytest models exit=$?"; tail -5 gpurun_out/pytest_models.log
for s in 0 1; do
  IVV_EPI2=$s timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_epi2_$s.json 2> gpurun_out/bench_epi2_$s.err
  python -c "
import json
j=json.load(open('gpurun_out/bench_epi2_$s.json')); print('EPI2=$s', j['value'], j['ms_per_step'], j['clocks'])"
done
timeout -s KILL 300 python tools/graph_timeline.py gpurun_out/graph_timeline_epi2.json > gpurun_out/graph_timeline_epi2.txt 2>&1; head -40 gpurun_out/graph_timeline_epi2.txt
