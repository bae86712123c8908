#!/bin/bash
# Quick check of a kernel change: kernel parity tests, per-shape timings, clip bench (no CPU baseline leg).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels exit=$?"; tail -5 gpurun_out/pytest_kernels.log
timeout -s KILL 300 python tools/gemm_ab.py ${AB_SETTINGS:-IVV_X=1} > gpurun_out/gemm_ab_quick.txt 2>&1; cat gpurun_out/gemm_ab_quick.txt
timeout -s KILL 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['clocks'])"
