#!/bin/bash
# Quick check of a kernel change: kernel parity tests, per-op profile of one UNet forward, clip bench (no CPU baseline).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels exit=$?"; tail -5 gpurun_out/pytest_kernels.log
timeout -s KILL 300 python tools/profile_ops.py unet > gpurun_out/per_op_unet.txt 2>&1; head -60 gpurun_out/per_op_unet.txt
timeout -s KILL 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['clocks'])"
