#!/bin/bash
# Attention kernel check: parity tests (each group under its own timeout), then the variant timing table.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_info.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -k "attention or variants" > gpurun_out/attn_tests.log 2>&1
echo "attention tests exit=$?"
grep -E "^\[|passed|failed|rror" gpurun_out/attn_tests.log | tail -40
timeout -s KILL 700 python tools/attn_bench.py > gpurun_out/attn_bench.txt 2>&1
echo "attn bench exit=$?"
cat gpurun_out/attn_bench.txt
