#!/bin/bash
# consistency of the new GEMM variants at the 48x72 (configs[4]) and 32x48 shapes, then the configs[4] bench line
mkdir -p gpurun_out
timeout -s KILL 300 python tools/variant_consistency.py > gpurun_out/variant_consistency.txt 2>&1; echo "consistency exit=$?"; grep -v Warn gpurun_out/variant_consistency.txt | tail -4
timeout -s KILL 420 python bench.py --config long --steps 2 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err; echo "bench long exit=$?"; cut -c1-330 gpurun_out/bench_long.json
