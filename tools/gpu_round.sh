#!/bin/bash
# One measurement round: all GPU tests (grouped timeouts), per-op profile of one forward, attention variants, bench.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
if [ -z "$SKIP_TESTS" ]; then
  timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest.log
fi
timeout -s KILL 300 python tools/profile_ops.py unet > gpurun_out/per_op_unet.txt 2>&1; echo "profile exit=$?"; head -45 gpurun_out/per_op_unet.txt
ATTN_BENCH_ONE=1 timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/attn_bench.txt 2>&1; cat gpurun_out/attn_bench.txt
if [ -n "$DO_BENCH" ]; then
  timeout -s KILL 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit=$?"; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
