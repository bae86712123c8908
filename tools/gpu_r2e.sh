#!/bin/bash
# all GPU tests after the dispatch clean-up + F<=64, sanitizer pass, long-form bench (configs[4] chained form)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/pytest.log
bash tools/gpu_sanitize.sh
timeout -s KILL 900 python bench.py --config long --steps 2 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err
echo "bench long exit=$?"; tail -3 gpurun_out/bench_long.err; cat gpurun_out/bench_long.json | cut -c1-700
