#!/bin/bash
# End-of-round measurement set: all GPU tests, smoke, the bench lines of every named config, reference arm, in-graph timeline.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt; nproc >> gpurun_out/gpu_info.txt
timeout -s KILL 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout -s KILL 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"; cut -c1-300 gpurun_out/bench_ref.json
for cfg in flow c1; do
  timeout -s KILL 600 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "bench $cfg exit=$?"; cut -c1-300 gpurun_out/bench_$cfg.json
done
timeout -s KILL 300 python tools/graph_timeline.py gpurun_out/graph_timeline_final.json > gpurun_out/graph_timeline_final.txt 2>&1; head -12 gpurun_out/graph_timeline_final.txt
timeout -s KILL 900 python bench.py --config long --steps 2 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err; echo "bench long exit=$?"; cut -c1-300 gpurun_out/bench_long.json
