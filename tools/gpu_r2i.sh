#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -s -k "layernorm_folded" > gpurun_out/pytest_fold.log 2>&1; echo "pytest fold exit=$?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/pytest_fold.log | tail -8
for s in 0 1; do
  IVV_LN_FOLD=$s timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_fold_$s.json 2> gpurun_out/bench_fold_$s.err
  python -c "
import json
j=json.load(open('gpurun_out/bench_fold_$s.json')); print('LN_FOLD=$s', j['value'], j['ms_per_step'], j['gpu_launches'], j['clocks'])"
done
timeout -s KILL 300 python tools/graph_timeline.py gpurun_out/graph_timeline_fold.json > gpurun_out/graph_timeline_fold.txt 2>&1; head -24 gpurun_out/graph_timeline_fold.txt
