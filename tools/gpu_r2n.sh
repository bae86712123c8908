#!/bin/bash
# concat elimination bring-up (+ everything else): all GPU tests, A/B bench, then the evidence pass
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest.log
for s in 0 1; do
  IVV_NO_CONCAT=$s timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-family > gpurun_out/bench_nocat_$s.json 2> gpurun_out/bench_nocat_$s.err
  python -c "
import json
j=json.load(open('gpurun_out/bench_nocat_$s.json')); print('NO_CONCAT=$s', j['value'], j['ms_per_step'], j['gpu_launches'], j['clocks'])"
done
bash tools/gpu_r2m.sh
