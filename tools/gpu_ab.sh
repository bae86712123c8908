#!/bin/bash
# A/B of an env switch on the full bench: tests first, then bench with each setting (no CPU baseline leg).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/pytest.log
for setting in ${AB_SETTINGS}; do
  echo "=== $setting"
  env $setting timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${setting}.json 2> gpurun_out/bench_${setting}.err
  echo "exit=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${setting}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
