#!/bin/bash
# Double-staging residual GEMMs + cp.async temporal attention: kernel parity tests, per-shape A/B, clip bench A/B.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels exit=$?"; tail -15 gpurun_out/pytest_kernels.log
timeout -s KILL 300 python tools/gemm_ab.py IVV_DS=0 IVV_DS=1 > gpurun_out/gemm_ab_ds.txt 2>&1; grep -v "taps=9" gpurun_out/gemm_ab_ds.txt
for setting in "IVV_DS=0" "IVV_DS=1"; do
  tag=$(echo $setting | tr ' =' '__')
  echo "=== $setting"
  env $setting timeout -s KILL 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
  echo "exit=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['clocks'])"
done
