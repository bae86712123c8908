#!/bin/bash
# Round-2 first call: all GPU tests (incl. full-size parity + calibration), then the bench lines of the named configs.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt; nproc >> gpurun_out/gpu_info.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q -s --durations=15 ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"
grep -E "^\[|passed|failed|error|Error|PyTorch fp16" gpurun_out/pytest.log | tail -60
for cfg in c2 flow c1; do
  timeout -s KILL 600 python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "bench $cfg exit=$?"; tail -3 gpurun_out/bench_$cfg.err; python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_$cfg.json"))
    print({k: j[k] for k in ("value", "ms_per_step", "gpu_launches")}, j["e2e"]["value"], j["step_tensor_roofline"]["frac"], j["roofline"]["frac"], j["roofline"].get("family", {}).get("gemm_family"))
except Exception as e:
    print("no json", e)
PY
done
