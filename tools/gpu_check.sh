#!/bin/bash
# Runs the GPU kernel parity tests group by group, each under its own timeout so that one hung kernel
# cannot take the whole gpurun call with it. Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
for grp in "linear" "conv3x3" "groupnorm or layernorm or softmax_rows" "self_attention" "cross_attention" "temporal_attention" "layout or timestep or warp or flow or cfg"; do
  name=$(echo "$grp" | tr ' ' '_')
  echo "=== group: $grp"
  timeout -s KILL ${GROUP_TIMEOUT:-240} python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -k "$grp" > "gpurun_out/check_${name}.log" 2>&1
  echo "exit=$?"
  grep -E "^\[|passed|failed|error|Error" "gpurun_out/check_${name}.log" | tail -40
done
