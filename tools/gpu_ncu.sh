#!/bin/bash
# ncu evidence: (1) launch list of one eager forward + decode, (2) --set full capture of the hot kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python tools/one_forward.py > gpurun_out/ncu_list.log 2>&1
echo "launch list exit=$?"
timeout -s KILL 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|attention|temporal_attn|gn_|layernorm' \
   -s 16 -c 16 -f -o gpurun_out/hot_kernels python tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
