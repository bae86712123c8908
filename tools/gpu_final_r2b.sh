#!/bin/bash
# End-of-round set (third session): all GPU tests, smoke, bench lines (configs[1] default, reference arm, flow, c1), in-graph
# timeline, ncu --set full of the GEMM variants added this session.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt; nproc >> gpurun_out/gpu_info.txt
timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout -s KILL 300 python tools/graph_timeline.py gpurun_out/graph_timeline_final3.json > gpurun_out/graph_timeline_final3.txt 2>&1; sed -n 4,12p gpurun_out/graph_timeline_final3.txt
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc' -f -o gpurun_out/new_gemms python tools/ncu_new_gemms.py > gpurun_out/ncu_new.log 2>&1; echo "ncu exit=$?"; tail -2 gpurun_out/ncu_new.log
python tools/ncu_extract.py gpurun_out/new_gemms.ncu-rep --by-grid > gpurun_out/new_gemms_ncu.csv 2>/dev/null; wc -l gpurun_out/new_gemms_ncu.csv
timeout -s KILL 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"; cut -c1-300 gpurun_out/bench_ref.json
for cfg in flow c1; do
  timeout -s KILL 600 python bench.py --config $cfg --no-cpu-baseline --no-family > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "bench $cfg exit=$?"; cut -c1-300 gpurun_out/bench_$cfg.json
done
