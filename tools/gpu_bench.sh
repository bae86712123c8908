#!/bin/bash
# bench + ncu launch list (separate runs: a number printed under ncu is never a bench value)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 1500 python bench.py --steps ${STEPS:-2} --warmup ${WARMUP:-3} ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ -n "$NCU_LIST" ]; then
  timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
     python tools/one_forward.py > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit=$?"; tail -2 gpurun_out/ncu_list.log
  python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; head -40 gpurun_out/launch_summary.txt
fi
