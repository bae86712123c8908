#!/bin/bash
# End-of-session measurement set, ordered by importance (the GPU budget may cut the tail): GPU tests, smoke, default
# bench (with CPU baseline), ncu full capture of the hot kernels, reference arm, per-op profile, ncu launch list.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|attention|temporal_attn|gn_|layernorm' \
   -s 16 -c 16 -f -o gpurun_out/hot_kernels python tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"; tail -2 gpurun_out/ncu_full.log
timeout -s KILL 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"; cat gpurun_out/bench_ref.json
timeout -s KILL 200 python tools/profile_ops.py unet > gpurun_out/per_op_unet.txt 2>&1; head -12 gpurun_out/per_op_unet.txt
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python tools/one_forward.py > gpurun_out/ncu_list.log 2>&1
echo "launch list exit=$?"
timeout -s KILL 200 python tools/profile_ops.py vae > gpurun_out/per_op_vae.txt 2>&1; head -5 gpurun_out/per_op_vae.txt
