#!/bin/bash
# End-of-round measurement set: all GPU tests, smoke, default bench (with CPU baseline), reference arm, per-op profiles,
# attention variants, ncu launch list + full capture. Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json
timeout -s KILL 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"; cat gpurun_out/bench_ref.json
timeout -s KILL 300 python tools/profile_ops.py unet > gpurun_out/per_op_unet.txt 2>&1
timeout -s KILL 300 python tools/profile_ops.py vae > gpurun_out/per_op_vae.txt 2>&1
timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/attn_bench.txt 2>&1; cat gpurun_out/attn_bench.txt
if [ -n "$DO_NCU" ]; then bash tools/gpu_ncu.sh; fi
