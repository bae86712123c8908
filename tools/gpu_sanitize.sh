#!/bin/bash
# compute-sanitizer over one launch of every hot-path kernel family (tools/sanitize_kernels.py): memcheck (out-of-bounds /
# misaligned accesses, incl. the TMA / mbarrier kernels) and racecheck (shared-memory hazards).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 120 python tools/sanitize_kernels.py > gpurun_out/sanitize_plain.txt 2>&1; echo "plain exit=$?"; tail -3 gpurun_out/sanitize_plain.txt
timeout -s KILL 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_kernels.py > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|Invalid|Misaligned|ok:" gpurun_out/sanitize_memcheck.txt | head -20
timeout -s KILL 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tools/sanitize_kernels.py > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck exit=$?"; grep -E "RACECHECK SUMMARY|hazard|ok:" gpurun_out/sanitize_racecheck.txt | head -20
