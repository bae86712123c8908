#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over tools/sanitize_kernels.py incl. the GEMM variants of the third session.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 120 python tools/sanitize_kernels.py > gpurun_out/sanitize_plain.txt 2>&1; echo "plain exit=$?"; tail -2 gpurun_out/sanitize_plain.txt | head -1
$T 240 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_kernels.py > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|Invalid|Misaligned|ok:" gpurun_out/sanitize_memcheck.txt | head -10
$T 240 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 60 python tools/sanitize_kernels.py > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck exit=$?"; grep -E "RACECHECK SUMMARY|ok:" gpurun_out/sanitize_racecheck.txt | head -5; grep "Race reported" gpurun_out/sanitize_racecheck.txt | sed 's/+0x.*//' | sort | uniq -c | head -30
grep "and Read access" gpurun_out/sanitize_racecheck.txt | sed 's/+0x.*//' | sort | uniq -c | head
