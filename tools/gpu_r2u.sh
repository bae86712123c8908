#!/bin/bash
# persistent one-tile attention as default: remaining tests; compute-sanitizer over the extended kernel list; GN CTA-count variants.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention or groupnorm or geglu" > gpurun_out/t_u.log 2>&1; echo "attention/gn/geglu exit=$?"; tail -3 gpurun_out/t_u.log
for v in "" gncps2 gncps8; do
  lib=$PWD/insv2v_b200/libivv_b200${v:+_$v}.so
  IVV_LIB_PATH=$lib $T 300 python tools/graph_timeline.py gpurun_out/tl3_$v.json > gpurun_out/tl3_$v.txt 2>&1; echo "variant '$v'"; sed -n 4p gpurun_out/tl3_$v.txt
  python - <<PY
import re
t=0; big=0
for l in open('gpurun_out/tl3_$v.txt'):
    if 'groupnorm' in l:
        m=re.search(r"\)\s+(\d+)\s+([\d.]+)\s+([\d.]+)",l); t+=float(m.group(2))
        if '73728' in l: big+=float(m.group(2))
print('  groupnorm total ms', round(t,3), ' level-0 (two-kernel) ms', round(big,3))
PY
done
$T 120 python tools/sanitize_kernels.py > gpurun_out/sanitize_plain.txt 2>&1; echo "plain exit=$?"; tail -2 gpurun_out/sanitize_plain.txt | head -1
$T 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_kernels.py > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|Invalid|Misaligned|ok:" gpurun_out/sanitize_memcheck.txt | head -10
$T 600 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 40 python tools/sanitize_kernels.py > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck exit=$?"; grep -E "RACECHECK SUMMARY|ok:" gpurun_out/sanitize_racecheck.txt | head -5; grep "Race reported" gpurun_out/sanitize_racecheck.txt | sed 's/+0x.*//' | sort | uniq -c | head -30
