#!/bin/bash
# KB2 on by default: bit-identity of a full UNet3D forward against all variants off, kernel tests, smoke.
mkdir -p gpurun_out
timeout -s KILL 100 python tools/variant_consistency.py 16,32,48 > gpurun_out/variant_consistency2.txt 2>&1; echo "consistency exit=$?"; grep -v Warn gpurun_out/variant_consistency2.txt | tail -2
timeout -s KILL 60 python -m pytest tests/test_kernels_gpu.py -q -k "linear or layernorm_folded" > gpurun_out/t_ad.log 2>&1; echo "tests exit=$?"; tail -1 gpurun_out/t_ad.log
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke2.log
