#!/bin/bash
# ncu evidence for every hot-path kernel: --set full capture inside the profiler window of tools/ncu_all_kernels.py.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 1500 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:'ivv::' \
   -f -o gpurun_out/all_kernels python tools/ncu_all_kernels.py > gpurun_out/ncu_all.log 2>&1
echo "full capture exit=$?"; tail -3 gpurun_out/ncu_all.log; ls -la gpurun_out/*.ncu-rep
