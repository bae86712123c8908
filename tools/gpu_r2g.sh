#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for v in "IVV_X=1" "IVV_NO_WS=1"; do
  env $v timeout -s KILL 600 ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled -k regex:'ivv::' \
     -f -o gpurun_out/linears_$v python tools/ncu_linears.py > gpurun_out/ncu_linears_$v.log 2>&1
  echo "$v exit=$?"; tail -2 gpurun_out/ncu_linears_$v.log
done
ls -la gpurun_out/*.ncu-rep
