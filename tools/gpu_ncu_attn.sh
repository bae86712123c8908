#!/bin/bash
# ncu --set full of one S=1536, d=40 attention launch per kernel variant
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
i=0
for v in "IVV_ATTN_PAIR=1" "IVV_ATTN_PAIR=0"; do
  env $v timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:'attention' -s 1 -c 1 -f \
     -o gpurun_out/attn_v$i python tools/ncu_one_attn.py > gpurun_out/ncu_attn_$i.log 2>&1
  echo "$v exit=$?"; tail -2 gpurun_out/ncu_attn_$i.log
  i=$((i+1))
done
ls -la gpurun_out/*.ncu-rep
