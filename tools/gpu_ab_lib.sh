#!/bin/bash
# A/B of two builds of the library on the same box: GEMM tile sweep (auto tiles) and attention, alternating.
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in libivv_b200.so libivv_b200_lane0.so; do
    echo "=== $lib (rep $rep)"
    IVV_LIB_PATH=$PWD/insv2v_b200/$lib timeout -s KILL 200 python tools/tile_sweep.py child 2>&1 | sed "s/^/[$lib] /"
  done
done | tee gpurun_out/ab_lib.txt
