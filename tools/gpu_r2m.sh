#!/bin/bash
# Evidence pass: ncu --set full of every hot-path kernel (one launch each), RAFT kernels with the speed-of-light sections,
# launch lists (configs[1] forward + decode, one 384x576 forward). Reports are reduced to CSV ON THE BOX: the .ncu-rep
# files are too large to travel back.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
NCU_RAFT=0 timeout -s KILL 1200 ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled -k regex:'ivv::' \
   -f -o /tmp/all_kernels python tools/ncu_all_kernels.py > gpurun_out/ncu_all.log 2>&1
echo "full capture exit=$?"; tail -2 gpurun_out/ncu_all.log
python tools/ncu_extract.py /tmp/all_kernels.ncu-rep --by-grid > gpurun_out/all_kernels_ncu_full.csv
ncu -i /tmp/all_kernels.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/all_kernels_ncu_raw.csv.gz
NCU_RAFT=only timeout -s KILL 600 ncu --section SpeedOfLight --section LaunchStats --section MemoryWorkloadAnalysis --clock-control none --profile-from-start off \
   --kernel-name-base demangled -k regex:'ivv::' -f -o /tmp/raft_kernels python tools/ncu_all_kernels.py > gpurun_out/ncu_raft.log 2>&1
echo "raft capture exit=$?"; tail -1 gpurun_out/ncu_raft.log
ncu -i /tmp/raft_kernels.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/raft_kernels_ncu_raw.csv.gz
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv \
   python tools/one_forward.py > gpurun_out/ncu_list_c2.log 2>&1; echo "launch list c2 exit=$?"
SHAPE=16,48,72 DECODE=0 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_long.csv \
   python tools/one_forward.py > gpurun_out/ncu_list_long.log 2>&1; echo "launch list 384x576 exit=$?"
gzip -f gpurun_out/launches_c2.csv gpurun_out/launches_long.csv
ls -la gpurun_out/
