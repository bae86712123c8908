#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
timeout -s KILL 600 python tools/graph_timeline.py gpurun_out/graph_timeline.json > gpurun_out/graph_timeline.txt 2>&1; echo "timeline exit=$?"; head -80 gpurun_out/graph_timeline.txt
timeout -s KILL 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s > gpurun_out/pytest_fullsize.log 2>&1; echo "pytest exit=$?"
grep -E "^\[|^\.\[|^F\[|passed|failed|rror" gpurun_out/pytest_fullsize.log | tail -40
