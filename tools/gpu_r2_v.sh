#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py tests/test_gpu_programs.py tests/test_gpu_net.py -x -q -m gpu -k "se_ or teacher or senet or fullsize or c2 or c5" > gpurun_out/pytest_v.log 2>&1
echo "tests exit=$?"; tail -5 gpurun_out/pytest_v.log
for b in 256 32; do timeout 300 python tools/ab_options.py $b gate 2>&1 | grep teacher | sed "s/^/B=$b /"; done
