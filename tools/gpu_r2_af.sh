#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -x -q -m gpu -k "se_ or teacher or c5" > gpurun_out/pytest_af.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/pytest_af.log
timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown_af.txt 2>&1
awk '/---- total/{f=1} f' gpurun_out/op_breakdown_af.txt | grep "total\|squeeze"
for b in 256 32; do timeout 300 python tools/ab_options.py $b gate 2>&1 | grep teacher | sed "s/^/B=$b /"; done
