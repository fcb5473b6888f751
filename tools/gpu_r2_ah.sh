#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vl_ops.py tests/test_gpu_programs.py tests/test_gpu_fullsize.py tests/test_gpu_net.py -x -q -m gpu > gpurun_out/pytest_ah.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/pytest_ah.log
timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown_ah.txt 2>&1
awk '/---- total/{f=1} f' gpurun_out/op_breakdown_ah.txt | head -6
grep "op_conv_fwd " gpurun_out/op_breakdown_ah.txt | head -12
grep "257x74" gpurun_out/op_breakdown_ah.txt
for b in 256; do timeout 300 python tools/ab_options.py $b gate 2>&1 | grep teacher | sed "s/^/B=$b /"; done
