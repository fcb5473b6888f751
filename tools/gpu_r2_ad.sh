#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_programs.py tests/test_gpu_net.py tests/test_gpu_fullsize.py -x -q -m gpu -k "teacher or senet or se_ or c5 or fullsize" > gpurun_out/pytest_ad.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/pytest_ad.log
timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown_ad.txt 2>&1
grep "op_conv_fwd_nc" gpurun_out/op_breakdown_ad.txt
grep -- "---- total" gpurun_out/op_breakdown_ad.txt
for b in 256 32; do timeout 300 python tools/ab_options.py $b gate 2>&1 | grep teacher | sed "s/^/B=$b /"; done
