#!/bin/bash
# per-op timings after the cluster gate (256 and 32 pairs), SE tests once more with the squeeze rule
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_programs.py tests/test_gpu_net.py -x -q -m gpu -k "teacher or senet or se_" > gpurun_out/pytest_se.log 2>&1
echo "se tests exit=$?"; tail -3 gpurun_out/pytest_se.log
timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown_256.txt 2>&1; echo "op256 exit=$?"
timeout 300 python tools/op_breakdown.py 32 > gpurun_out/op_breakdown_32.txt 2>&1; echo "op32 exit=$?"
grep "op_se_gate\|op_se_squeeze" gpurun_out/op_breakdown_256.txt | head -40
echo ---- 32
awk '/---- total/{f=1} f' gpurun_out/op_breakdown_32.txt | head -30
