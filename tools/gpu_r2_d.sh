#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_isolation.py tests/test_gpu_coupling.py -q -s --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_parity.log 2>&1
echo "parity exit=$?"; grep -E "^(FAILED|ERROR)|passed|failed|AssertionError" gpurun_out/pytest_parity.log | tail -30
grep -A45 "fp32-equivalent student step (hot-cross-ent), N = 16" gpurun_out/pytest_parity.log | head -50
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_isolation.py --deselect tests/test_gpu_coupling.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?"; tail -8 gpurun_out/pytest_gpu.log
