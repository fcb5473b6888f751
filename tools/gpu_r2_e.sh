#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_net.py -q -x -p no:cacheprovider > gpurun_out/pytest_net.log 2>&1
echo "net exit=$?"; tail -40 gpurun_out/pytest_net.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_isolation.py tests/test_gpu_coupling.py -q -p no:cacheprovider > gpurun_out/pytest_parity2.log 2>&1
echo "parity exit=$?"; tail -5 gpurun_out/pytest_parity2.log
