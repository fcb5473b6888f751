#!/bin/bash
# Round-2 visit B: parity mode, isolation, coupling, guard tests; SE_LIN op breakdown.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_isolation.py tests/test_gpu_coupling.py -q -s --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_parity.log 2>&1
echo "parity exit=$?"; grep -v "^$" gpurun_out/pytest_parity.log | tail -120
timeout 600 python -m pytest tests/test_gpu_programs.py -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_programs.log 2>&1
echo "programs exit=$?"; tail -30 gpurun_out/pytest_programs.log
XEMO_SE_LIN=1 timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown_selin.txt 2>&1; echo "op exit=$?"
grep -E "se_|total" gpurun_out/op_breakdown_selin.txt | sort | uniq -c | sort -rn | head -5
