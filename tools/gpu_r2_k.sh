#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_callers.py tests/test_gpu_programs.py -q -p no:cacheprovider -k "deterministic or resumes or teacher" > gpurun_out/pytest_det.log 2>&1
echo "det exit=$?"; grep -E "passed|failed|^FAILED|^E  .*(Assert|assert)" gpurun_out/pytest_det.log | head -12
for v in 0 1; do XEMO_SE_GATE_SPLIT=$v timeout 200 python tools/ab_options.py 256 se_gate 2>/dev/null; XEMO_SE_GATE_SPLIT=$v timeout 200 python tools/ab_options.py 32 se_gate 2>/dev/null; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_b32.csv python bench.py --steps 2 --warmup 3 --scaling weak --per-gpu-batch 32 --overlap 0 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_b32.log 2>&1; echo "ncu b32 exit=$?"
python tools/launch_summary.py gpurun_out/launches_b32.csv 194 | head -30
timeout 300 python bench.py --steps 20 --warmup 5 --scaling weak --per-gpu-batch 32 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_b32.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/bench_b32.json').read().strip().splitlines()[-1]);print('B=32', d['value'], d['ms_per_step'], d['kernels_per_step'])"
