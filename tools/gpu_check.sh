#!/bin/bash
# One GPU-box visit: parity tests, per-op timing of an eager step, stem diagnostics, default bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown_new.txt 2>&1; echo "op_new exit=$?"
tail -32 gpurun_out/op_breakdown_new.txt
timeout 300 python tests/tools/stem_diag.py 16 300 > gpurun_out/stem_diag.txt 2>&1; timeout 300 python tests/tools/stem_diag.py 8 100 >> gpurun_out/stem_diag.txt 2>&1
cat gpurun_out/stem_diag.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
cat gpurun_out/bench_n1.json
