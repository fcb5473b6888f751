#!/bin/bash
# Round-2 first visit: experimental options on hardware (parity + A/B timings), default gpu tests, bench.
mkdir -p gpurun_out
XEMO_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu_experimental -q -p no:cacheprovider > gpurun_out/pytest_exp.log 2>&1
echo "exp exit=$?"; tail -15 gpurun_out/pytest_exp.log
timeout 600 python tools/ab_options.py 256 > gpurun_out/ab_options.json 2> gpurun_out/ab_options.err; echo "ab exit=$?"
for v in 1 2; do XEMO_CONV_COSTMODEL=$v timeout 300 python tools/ab_options.py 256 costmodel > gpurun_out/ab_costmodel$v.json 2>> gpurun_out/ab_options.err; cat gpurun_out/ab_costmodel$v.json; done
cat gpurun_out/ab_options.json; tail -3 gpurun_out/ab_options.err
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
cut -c1-400 gpurun_out/bench_n1.json
