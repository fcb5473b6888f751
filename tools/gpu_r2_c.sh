#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/tools/parity_diag.py 4 100 > gpurun_out/parity_diag_n4.txt 2>&1; echo "diag exit=$?"
cat gpurun_out/parity_diag_n4.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_new.json 2> gpurun_out/bench_n1_new.err; echo "bench exit=$?"
cut -c1-1500 gpurun_out/bench_n1_new.json; tail -5 gpurun_out/bench_n1_new.err
timeout 300 python bench.py --steps 10 --warmup 3 --scaling weak --per-gpu-batch 32 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "b32 exit=$?"
cut -c1-400 gpurun_out/bench_b32.json; tail -3 gpurun_out/bench_b32.err
timeout 300 python tools/op_breakdown.py 32 > gpurun_out/op_breakdown_b32.txt 2>&1; tail -40 gpurun_out/op_breakdown_b32.txt
