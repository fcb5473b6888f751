#!/bin/bash
# Round-end GPU visit: parity tests, op breakdown, both bench arms, ncu launch list + conv DRAM traffic + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown.txt 2>&1; echo "op exit=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "launch list exit=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_ --csv --log-file gpurun_out/conv_traffic.csv python tools/op_breakdown.py 256 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic exit=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_fprop --launch-skip 127 -c 1 -o gpurun_out/full_fprop_conv2 python tools/op_breakdown.py 256 > gpurun_out/ncu_full1.log 2>&1; echo "full1 exit=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad --launch-skip 14 -c 1 -o gpurun_out/full_wgrad_conv2 python tools/op_breakdown.py 256 > gpurun_out/ncu_full2.log 2>&1; echo "full2 exit=$?"
cat gpurun_out/bench_n1.json | cut -c1-300
cat gpurun_out/bench_ref.json | cut -c1-300
grep -E "total|op_conv_wgrad .*calls|op_conv_fwd .*calls|op_face" gpurun_out/op_breakdown.txt
