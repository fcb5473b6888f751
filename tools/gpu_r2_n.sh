#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vl_ops.py tests/test_gpu_programs.py tests/test_gpu_net.py -q -p no:cacheprovider > gpurun_out/pytest_n.log 2>&1
echo "tests exit=$?"; grep -E "passed|failed|^FAILED|^E  .*(Assert|assert)" gpurun_out/pytest_n.log | head -12
for m in 0 1; do
XEMO_CONV_2CTA=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_2cta$m.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/bench_2cta$m.json').read().strip().splitlines()[-1]);r=d['roofline'];print('XEMO_CONV_2CTA=$m', d['value'], d['ms_per_step'], 'conv frac', r['frac'], 'step', r['step_frac'], r['teacher_forward'], r['student_step'])"
done
timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown_2cta.txt 2>&1; grep -E "calls" gpurun_out/op_breakdown_2cta.txt | head -8; grep -E "5x5|3x3" gpurun_out/op_breakdown_2cta.txt | head -30
