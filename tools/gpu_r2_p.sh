#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py tests/test_gpu_programs.py -q -p no:cacheprovider -k "full_height or training_step or distillation_step" > gpurun_out/pytest_p.log 2>&1
echo "tests exit=$?"; grep -E "passed|failed|^FAILED|^E  .*(Assert|assert)" gpurun_out/pytest_p.log | head -12
for m in 0 1; do
XEMO_DGRAD_FULLHEIGHT=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_fh$m.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/bench_fh$m.json').read().strip().splitlines()[-1]);r=d['roofline'];print('XEMO_DGRAD_FULLHEIGHT=$m', d['value'], d['ms_per_step'], 'conv frac', r['frac'], 'step', r['step_frac'], r['student_step'])"
done
