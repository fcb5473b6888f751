#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_fullsize.py -q -p no:cacheprovider > gpurun_out/pytest_net.log 2>&1
echo "net+fullsize exit=$?"; grep -E "passed|failed|^FAILED|^E  .*(Assert|assert)" gpurun_out/pytest_net.log | head -30
timeout 600 python -m pytest tests/test_gpu_programs.py -q -p no:cacheprovider -k "se_blocks or teacher" > gpurun_out/pytest_se.log 2>&1
echo "se exit=$?"; tail -5 gpurun_out/pytest_se.log
timeout 400 python tools/ab_options.py 256 se_lin > gpurun_out/ab_se_hybrid.json 2> gpurun_out/ab_se_hybrid.err; cat gpurun_out/ab_se_hybrid.json; tail -3 gpurun_out/ab_se_hybrid.err
for ov in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --scaling weak --per-gpu-batch 32 --overlap $ov --no-cpu-baseline --no-parity-mode --watchdog 250 > gpurun_out/bench_b32_ov$ov.json 2> gpurun_out/bench_b32_ov$ov.err; echo "b32 ov$ov exit=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_b32_ov$ov.json').read().strip().splitlines()[-1]);print('B=32 overlap $ov', d['value'], d['ms_per_step'], d['kernels_per_step'])" || tail -5 gpurun_out/bench_b32_ov$ov.err
done
for ov in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --overlap $ov --no-cpu-baseline --no-parity-mode --watchdog 250 > gpurun_out/bench_b256_ov$ov.json 2> gpurun_out/bench_b256_ov$ov.err; echo "b256 ov$ov exit=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_b256_ov$ov.json').read().strip().splitlines()[-1]);print('B=256 overlap $ov', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['teacher_forward'], d['roofline']['student_step'])" || tail -5 gpurun_out/bench_b256_ov$ov.err
done
