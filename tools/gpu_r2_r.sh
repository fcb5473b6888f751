#!/bin/bash
# SE gate as a cluster kernel: correctness, A/B at 256 and 32 faces, squeeze block-size sweep, parity-mode timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vl_ops.py tests/test_gpu_programs.py tests/test_gpu_net.py -x -q -m gpu -k "se_gate or scratch_pool or teacher or senet or SE or se_" > gpurun_out/pytest_gate.log 2>&1
echo "gate tests exit=$?"; tail -4 gpurun_out/pytest_gate.log
for b in 256 32; do
  for v in 0 1; do
    XEMO_SE_GATE_CLUSTER=$v timeout 300 python tools/ab_options.py $b gate 2>&1 | grep teacher
  done
done
timeout 300 python tools/ab_options.py 256 squeeze 2>&1 | grep teacher
timeout 300 python tools/ab_options.py 32 squeeze 2>&1 | grep teacher
timeout 600 python bench.py --steps 10 --warmup 3 --watchdog 500 > gpurun_out/bench_gate.json 2> gpurun_out/bench_gate.err
echo "bench exit=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_gate.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['teacher_forward'], d['parity_mode']['ms_per_step'])
PY
timeout 300 python bench.py --steps 10 --warmup 3 --scaling weak --per-gpu-batch 32 --no-cpu-baseline --no-parity-mode --watchdog 250 > gpurun_out/bench_gate_b32.json 2> gpurun_out/bench_gate_b32.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_gate_b32.json').read().strip().splitlines()[-1])
print('B=32', d['value'], d['ms_per_step'])
PY
