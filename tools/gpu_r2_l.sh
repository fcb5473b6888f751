#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_callers.py -q -p no:cacheprovider > gpurun_out/pytest_l.log 2>&1
echo "tests exit=$?"; grep -E "passed|failed|^FAILED|^E  .*(Assert|assert)" gpurun_out/pytest_l.log | head -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --watchdog 300 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "bench2 exit=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2b.json").read().strip().splitlines()[-1])
    print("N=2", d["value"], d["ms_per_step"], d["scaling"], d.get("weak_scaling"), d["e2e"]["value"])
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/bench_n2b.err").read()[-3000:])
PY
timeout 300 python bench.py --steps 20 --warmup 5 --scaling weak --per-gpu-batch 32 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_b32b.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/bench_b32b.json').read().strip().splitlines()[-1]);print('B=32', d['value'], d['ms_per_step'], d['kernels_per_step'])"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_n1b.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/bench_n1b.json').read().strip().splitlines()[-1]);print('B=256', d['value'], d['ms_per_step'], d['kernels_per_step'], d['e2e']['value'])"
