#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --watchdog 240 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench8 exit=$?"
grep -v "^$" gpurun_out/bench_n8.err | grep -v "OMP_NUM\|\*\*\*\*" | tail -15
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
    print("N=8", d["value"], d["ms_per_step"], d["scaling"], d.get("weak_scaling"), d["e2e"]["value"], d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
