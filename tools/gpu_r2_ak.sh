#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 --watchdog 250 --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench4 exit=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n4.json").read().strip().splitlines()[-1])
print("N=4", d["value"], d["ms_per_step"], d["scaling"], d.get("weak_scaling"), d["e2e"]["value"])
PY
