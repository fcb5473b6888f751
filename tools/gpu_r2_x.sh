#!/bin/bash
# 8 GPUs with the final code: bench (strong + weak in one line), N=4 as well
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --watchdog 300 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench$n exit=$?"
grep -v "^$" gpurun_out/bench_n$n.err | grep -v "OMP_NUM\|\*\*\*\*" | tail -5
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", d["value"], d["ms_per_step"], d["scaling"], d.get("weak_scaling"), d["e2e"]["value"])
except Exception as e:
    print("unreadable", e)
PY
done
