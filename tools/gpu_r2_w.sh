#!/bin/bash
# 2 GPUs with the final code: the data-parallel test, then the bench (strong + weak in one line) and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_net.py -q -p no:cacheprovider -k "two_gpus" > gpurun_out/pytest_net2.log 2>&1
echo "dp test exit=$?"; tail -3 gpurun_out/pytest_net2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --watchdog 300 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 exit=$?"
grep -v "^$" gpurun_out/bench_n2.err | tail -5
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
    print("N=2", d["value"], d["ms_per_step"], d["scaling"], d.get("weak_scaling"), d["e2e"]["value"])
except Exception as e:
    print("unreadable", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref2 exit=$?"; cut -c1-300 gpurun_out/bench_ref_n2.json
