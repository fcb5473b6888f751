#!/bin/bash
# 2-GPU visit: C ABI tests (incl. the data-parallel check), strong + weak scaling bench at N = 2, 1-GPU bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_net.py -q -p no:cacheprovider > gpurun_out/pytest_net.log 2>&1
echo "net exit=$?"; tail -30 gpurun_out/pytest_net.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 exit=$?"
cut -c1-600 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 exit=$?"
cut -c1-600 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json", "gpurun_out/bench_n2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["scaling"], d.get("weak_scaling"), d["e2e"]["value"], d.get("parity_mode"), d["roofline"] and {k: d["roofline"][k] for k in ("frac", "step_frac", "teacher_forward", "student_step")})
    except Exception as e:
        print(f, "unreadable", e)
PY
