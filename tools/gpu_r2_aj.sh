#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_programs.py tests/test_gpu_callers.py -x -q -m gpu > gpurun_out/pytest_aj.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/pytest_aj.log
timeout 300 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c5.json").read().strip().splitlines()[-1])
print("c5", round(d["value"], 1), {k: (round(v['teacher_ms'], 3), round(v['student_ms'], 3)) for k, v in d['sweep'].items()})
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("n1", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), r["frac"], r["step_frac"], r["teacher_forward"], r["student_step"], d["parity_mode"]["ms_per_step"])
d = json.loads(open("gpurun_out/bench_c3.json").read().strip().splitlines()[-1])
print("c3", round(d["value"], 1), round(d["ms_per_step"], 3), d["e2e"]["value"], d["roofline"]["frac"])
PY
