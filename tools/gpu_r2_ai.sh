#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 1 2; do
XEMO_GRID_WAVES=$w timeout 300 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/bench_c5_w$w.json 2> gpurun_out/bench_c5_w$w.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c5_w$w.json").read().strip().splitlines()[-1])
print("waves $w", {k: (round(v['teacher_ms'], 3), round(v['student_ms'], 3)) for k, v in d['sweep'].items()})
PY
done
