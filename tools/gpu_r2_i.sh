#!/bin/bash
# Full validation + measurements on one GPU.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -20; grep -A16 "slowest" gpurun_out/pytest_gpu.log | head -18
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
for c in c2 c3 c5; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "$c exit=$?"; tail -2 gpurun_out/bench_$c.err; done
python - <<'PY'
import json
for f in ("bench_n1", "bench_ref", "bench_c2", "bench_c3", "bench_c5"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["value"], 1), d["unit"], round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "frac", r.get("frac"), r.get("step_frac"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_bench.log 2>&1; echo "launch list exit=$?"
timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown.txt 2>&1; echo "op exit=$?"; tail -35 gpurun_out/op_breakdown.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_ --csv --log-file gpurun_out/conv_traffic.csv python tools/op_breakdown.py 256 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic exit=$?"
python tools/conv_traffic.py gpurun_out/conv_traffic.csv 0 256 gpurun_out/conv_traffic.json
