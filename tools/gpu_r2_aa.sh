#!/bin/bash
# occupancy-sized grids (XEMO_GRID_WAVES) and the filter-gradient item order: tests, per-op times, bench, ncu of conv2's wgrad
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vl_ops.py tests/test_gpu_programs.py tests/test_gpu_net.py -x -q -m gpu > gpurun_out/pytest_aa.log 2>&1
echo "tests exit=$?"; tail -4 gpurun_out/pytest_aa.log
for w in 1 2 4; do
  XEMO_GRID_WAVES=$w timeout 300 python tools/op_breakdown.py 256 > gpurun_out/op_breakdown_w$w.txt 2>&1
  echo "== waves $w"; grep -- "---- total" gpurun_out/op_breakdown_w$w.txt
  awk '/---- total/{f=1} f' gpurun_out/op_breakdown_w$w.txt | grep "maxpool\|bn_bwd\|stem_pool\|affine\|excite\|im2col\|s2d\|wgrad"
  XEMO_GRID_WAVES=$w timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-mode --watchdog 250 > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_w$w.json").read().strip().splitlines()[-1])
print("waves $w:", round(d["value"], 1), round(d["ms_per_step"], 3), d["roofline"]["teacher_forward"]["ms"], d["roofline"]["student_step"]["ms"])
PY
done
grep "op_conv_wgrad" gpurun_out/op_breakdown_w1.txt | head -9
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel --launch-skip 14 -c 1 -o gpurun_out/full_wgrad_conv2_order -f python tools/op_breakdown.py 256 > gpurun_out/ncu_full_wgrad_order.log 2>&1; echo "ncu exit=$?"
ncu -i gpurun_out/full_wgrad_conv2_order.ncu-rep --page raw --csv > gpurun_out/full_wgrad_conv2_order.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/full_wgrad_conv2_order.csv")))
H, U, V = rows[0], rows[1], rows[2]
for i, h in enumerate(H):
    if h in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_lookup_hit.sum", "lts__t_sectors_lookup_miss.sum"):
        print("   %-80s %s %s" % (h, V[i], U[i]))
PY
