#!/bin/bash
# ncu --set full + source page of the teacher's 56x56 c64 -> k256 1x1 convolution (HBM-bound, plain epilogue)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_fprop_kernel<.int.64, .bool.0, .int.1" --launch-skip 32 -c 1 -o gpurun_out/full_fprop_1x1_56 -f python tools/op_breakdown.py 256 > gpurun_out/ncu_full_1x1_56.log 2>&1; echo "ncu exit=$?"
ncu -i gpurun_out/full_fprop_1x1_56.ncu-rep --page raw --csv > gpurun_out/full_fprop_1x1_56.csv 2>/dev/null
ncu -i gpurun_out/full_fprop_1x1_56.ncu-rep --page source --csv > gpurun_out/full_fprop_1x1_56_src.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/full_fprop_1x1_56_src.csv 30 > gpurun_out/full_fprop_1x1_56_hot.txt 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/full_fprop_1x1_56.csv")))
H, U, V = rows[0], rows[1], rows[2]
for i, h in enumerate(H):
    if h in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__cluster_dim_x") or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        print("   %-90s %s %s" % (h, V[i], U[i]))
PY
head -34 gpurun_out/full_fprop_1x1_56_hot.txt
