#!/bin/bash
# ncu --set full of the teacher's HBM-bound expand convolution with the SE excite folded in (56x56, c64 -> k256, + shortcut)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_fprop_kernel<.int.64, .bool.1" --launch-skip 7 -c 1 -o gpurun_out/full_fprop_nc56 -f python tools/op_breakdown.py 256 > gpurun_out/ncu_full_nc56.log 2>&1; echo "ncu exit=$?"
tail -3 gpurun_out/ncu_full_nc56.log
ncu -i gpurun_out/full_fprop_nc56.ncu-rep --page raw --csv > gpurun_out/full_fprop_nc56.csv 2>/dev/null
ncu -i gpurun_out/full_fprop_nc56.ncu-rep --page source --csv > gpurun_out/full_fprop_nc56_src.csv 2>/dev/null
python tools/ncu_hot.py gpurun_out/full_fprop_nc56_src.csv 40 > gpurun_out/full_fprop_nc56_hot.txt 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/full_fprop_nc56.csv")))
H, U, V = rows[0], rows[1], rows[2]
for i, h in enumerate(H):
    if h in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "lts__t_sectors_lookup_hit.sum", "lts__t_sectors_lookup_miss.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
             "launch__grid_size", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed") or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled"):
        print("   %-90s %s %s" % (h, V[i], U[i]))
PY
head -45 gpurun_out/full_fprop_nc56_hot.txt
