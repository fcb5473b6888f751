#!/bin/bash
# ncu --set full of the student's first layer: stem convolution (forward), its filter gradient, pooling forward with BN + ReLU folded in
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -o gpurun_out/full_$1 -f python tools/op_breakdown.py 256 > gpurun_out/ncu_full_$1.log 2>&1; echo "ncu $1 exit=$?"
  ncu -i gpurun_out/full_$1.ncu-rep --page raw --csv > gpurun_out/full_$1.csv 2>/dev/null
  ncu -i gpurun_out/full_$1.ncu-rep --page source --csv > gpurun_out/full_$1_src.csv 2>/dev/null
  python tools/ncu_hot.py gpurun_out/full_$1_src.csv 14 > gpurun_out/full_$1_hot.txt 2>&1
}
cap stem_fwd "conv_fprop_kernel<.int.32, .bool.0" 1
cap stem_wgrad "conv_wgrad_kernel" 15
cap pool1_fwd "maxpool_fwd_h2_kernel<.bool.1, .int.3, .int.3, .bool.1" 1
python - <<'PY'
import csv
for n in ("stem_fwd", "stem_wgrad", "pool1_fwd"):
    try:
        rows = list(csv.reader(open("gpurun_out/full_%s.csv" % n)))
        H, U, V = rows[0], rows[1], rows[2]
        print("==", n)
        for i, h in enumerate(H):
            if h in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
                     "lts__t_sectors_lookup_hit.sum", "lts__t_sectors_lookup_miss.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
                     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
                     "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"):
                print("   %-90s %s %s" % (h, V[i], U[i]))
        print(open("gpurun_out/full_%s_hot.txt" % n).read())
    except Exception as e:
        print(n, "unreadable", e)
PY
