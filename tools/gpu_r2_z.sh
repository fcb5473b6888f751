#!/bin/bash
# ncu --set full of the next three kernels of the step after the paired forward convolution:
# conv2's filter gradient, the student's first pooling (BN + ReLU folded in, winner recorded), conv2's BN backward apply
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 -c 1 -o gpurun_out/full_$1 -f python tools/op_breakdown.py 256 > gpurun_out/ncu_full_$1.log 2>&1; echo "ncu $1 exit=$?"
  ncu -i gpurun_out/full_$1.ncu-rep --page raw --csv > gpurun_out/full_$1.csv 2>/dev/null
}
cap wgrad_conv2 conv_wgrad_kernel 14
cap pool1_fwd "maxpool_fwd_h2_kernel" 4
cap bn_bwd_apply_conv2 bn_bwd_apply_kernel 11
cap pool1_bwd maxpool_bwd_3x3s2_h2_kernel 3
python - <<'PY'
import csv
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_lookup_hit.sum",
        "lts__t_sectors_lookup_miss.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__cycles_active.avg",
        "launch__grid_dim_x", "launch__grid_dim_y"]
for n in ("wgrad_conv2", "pool1_fwd", "bn_bwd_apply_conv2", "pool1_bwd"):
    try:
        rows = list(csv.reader(open("gpurun_out/full_%s.csv" % n)))
        H, U, V = rows[0], rows[1], rows[2]
        print("==", n)
        for i, h in enumerate(H):
            if h in want:
                print("   %-80s %s %s" % (h, V[i], U[i]))
    except Exception as e:
        print(n, "unreadable", e)
PY
