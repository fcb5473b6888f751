#!/bin/bash
mkdir -p gpurun_out
M="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed_pipe_uniform.sum"
for m in 0 1; do
  for c in conv2 conv3 res4; do
    XEMO_CONV_2CTA=$m timeout 120 python tools/conv_one.py $c 256
  done
  XEMO_CONV_2CTA=$m timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_fprop --launch-skip 2 -c 1 -o gpurun_out/full_fprop_conv2_2cta$m -f python tools/conv_one.py conv2 256 > gpurun_out/ncu_full_conv2_$m.log 2>&1; echo "ncu full $m exit=$?"
  ncu -i gpurun_out/full_fprop_conv2_2cta$m.ncu-rep --page raw --csv > gpurun_out/full_fprop_conv2_2cta$m.csv 2>/dev/null
done
python - <<'PY'
import csv
for m in (0, 1):
    try:
        rows = list(csv.reader(open("gpurun_out/full_fprop_conv2_2cta%d.csv" % m)))
        H, U, V = rows[0], rows[1], rows[2]
        want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_lookup_hit.sum", "lts__t_sectors_lookup_miss.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__cluster_dim_x", "launch__grid_size"]
        print("XEMO_CONV_2CTA=%d" % m)
        for i, h in enumerate(H):
            if h in want or "pipe_tensor" in h and "pct" in h:
                print("   %-80s %s %s" % (h, V[i], U[i]))
    except Exception as e:
        print(m, "unreadable", e)
PY
