"""DRAM traffic of the tcgen05 convolution launches of ONE distillation step, from an ncu csv captured with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_ ...
Writes profiles/r02_conv_traffic.json (bench.py points at it from roofline.traffic_note; the run itself reports null).
    python tools/conv_traffic.py gpurun_out/conv_traffic.csv <conv_launches_per_step, 0 = half of the capture (op_breakdown
    runs two steps)> <per_gpu_batch> [out.json]"""
import csv
import json
import sys

path, per_step, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = list(csv.reader(open(path)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[h]
ki, mi, vi, ui, idc = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("ID")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}
launch = {}
order = []
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    if r[idc] not in launch:
        launch[r[idc]] = {"kernel": r[ki].split("(")[0]}
        order.append(r[idc])
    launch[r[idc]][r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
if per_step <= 0:
    per_step = len(order) // 2
ids = order[-per_step:]  # the last complete step in the capture
tot_r = sum(launch[i].get("dram__bytes_read.sum", 0) for i in ids)
tot_w = sum(launch[i].get("dram__bytes_write.sum", 0) for i in ids)
tot_t = sum(launch[i].get("gpu__time_duration.sum", 0) for i in ids)
out = {"per_gpu_batch": batch, "conv_launches": len(ids), "dram_bytes_read": tot_r, "dram_bytes_write": tot_w,
       "dram_bytes": tot_r + tot_w, "ncu_serialised_seconds": tot_t,
       "note": "sum over the conv_fprop_kernel / conv_wgrad_kernel launches of one step (ncu, cold-cache, serialised)"}
json.dump(out, open(sys.argv[4] if len(sys.argv) > 4 else "profiles/r02_conv_traffic.json", "w"), indent=1)
print(json.dumps(out))
