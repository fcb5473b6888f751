"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of one step.
    python tools/launch_summary.py gpurun_out/launches.csv [kernels_per_step] [marker]"""
import collections
import csv
import sys

path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 236
marker = sys.argv[3] if len(sys.argv) > 3 else "face_u8"
rows = list(csv.reader(open(path)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
    seq.append((r[ki].split("(")[0][:64], v))
starts = [i for i, (n, _) in enumerate(seq) if marker in n]
start = starts[1] if len(starts) > 1 and starts[1] + per_step <= len(seq) else starts[0]
step = seq[start:start + per_step]
tot = sum(v for _, v in step)
agg = collections.OrderedDict()
for n, v in step:
    a = agg.setdefault(n, [0.0, 0])
    a[0] += v
    a[1] += 1
print("launches in file: %d; step of %d kernels starting at #%d: %.3f ms (serialised, cold-cache)" % (len(seq), len(step), start, tot))
for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-66s %8.3f ms %4d launches %5.1f%%" % (n, v, c, 100 * v / tot))
