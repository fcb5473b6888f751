"""Top stall-sample SASS lines of an ncu report's source page.
    ncu -i X.ncu-rep --page source --csv > /tmp/src.csv ; python tools/ncu_hot.py /tmp/src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[h]
si, src, ex = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(H) if c.startswith("stall_")]
data = []
for k, r in enumerate(rows[h + 1:]):
    if len(r) <= si:
        continue
    try:
        data.append((int(r[si]), k, r))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, " stall columns:", [H[i] for i in stall_cols][:20])
for s, k, r in sorted(data, reverse=True)[:n]:
    top = sorted(((int(r[i] or 0), H[i]) for i in stall_cols), reverse=True)[:2] if stall_cols else []
    print("%6d %5.1f%%  #%4d  %-70s %s" % (s, 100.0 * s / max(tot, 1), k, r[src].strip()[:70], top))
