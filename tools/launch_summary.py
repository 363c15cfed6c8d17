"""Per-kernel summary of an ncu launch list (--csv --metrics gpu__time_duration.sum[,smsp__inst_executed.sum])."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
t, n = collections.defaultdict(list), collections.defaultdict(list)
for r in rows:
    (t if r[-3] == "gpu__time_duration.sum" else n)[r[4]].append(float(r[-1].replace(",", "")))
tot = sum(sum(v) for v in t.values())
for k, v in t.items():
    inst = n.get(k)
    extra = f"  inst {sum(inst) / len(inst):12.4g}" if inst else ""
    print(f"{k[:60]:60s} n={len(v):3d} mean={sum(v) / len(v) / 1e3:9.1f} us  min={min(v) / 1e3:9.1f} max={max(v) / 1e3:9.1f}  share={100 * sum(v) / tot:5.1f}%{extra}")
