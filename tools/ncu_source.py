"""Top source lines of each kernel in an ncu report (needs -lineinfo and --import-source on).
    python tools/ncu_source.py report.ncu-rep [top N] [kernel substring]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, order, fn, path = {}, [], None, None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
        if fn not in agg:
            agg[fn] = {}
            order.append(fn)
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and r[0] not in ("", "...") and r[0].isdigit():
        try:
            smp = float(r[hdr["# Samples"]] or 0)
            ins = float(r[hdr["Instructions Executed"]] or 0)
            thr = float(r[hdr["Thread Instructions Executed"]] or 0)
        except Exception:
            continue
        key = (path, int(r[0]), r[1].strip())
        a = agg[fn].setdefault(key, [0.0, 0.0, 0.0])
        a[0] += smp; a[1] += ins; a[2] += thr
for fn in order:
    if want and want not in fn:
        continue
    d = agg[fn]
    ts = sum(v[0] for v in d.values()) or 1.0
    ti = sum(v[1] for v in d.values()) or 1.0
    print(f"==== {fn}: samples {ts:.0f}, warp instructions {ti:.4g}")
    for (p, ln, src), v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        lanes = v[2] / v[1] if v[1] else 0
        print(f"  {100 * v[0] / ts:5.1f}% smp {100 * v[1] / ti:5.1f}% inst  lanes {lanes:4.1f}  {p}:{ln:<4d} {src[:120]}")
