"""Aggregate an ncu `--page source --print-source cuda,sass --csv` export by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur = None; agg = {}; hdr = None
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; ci = {n: i for i, n in enumerate(hdr)}; continue
    if hdr is None or len(r) < 10 or r[0] == '': continue
    try:
        samples = int(r[ci['# Samples']]); inst = int(r[ci['Instructions Executed']])
        exc = int(r[ci['L1 Wavefronts Shared Excessive']])
    except Exception: continue
    agg[(cur, int(r[0]), r[1].strip()[:84])] = (samples, inst, exc)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values()); te = sum(v[2] for v in agg.values())
print("total samples", ts, "warp inst", ti, "excess smem wavefronts", te)
byfile = {}
for (f, l, s), v in agg.items():
    b = byfile.setdefault(f, [0, 0]); b[0] += v[0]; b[1] += v[1]
print({k: (round(100 * v[0] / ts, 1), round(100 * v[1] / ti, 1)) for k, v in byfile.items()})
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:14s}{k[1]:4d} {100 * v[0] / ts:5.1f}%s {100 * v[1] / ti:5.1f}%i  {k[2]}")
