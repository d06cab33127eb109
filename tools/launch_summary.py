"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the run)."""
import collections
import csv
import re
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ci = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) < len(h) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("ps3d::", "")
    val = float(r[ci["Metric Value"]].replace(",", ""))
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[r[ci["Metric Unit"]]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(v[1] for v in agg.values())
lines = [f"# {title}", "# ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
         f"# total {tot / 1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches",
         "kernel,launches,total_ms,share_pct,avg_ms"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k},{v[0]},{v[1] / 1e6:.3f},{100 * v[1] / tot:.1f},{v[1] / 1e6 / v[0]:.4f}")
open(dst, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
