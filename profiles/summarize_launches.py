"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
usage: python profiles/summarize_launches.py <launches.csv> <title> > profiles/<name>.md"""
import collections
import csv
import re
import sys

path, title = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    m = re.search(r"lnb_items_kernel<(\w+)>", name) or re.search(r"(lnb_\w+)", name)
    key = m.group(1) if m else name.split("(")[0][:70]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r.get("Metric Unit", "ns"), 1e-6)
    agg[key][0] += 1
    agg[key][1] += float(r["Metric Value"].replace(",", "")) * scale
tot = sum(v[1] for v in agg.values()) or 1.0
print(f"# {title}\n")
print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
