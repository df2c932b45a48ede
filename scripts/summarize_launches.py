"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python scripts/summarize_launches.py gpurun_out/launches_r1g.csv > profiles/r1g_launches_summary.md"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
reader = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for rec in reader:
    if rec.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(rec["Metric Value"].replace(",", ""))
    unit = rec.get("Metric Unit", "ns")
    ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = rec["Kernel Name"]
    name = name.split("(")[0].replace("void ", "")
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print(f"# Launch list of `{path.split('/')[-1]}` (ncu --metrics gpu__time_duration.sum --clock-control none)\n")
print("Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| `{name}` | {cnt[name]} | {ms:.3f} | {100 * ms / total:.1f} % |")
print(f"\ntotal {total:.1f} ms over {sum(cnt.values())} launches")
