#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table
(kernel, launches, total ms, share):  python scripts/launch_summary.py in.csv out.md "<command>" """
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").strip()
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        t = float(r[iv].replace(",", "")) * scale
        n, tot = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, tot + t)
    total = sum(t for _, t in agg.values())
    lines = [f"# Launch list of `{src.split('/')[-1]}` (ncu --metrics gpu__time_duration.sum --clock-control none)\n",
             "Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n",
             f"Profiled command: `{cmd}`\n", "| kernel | launches | total ms | ms per launch | share |", "|---|---|---|---|---|"]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{name}` | {n} | {t:.3f} | {t / n:.3f} | {100 * t / total:.1f} % |")
    lines.append(f"\ntotal {total:.1f} ms over {sum(n for n, _ in agg.values())} launches")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main()
