#!/usr/bin/env python
"""
Turn `ncu --set full` captures (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into the
table bench.py reads for `roofline.traffic`: profiles/r2_traffic.json.

    python scripts/ncu_traffic.py [--tag r2z] 7.2e9 gpurun_out/prof_r2_build.ncu-rep gpurun_out/prof_r2_products.ncu-rep

Per kernel: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) and the
algorithmic bytes of that launch (G bytes the launch writes or reads: rows x cols x 8, from the
grid the capture script used: scripts/prof_one.py tmi_vector 300 -> 300 x 3e6 x 8 B).
Also writes a markdown summary of the headline counters next to it.
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
]


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1.0)


def read(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        kernels.append(d)
    return kernels


def main():
    args = list(sys.argv[1:])
    tag = "r2"
    if args[0] == "--tag":                      # profiles/<tag>_traffic.json, <tag>_kernels_summary.md
        tag = args[1]
        args = args[2:]
    algorithmic = float(args[0])                # bytes of G one captured launch touches
    table, md = {"kernels": []}, [f"# ncu --set full, {tag} (one launch each; cold cache, profiler clocks)\n"]
    for rep in args[1:]:
        for d in read(rep):
            name = d["Kernel Name"][0]
            rd = to_bytes(*d["dram__bytes_read.sum"])
            wr = to_bytes(*d["dram__bytes_write.sum"])
            short = "pf_march_kernel" if "pf_march" in name else ("gemv_kernel" if "gemv" in name else
                    ("column_kernel" if "column" in name else name.split("(")[0].split("::")[-1]))
            table["kernels"].append({"match": short, "kernel": name, "dram_bytes": rd + wr, "dram_read": rd,
                                     "dram_write": wr, "algorithmic_bytes": algorithmic,
                                     "source": f"profiles/{tag}_kernels_summary.md (ncu --set full, {Path(rep).name})"})
            md.append(f"\n## {name}\n\n| metric | value | unit |\n|---|---|---|")
            for k in WANT:
                if k in d:
                    md.append(f"| `{k}` | {d[k][0]} | {d[k][1]} |")
            for k in sorted(d):
                if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k][0] or 0) >= 0.05:
                    md.append(f"| `{k}` | {d[k][0]} | {d[k][1]} |")
    (ROOT / "profiles" / f"{tag}_traffic.json").write_text(json.dumps(table, indent=1) + "\n")
    (ROOT / "profiles" / f"{tag}_kernels_summary.md").write_text("\n".join(md) + "\n")
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
