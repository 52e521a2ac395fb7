#!/usr/bin/env python
"""Turn an `ncu --set full` report into a markdown table of the roofline-relevant metrics per kernel launch.

    python tools/ncu_table.py gpurun_out/prof_x.ncu-rep [--title "..."] > profiles/rNN_x_ncu.md
"""
import argparse
import csv
import io
import re
import subprocess

METRICS = [
    ("gpu__time_duration.sum", "duration us", lambda v, u: v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %", None),
    ("dram__bytes_read.sum", "DRAM read MB", lambda v, u: v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)),
    ("dram__bytes_write.sum", "DRAM write MB", lambda v, u: v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", None),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", None),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", None),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX %", None),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", None),
    ("smsp__inst_executed.sum", "warp instr M", lambda v, u: v * 1e-6),
    ("launch__registers_per_thread", "regs", None),
    ("launch__block_size", "threads", None),
    ("launch__grid_size", "CTAs", None),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--title", default="")
    ap.add_argument("--command", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    if a.title:
        print(f"# {a.title}\n")
    if a.command:
        print(f"Command: `{a.command}`\n")
    print("| kernel | " + " | ".join(m[1] for m in METRICS) + " | DRAM GB/s |")
    print("|---|" + "---:|" * (len(METRICS) + 1))
    for r in rows[2:]:
        name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]).replace("void ", "").replace("unnamed>::", "").replace("<unnamed>::", "")
        vals = []
        num = {}
        for key, label, conv in METRICS:
            if key not in col or r[col[key]] == "":
                vals.append("-")
                continue
            v = float(r[col[key]].replace(",", ""))
            if conv:
                v = conv(v, units[col[key]])
            num[label] = v
            vals.append(f"{v:.1f}" if v < 1000 else f"{v:.0f}")
        gbs = "-"
        if all(k in num for k in ("duration us", "DRAM read MB", "DRAM write MB")) and num["duration us"] > 0:
            gbs = f"{(num['DRAM read MB'] + num['DRAM write MB']) / num['duration us'] * 1e3:.0f}"
        print(f"| `{name}` | " + " | ".join(vals) + f" | {gbs} |")


if __name__ == "__main__":
    main()
