"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of one bench step into a markdown table.

    python tools/summarize_launches.py gpurun_out/launches.csv [--step-launches N] > profiles/rNN_launches_summary.md

The capture holds warm-up + timed steps (+ the roofline GEMM repetitions) back to back; the last complete step (from its
text_embed_kernel to the end of the loss) is summarised, or the last `--step-launches` launches when given."""
import argparse
import csv
import re
import sys
from collections import OrderedDict


def group_of(name: str) -> str:
    if "attention" in name:
        return "attention"
    if "layernorm" in name and "eot" not in name:
        return "LayerNorm"
    if "adapter" in name:
        return "adapters"
    if any(k in name for k in ("conv_gemm", "front_conv", "patch_pool", "im2col")):
        return "conv path"
    m = re.search(r"gemm_tcgen05_kernel<(\d+), (\d+), (\d+), (\d+), (\d+), (\d+)(?:, (\d+))?>", name)
    if m:
        bn, epi, cg, np_, ln, ne, conv = m.groups()
        if conv == "1" or epi == "2":
            return "conv path"
        if bn == "256" and epi in ("0", "1", "3"):
            return "shared-block GEMMs"
        return "other GEMMs (projections, adapter pw-conv, last_conv)"
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--step-launches", type=int, default=0)
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("msclip::", "").replace("(anonymous namespace)::", "")
        rows.append((name.strip(), ns))
    if a.step_launches:
        rows = rows[-a.step_launches:]
    else:
        # one step = from its first kernel (text_embed_kernel) to the end of the loss (finish_loss / loss_reduce)
        starts = [i for i, (nm, _) in enumerate(rows) if "text_embed" in nm]
        if starts:
            lo = starts[-1]
            ends = [i for i in range(lo, len(rows)) if "finish_loss" in rows[i][0] or "loss_reduce" in rows[i][0]]
            rows = rows[lo:(ends[-1] + 1) if ends else len(rows)]
    total = sum(ns for _, ns in rows)
    per = OrderedDict()
    for name, ns in rows:
        c, t = per.get(name, (0, 0.0))
        per[name] = (c + 1, t + ns)
    print(f"One step = {len(rows)} launches, {total / 1e6:.1f} ms summed kernel time under ncu (serialised, cold caches).\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for name, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {c} | {t / 1e6:.2f} | {100 * t / total:.1f}% |")
    groups = OrderedDict()
    for name, (c, t) in per.items():
        g = group_of(name)
        gc, gt = groups.get(g, (0, 0.0))
        groups[g] = (gc + c, gt + t)
    print("\n| group | launches | ms | share |\n|---|---:|---:|---:|")
    for g, (c, t) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        print(f"| {g} | {c} | {t / 1e6:.1f} | {100 * t / total:.1f}% |")


if __name__ == "__main__":
    main()
