#!/usr/bin/env python
"""Kernel timeline of un-profiled-speed bench steps (nsys is not in the image: torch.profiler / CUPTI activity records
see every kernel of the process, including the ones libmsclip_b200.so launches).

    python tools/timeline.py [--steps 3] [--batch 4096] > profiles/rNN_timeline.md

Reports, per step: wall time from the first kernel's start to the last kernel's end, summed kernel time, idle time
between kernels (launch gaps + drain / ramp), the gap distribution and the largest gaps with the kernels around them."""
import argparse
import ctypes as C
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
from torch.profiler import ProfilerActivity, profile   # noqa: E402
from msclip_b200 import _lib, synth           # noqa: E402
from msclip_b200.config import MSCLIPConfig   # noqa: E402
from msclip_b200.model import CLIP            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--layers", type=int, default=12)
    a = ap.parse_args()
    cfg = MSCLIPConfig(layers=a.layers)
    sd = synth.synth_state_dict(cfg, seed=0)
    model = CLIP(cfg)
    model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    model = model.cuda().eval()
    model._sync_weights()
    B = a.batch
    img = torch.randn(B, 3, 224, 224, device="cuda")
    tok = torch.from_numpy(synth.synth_tokens(B, 1234)).cuda()
    parts, loss = torch.zeros(2, device="cuda"), torch.zeros((), device="cuda")
    L, h = _lib.lib(), model._handle
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        _lib.check(L.msclip_forward_loss(h, C.c_void_p(img.data_ptr()), _lib.F32, C.c_void_p(tok.data_ptr()), B,
                                         C.c_void_p(parts.data_ptr()), C.c_void_p(loss.data_ptr()), sp))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
            ev.append((e.time_range.start, e.time_range.end, e.name))
    ev.sort()
    # split into steps at text_embed_kernel (first kernel of a step)
    starts = [i for i, (_, _, n) in enumerate(ev) if "text_embed" in n]
    print(f"# Kernel timeline of {a.steps} bench steps at full speed (torch.profiler / CUPTI; batch {B}, {a.layers} layers)\n")
    print("| step | kernels | wall ms | summed kernel ms | idle ms | idle % | median gap us | gaps > 10 us |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    gap_after = defaultdict(list)
    for si, lo in enumerate(starts):
        hi = starts[si + 1] if si + 1 < len(starts) else len(ev)
        seg = ev[lo:hi]
        wall = (seg[-1][1] - seg[0][0]) / 1e3
        busy = sum(e - s for s, e, _ in seg) / 1e3
        gaps = [(seg[i + 1][0] - seg[i][1], seg[i][2], seg[i + 1][2]) for i in range(len(seg) - 1)]
        gs = sorted(g for g, _, _ in gaps)
        for g, prev, nxt in gaps:
            gap_after[prev.split("<")[0].split("(")[0][-40:]].append(g)
        print(f"| {si} | {len(seg)} | {wall:.2f} | {busy:.2f} | {wall - busy:.2f} | {100 * (wall - busy) / wall:.1f} | "
              f"{gs[len(gs) // 2]:.1f} | {sum(1 for g in gs if g > 10)} |")
    print("\n| gap follows kernel | count | mean gap us | total ms |\n|---|---:|---:|---:|")
    for k, v in sorted(gap_after.items(), key=lambda kv: -sum(kv[1]))[:12]:
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / 1e3:.2f} |")


if __name__ == "__main__":
    main()
