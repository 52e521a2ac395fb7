#!/usr/bin/env python
"""gpurun_out/kernel_bench.json (tools/kernel_bench.py) [+ gpurun_out/train_bench.json] -> markdown tables for profiles/."""
import json
import sys

kb = json.load(open(sys.argv[1]))
pk = kb["peaks"]
print("# Round 2 - per-kernel timing at the benchmark shapes (CUDA events, no profiler)\n")
print(f"Command: `python tools/kernel_bench.py` on the B200 (batch {kb['batch']}; text M = {kb['batch'] * 77} tokens, image M = {kb['batch'] * 50}; "
      f"convolutions and the fused front kernel on one chunk of 256 images).  Peaks: {pk['tflops']} TFLOP/s bf16 burst, {pk['hbm']} GB/s HBM "
      "(MEASURED_PEAKS.json).  `cuBLAS` rows are a comparator run in the same process at the same moment (torch.matmul: plain bf16 GEMM "
      "without bias / activation / residual), not part of the product path.  Everything runs back to back under the 1 kW power cap.\n")
print("| GEMM (M x N x K) | kernel / tiling | ms | TFLOP/s | of measured burst peak | vs cuBLAS plain |\n|---|---|---:|---:|---:|---:|")
cub = {}
for g in kb["gemm"]:
    tag = f"{g['name']} {g['M']}x{g['N']}x{g['K']}"
    if isinstance(g["pair"], str):
        cub[g["name"]] = g["tflops"]
        print(f"| {tag} | cuBLAS (no epilogue) | {g['ms']:.3f} | {g['tflops']:.0f} | {100 * g['tflops'] / pk['tflops']:.1f}% |  |")
    else:
        kind = {1: "ours, CTA pair 256x256 + fused epilogue (default)", 0: "ours, single CTA 128x256 + fused epilogue"}.get(g["pair"], f"ours, pair mode {g['pair']}")
        print(f"| {tag} | {kind} | {g['ms']:.3f} | {g['tflops']:.0f} | {100 * g['tflops'] / pk['tflops']:.1f}% | {100 * g['tflops'] / cub.get(g['name'], g['tflops']):.0f}% |")
print("\n| kernel | ms | TFLOP/s | algorithmic GB/s | of HBM peak | notes |\n|---|---:|---:|---:|---:|---|")
for o in kb["other"]:
    notes = []
    for k in ("ms_gather_kernel", "ms_im2col", "ms_gemm", "ms_unfused"):
        if k in o:
            notes.append(f"{k[3:]} {o[k]:.3f} ms")
    print(f"| {o['name']} | {o['ms']:.3f} | {o.get('tflops', 0):.1f} | {o.get('GBps', 0):.0f} | {100 * o.get('frac_hbm', 0):.0f}% | {', '.join(notes)} |")
if len(sys.argv) > 2:
    tb = json.load(open(sys.argv[2]))
    print("\n## Backward kernels (tools/train_bench.py, same box)\n")
    print("| kernel | ms | TFLOP/s | of burst peak | cuBLAS dY^T.X ms | GB/s | of HBM peak |\n|---|---:|---:|---:|---:|---:|---:|")
    for k in tb["kernels"]:
        print(f"| {k['name']} | {k['ms']:.3f} | {k.get('tflops', 0):.0f} | {100 * k.get('frac_tensor', 0):.1f}% | "
              f"{k.get('cublas_ms', 0):.3f} | {k.get('GBps', 0):.0f} | {100 * k.get('frac_hbm', 0):.0f}% |")
    if "train_step" in tb:
        print("\nTraining step: `" + json.dumps(tb["train_step"]) + "`")
