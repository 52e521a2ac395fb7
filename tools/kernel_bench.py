#!/usr/bin/env python
"""Per-kernel timing on the B200 at the benchmark shapes (CUDA events, warm-up, L2 flushed between
launches by rotating over operand sets larger than the 126 MB L2).  Writes gpurun_out/kernel_bench.json.

    python tools/kernel_bench.py [--batch 4096] [--reps 10]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
from msclip_b200 import _lib                   # noqa: E402


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


WARM = 3


def time_ms(fn, reps, warm=None):
    warm = WARM if warm is None else warm
    s = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        fn()
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--modes", type=int, nargs="+", default=[1, 0], help="GEMM tiling modes to time (0, 1, 2, 4)")
    ap.add_argument("--only", default="", help="substring filter on the kernel names (profiling)")
    args = ap.parse_args()
    global WARM
    WARM = args.warm
    L = _lib.lib()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    peaks = {"tflops": 1653.1, "hbm": 6545.3}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            j = json.load(f)
        peaks = {"tflops": j["bf16_tflops"], "hbm": j["hbm_gbs"]}
    out = {"peaks": peaks, "batch": args.batch, "gemm": [], "other": []}
    B = args.batch
    shapes = []
    for tower, Lseq in (("text", 77), ("image", 50)):
        M = B * Lseq
        shapes += [(f"{tower}/qkv", M, 2304, 768, _lib.EPI_BF16), (f"{tower}/out_proj", M, 768, 768, _lib.EPI_RESID_F32),
                   (f"{tower}/fc1", M, 3072, 768, _lib.EPI_QGELU_BF16), (f"{tower}/fc2", M, 768, 3072, _lib.EPI_RESID_F32)]
    for name, M, N, K, epi in shapes:
        if args.only not in name:
            continue
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        f32 = epi in (_lib.EPI_RESID_F32, _lib.EPI_F32)
        o = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
        # comparator only (not on the product path): cuBLAS via torch.matmul on the same operands, same moment
        o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = time_ms(lambda: torch.matmul(a, w.t(), out=o16), args.reps)
        tf = 2.0 * M * N * K / ms / 1e9
        out["gemm"].append({"name": name, "M": M, "N": N, "K": K, "pair": "cublas(no epilogue)", "ms": ms, "tflops": tf})
        print(f"{name:16s} cuBLAS {ms:8.3f} ms {tf:7.1f} TF/s  (plain bf16 GEMM, no bias/activation/residual)", flush=True)
        del o16
        for pair in args.modes:
            L.msclip_op_set_gemm_pair_mode(pair)
            fn = lambda: _lib.check(L.msclip_op_gemm(ptr(a), K, ptr(w), K, M, N, K, 1.0, ptr(bias), ptr(o), N,
                                                     ptr(o) if epi == _lib.EPI_RESID_F32 else None, N, epi, sp))
            ms = time_ms(fn, args.reps)
            tf = 2.0 * M * N * K / ms / 1e9
            gb = (M * K * 2 + N * K * 2 + M * N * (4 if f32 else 2) * (2 if epi == _lib.EPI_RESID_F32 else 1)) / ms / 1e6
            out["gemm"].append({"name": name, "M": M, "N": N, "K": K, "pair": pair, "ms": ms, "tflops": tf,
                                "frac_tensor": tf / peaks["tflops"], "algorithmic_GBps": gb})
            print(f"{name:16s} pair={pair} {ms:8.3f} ms {tf:7.1f} TF/s ({100 * tf / peaks['tflops']:.1f}%)  {gb:7.0f} GB/s", flush=True)
        L.msclip_op_set_gemm_pair_mode(-1)
        del a, w, o
    # convolutions of the stem / parallel branch at one chunk of 256 images: implicit GEMM vs im2col + GEMM
    nbc = 256
    convs = [("stem0 3x3s2 48->96", 112, 48, 0, 48, 3, 2, 1, 96), ("stem1 3x3s2 96->192", 56, 96, 0, 96, 3, 2, 1, 192),
             ("stem2 3x3s2 192->384", 28, 192, 0, 192, 3, 2, 1, 384), ("stem3 3x3s2 384->768", 14, 384, 0, 384, 3, 2, 1, 768),
             ("branch1.conv2 3x3s2 48->48", 112, 48, 0, 48, 3, 2, 1, 48), ("branch2.conv2 3x3s2 96->96", 56, 96, 0, 96, 3, 2, 1, 96)]
    for name, H, cpix, coff, Cc, k, st, pd, N in convs:
        if args.only not in "conv/" + name:
            continue
        Ho = (H + 2 * pd - k) // st + 1
        x = torch.randn(nbc, H, H, cpix, device="cuda").to(torch.bfloat16)
        K = k * k * Cc
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        o = torch.empty(nbc * Ho * Ho, N, device="cuda", dtype=torch.bfloat16)
        col = torch.empty(nbc * Ho * Ho, K, device="cuda", dtype=torch.bfloat16)
        ms_i = time_ms(lambda: _lib.check(L.msclip_op_conv_gemm(ptr(x), H, H, cpix, coff, Cc, k, st, pd, None, 0, 0, 0, 0, 0, 0, 0, 0, nbc,
                                                                  Ho, Ho, ptr(w), K, N, ptr(bias), ptr(o), N, _lib.EPI_RELU_BF16, sp)), args.reps)
        kp = L.msclip_op_conv_tma_kpad(Cc, k, 0, 0)
        wp = torch.empty(N * kp, device="cuda", dtype=torch.bfloat16)
        # im2col-mode TMA feed (the default path); includes the (tiny) weight re-pack the engine does once at load time
        ms_t = time_ms(lambda: _lib.check(L.msclip_op_conv_tma(ptr(x), H, H, cpix, coff, Cc, k, st, pd, None, 0, 0, 0, 0, 0, 0, 0, 0, nbc,
                                                                 Ho, Ho, ptr(w), K, N, ptr(bias), ptr(o), N, _lib.EPI_RELU_BF16, ptr(wp), sp)),
                       args.reps)
        ms_c = time_ms(lambda: _lib.check(L.msclip_op_im2col_nhwc(ptr(x), nbc, H, H, cpix, coff, Cc, k, st, pd, ptr(col), K, 0, sp)), args.reps)
        ms_g = time_ms(lambda: _lib.check(L.msclip_op_gemm(ptr(col), K, ptr(w), K, nbc * Ho * Ho, N, K, 1.0, ptr(bias), ptr(o), N, None, 0,
                                                            _lib.EPI_RELU_BF16, sp)), args.reps)
        fl = 2.0 * nbc * Ho * Ho * N * K
        algo = (x.numel() * Cc // cpix + o.numel() + w.numel()) * 2      # read input once, write output once
        out["other"].append({"name": "conv/" + name, "ms": ms_t, "ms_gather_kernel": ms_i, "ms_im2col": ms_c, "ms_gemm": ms_g,
                             "tflops": fl / ms_t / 1e9, "GBps": algo / ms_t / 1e6, "frac_hbm": algo / ms_t / 1e6 / peaks["hbm"]})
        print(f"conv/{name:28s} tma-im2col {ms_t:7.3f} ms ({fl / ms_t / 1e9:6.1f} TF/s, {algo / ms_t / 1e6:6.0f} GB/s algorithmic = "
              f"{100 * algo / ms_t / 1e6 / peaks['hbm']:.0f}% HBM) | gather kernel {ms_i:7.3f} | im2col {ms_c:7.3f} + gemm {ms_g:7.3f} ms",
              flush=True)
        del x, w, o, col
    # fused 112 x 112 stage (front.cu) against the kernels it replaces, one chunk of 256 images
    if args.only in "front/fused":
        R, kk = 224, 16
        img = torch.randn(nbc, 3, R, R, device="cuda")
        w0 = (torch.randn(96, 32, device="cuda") / 5).to(torch.bfloat16)
        w0[:, 27:] = 0
        w1 = (torch.randn(48, 48, device="cuda") / 7).to(torch.bfloat16)
        b0, b1 = torch.randn(96, device="cuda"), torch.randn(48, device="cuda")
        pw, pb = torch.randn(kk * kk, 48, device="cuda") / kk, torch.randn(48, device="cuda")
        px = nbc * 112 * 112
        stem = torch.empty(px, 48, device="cuda", dtype=torch.bfloat16)
        y1 = torch.empty(px, 48, device="cuda", dtype=torch.bfloat16)
        p0s = torch.empty(px // 4, 48, device="cuda", dtype=torch.bfloat16)
        pooled = torch.empty(nbc * 49, 48, device="cuda", dtype=torch.bfloat16)
        ms_f = time_ms(lambda: _lib.check(L.msclip_op_front_conv(ptr(img), _lib.F32, nbc, R, R, ptr(w0), ptr(b0), ptr(w1), ptr(b1), ptr(pw),
                                                                  ptr(pb), kk, ptr(stem), ptr(y1), ptr(p0s), 48, ptr(pooled), sp)), args.reps)
        col0 = torch.empty(px, 32, device="cuda", dtype=torch.bfloat16)
        a1 = torch.empty(px, 96, device="cuda", dtype=torch.bfloat16)
        ms_u = time_ms(lambda: (_lib.check(L.msclip_op_im2col_first(ptr(img), _lib.F32, ptr(col0), nbc, R, R, sp)),
                                _lib.check(L.msclip_op_gemm(ptr(col0), 32, ptr(w0), 32, px, 96, 32, 1.0, ptr(b0), ptr(a1), 96, None, 0,
                                                            _lib.EPI_RELU_BF16, sp)),
                                _lib.check(L.msclip_op_gemm(ptr(a1[:, 48:]), 96, ptr(w1), 48, px, 48, 48, 1.0, ptr(b1), ptr(y1), 48, None, 0,
                                                            _lib.EPI_RELU_BF16, sp)),
                                _lib.check(L.msclip_op_patch_pool(ptr(a1), nbc, 112, 112, 96, 48, 48, kk, ptr(pw), ptr(pb), ptr(pooled), sp))),
                       args.reps)
        algo = img.numel() * 4 + (stem.numel() + y1.numel() + p0s.numel() + pooled.numel()) * 2
        out["other"].append({"name": "front/fused", "ms": ms_f, "ms_unfused": ms_u, "GBps": algo / ms_f / 1e6,
                             "frac_hbm": algo / ms_f / 1e6 / peaks["hbm"]})
        print(f"front/fused {ms_f:7.3f} ms ({algo / ms_f / 1e6:6.0f} GB/s algorithmic, {100 * algo / ms_f / 1e6 / peaks['hbm']:.1f}% of HBM) | "
              f"im2col + GEMM(96) + GEMM(48) + pool {ms_u:7.3f} ms", flush=True)
        del img, stem, y1, p0s, col0, a1
    # LayerNorm (HBM-bound): M x 768 fp32 in, bf16 out
    for tower, Lseq in (("text", 77), ("image", 50)):
        if args.only not in f"{tower}/layernorm":
            continue
        M = B * Lseq
        x = torch.randn(M, 768, device="cuda")
        w, b = torch.randn(768, device="cuda"), torch.randn(768, device="cuda")
        y = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)
        ms = time_ms(lambda: _lib.check(L.msclip_op_layernorm(ptr(x), 1, ptr(w), ptr(b), ptr(y), M, sp)), args.reps)
        gb = M * 768 * 6 / ms / 1e6
        out["other"].append({"name": f"{tower}/layernorm", "ms": ms, "GBps": gb, "frac_hbm": gb / peaks["hbm"]})
        print(f"{tower}/layernorm {ms:8.3f} ms {gb:7.0f} GB/s ({100 * gb / peaks['hbm']:.1f}% of HBM)", flush=True)
        del x, y
    # attention
    for tower, Lseq, causal in (("text", 77, 1), ("image", 50, 0)):
        if args.only not in f"{tower}/attention":
            continue
        M = B * Lseq
        qkv = (torch.randn(M, 2304, device="cuda") * 0.5).to(torch.bfloat16)
        o = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)
        ms = time_ms(lambda: _lib.check(L.msclip_op_attention(ptr(qkv), ptr(o), B, Lseq, 12, causal, sp)), args.reps)
        fl = 4.0 * Lseq * Lseq * 768 * B
        gb = M * (2304 + 768) * 2 / ms / 1e6
        out["other"].append({"name": f"{tower}/attention", "ms": ms, "tflops": fl / ms / 1e9, "GBps": gb, "frac_hbm": gb / peaks["hbm"]})
        print(f"{tower}/attention {ms:8.3f} ms {fl / ms / 1e9:7.1f} TF/s {gb:7.0f} GB/s ({100 * gb / peaks['hbm']:.1f}% of HBM)", flush=True)
        del qkv, o
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
