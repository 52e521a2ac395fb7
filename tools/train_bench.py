#!/usr/bin/env python
"""Backward kernels and the whole training step on the B200 at the benchmark shapes.

  * per kernel: weight-gradient GEMM (wgrad.cu) on the four linear layers of a block, attention backward, LayerNorm backward,
    QuickGELU backward - CUDA events, warm-up, operands far larger than the 126 MB L2
  * whole step: fused forward + loss, loss backward, tower backward (both towers, all blocks), fused AdamW + weight re-pack,
    MS-CLIP-S B/32, 12 layers, `--batch` pairs; algorithmic work = 3 x the forward's 23.549 GFLOP per pair for the transformer
    part (forward + dgrad + wgrad; the frozen convolutional front runs forward only)
Writes gpurun_out/train_bench.json.

    python tools/train_bench.py [--batch 4096] [--reps 5] [--skip-step]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
from msclip_b200 import _lib, synth            # noqa: E402
from msclip_b200.config import MSCLIPConfig    # noqa: E402


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def time_ms(fn, reps, warm=2):
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
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--patch", type=int, default=32, choices=[16, 32])
    ap.add_argument("--global-batch", type=int, default=0, help="training step over this many pairs on ONE GPU in micro-batches of "
                    "--batch pairs (GradCache-style: two forwards per micro-batch, one loss over all pairs); images held as bf16")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-step", action="store_true")
    ap.add_argument("--skip-kernels", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--profile-step", action="store_true", help="one warm step, then ONE step between cudaProfilerStart/Stop "
                    "(ncu --profile-from-start off)")
    args = ap.parse_args()
    L = _lib.lib("bf16")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    peaks = {"tflops": 1610.3, "tflops_sustained": 1373.3, "hbm": 6534.1}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            j = json.load(f)
        peaks = {"tflops": j["bf16_tflops"], "tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "hbm": j["hbm_gbs"]}
    out = {"peaks": peaks, "batch": args.batch, "kernels": []}
    B = args.batch
    if not args.skip_kernels:
        for tower, Lseq, causal in (("text", 77, 1), ("image", 50, 0)):
            M = B * Lseq
            for name, N, K in (("wgrad_qkv", 2304, 768), ("wgrad_out_proj", 768, 768), ("wgrad_fc1", 3072, 768), ("wgrad_fc2", 768, 3072)):
                if args.only not in f"{tower}/{name}":
                    continue
                dy = torch.randn(M, N, device="cuda").bfloat16()
                x = torch.randn(M, K, device="cuda").bfloat16()
                dw = torch.zeros(N, K, device="cuda")
                ws = torch.empty(L.msclip_op_wgrad_workspace(M, N, K), dtype=torch.uint8, device="cuda")
                ms = time_ms(lambda: _lib.check(L.msclip_op_wgrad(ptr(dy), N, ptr(x), K, M, N, K, ptr(dw), 1, ptr(ws), sp)), args.reps)
                # comparator: cuBLAS on the same operands (torch.matmul with a transposed view)
                o = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
                ms_c = time_ms(lambda: torch.matmul(dy.t(), x, out=o), args.reps)
                tf = 2.0 * M * N * K / ms / 1e9
                out["kernels"].append({"name": f"{tower}/{name}", "tokens": M, "N": N, "K": K, "ms": ms, "tflops": tf,
                                       "frac_tensor": tf / peaks["tflops"], "cublas_ms": ms_c, "cublas_tflops": 2.0 * M * N * K / ms_c / 1e9})
                print(f"{tower}/{name:15s} {ms:8.3f} ms {tf:7.1f} TF/s ({100 * tf / peaks['tflops']:.1f}%) | cuBLAS dy^T.x (bf16 out) {ms_c:8.3f} ms "
                      f"{2.0 * M * N * K / ms_c / 1e9:7.1f} TF/s", flush=True)
                del dy, x, dw, ws, o
            if args.only in f"{tower}/attention_bwd":
                qkv = (torch.randn(M, 2304, device="cuda") * 0.5).bfloat16()
                dctx = (torch.randn(M, 768, device="cuda") * 0.1).bfloat16()
                dqkv = torch.empty(M, 2304, device="cuda", dtype=torch.bfloat16)
                ms = time_ms(lambda: _lib.check(L.msclip_op_attention_bwd(ptr(qkv), ptr(dctx), ptr(dqkv), B, Lseq, 12, causal, sp)), args.reps)
                fl = 10.0 * Lseq * Lseq * 768 * B          # five L x L x 64 contractions per head
                gb = M * (2304 + 768 + 2304) * 2 / ms / 1e6
                out["kernels"].append({"name": f"{tower}/attention_bwd", "ms": ms, "tflops": fl / ms / 1e9, "GBps": gb, "frac_hbm": gb / peaks["hbm"]})
                print(f"{tower}/attention_bwd   {ms:8.3f} ms {fl / ms / 1e9:7.1f} TF/s {gb:7.0f} GB/s ({100 * gb / peaks['hbm']:.1f}% of HBM)", flush=True)
                del qkv, dctx, dqkv
            if args.only in f"{tower}/layernorm_bwd":
                x = torch.randn(M, 768, device="cuda")
                dy = torch.randn(M, 768, device="cuda")
                dx = torch.randn(M, 768, device="cuda")
                g = torch.randn(768, device="cuda")
                d16 = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)
                dg, db, dc = [torch.zeros(768, device="cuda") for _ in range(3)]
                ws = torch.empty(L.msclip_op_bwd_workspace(M), dtype=torch.uint8, device="cuda")
                ms = time_ms(lambda: _lib.check(L.msclip_op_layernorm_bwd(ptr(x), ptr(dy), ptr(g), ptr(dx), ptr(d16), ptr(dg), ptr(db), ptr(dc), M, 1,
                                                                          ptr(ws), sp)), args.reps)
                gb = M * 768 * (4 + 4 + 4 + 4 + 2) / ms / 1e6
                out["kernels"].append({"name": f"{tower}/layernorm_bwd", "ms": ms, "GBps": gb, "frac_hbm": gb / peaks["hbm"]})
                print(f"{tower}/layernorm_bwd   {ms:8.3f} ms {gb:7.0f} GB/s ({100 * gb / peaks['hbm']:.1f}% of HBM)", flush=True)
                del x, dy, dx, d16
            if args.only in f"{tower}/qgelu_bwd":
                u = torch.randn(M, 3072, device="cuda").bfloat16()
                da = torch.randn(M, 3072, device="cuda").bfloat16()
                db = torch.zeros(3072, device="cuda")
                ws = torch.empty(L.msclip_op_bwd_workspace(M), dtype=torch.uint8, device="cuda")
                ms = time_ms(lambda: _lib.check(L.msclip_op_qgelu_bwd(ptr(da), ptr(u), ptr(db), M, 3072, ptr(ws), sp)), args.reps)
                gb = M * 3072 * 6 / ms / 1e6
                out["kernels"].append({"name": f"{tower}/qgelu_bwd", "ms": ms, "GBps": gb, "frac_hbm": gb / peaks["hbm"]})
                print(f"{tower}/qgelu_bwd       {ms:8.3f} ms {gb:7.0f} GB/s ({100 * gb / peaks['hbm']:.1f}% of HBM)", flush=True)
                del u, da
    if not args.skip_step:
        from msclip_b200.model import CLIP
        from msclip_b200.optim import AdamW
        cfg = MSCLIPConfig(patch_size=args.patch, layers=args.layers)
        sd = synth.synth_state_dict(cfg, seed=0, logit_scale=2.6593)
        model = CLIP(cfg, precision="bf16")
        model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
        model = model.cuda().eval()
        img = torch.randn(B, 3, 224, 224, device="cuda")
        tok = torch.from_numpy(synth.synth_tokens(B, 1234, cfg.context_length, cfg.vocab_size)).cuda()
        if args.global_batch:
            G = args.global_batch
            del img, tok
            img = torch.randn(G, 3, 224, 224, device="cuda", dtype=torch.bfloat16)
            tok = torch.from_numpy(synth.synth_tokens(G, 1234, cfg.context_length, cfg.vocab_size)).cuda()
            model.enable_training()
            model.setup_data_parallel(G)
            opt = AdamW(model, lr=1e-4, weight_decay=0.05, lr_share=1e-4, wd_share=0.2)
            losses = []

            def gstep():
                opt.zero_grad()
                losses.append(model.loss_and_backward(img, tok, micro_batch=B))
                opt.step()
            ms = time_ms(gstep, max(1, args.reps // 2), warm=1)
            gf_pair, gf_conv = (23.549, 2.379) if args.patch == 32 else (49.617, 4.330)
            gf_step = 4.0 * (gf_pair - gf_conv) + 2.0 * gf_conv      # two forwards + dgrad + wgrad of the transformer part, two of the front
            res = {"global_batch": G, "micro_batch": B, "step_ms": ms, "pairs_per_s": G / ms * 1e3, "gflop_per_pair_executed": gf_step,
                   "tflops_executed": G / ms * gf_step, "loss_first": float(losses[0]), "loss_last": float(losses[-1]), "ln_G": __import__("math").log(G),
                   "device_bytes": int(L.msclip_device_bytes(model._handle)), "free_bytes": int(torch.cuda.mem_get_info()[0])}
            out["train_step_global_batch"] = res
            print(json.dumps(res), flush=True)
            with open(os.path.join(ROOT, "gpurun_out", "train_bench_global.json"), "w") as f:
                json.dump(out, f, indent=1)
            return
        if args.profile_step:
            model.enable_training()
            opt = AdamW(model, lr=1e-4, weight_decay=0.05, lr_share=1e-4, wd_share=0.2)
            for i in range(2):
                if i == 1:
                    torch.cuda.synchronize()
                    torch.cuda.cudart().cudaProfilerStart()
                opt.zero_grad()
                model.loss_and_backward(img, tok)
                opt.step()
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
            return
        torch.cuda.empty_cache()
        fwd_ms = time_ms(lambda: model.contrastive_loss(img, tok), args.reps)
        model.enable_training()
        opt = AdamW(model, lr=1e-4, weight_decay=0.05, lr_share=1e-4, wd_share=0.2)
        fwd_tape_ms = time_ms(lambda: model.contrastive_loss(img, tok), args.reps)
        losses = []

        def step():
            opt.zero_grad()
            losses.append(model.loss_and_backward(img, tok))
            opt.step()

        launches0 = model.launch_count()
        t0 = time.time()
        step_ms = time_ms(step, args.reps, warm=2)
        wall = time.time() - t0

        # per-step device times (a bimodal step time would otherwise hide in the mean; no subprocesses here: forking a process
        # that maps 50 GB of device memory stalls the launching thread for hundreds of milliseconds)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
        evs[0].record()
        for i in range(8):
            step()
            evs[i + 1].record()
        torch.cuda.synchronize()
        each = [evs[i].elapsed_time(evs[i + 1]) for i in range(8)]
        print("per-step ms:", [round(x, 1) for x in each], flush=True)

        def fb():
            opt.zero_grad()
            model.loss_and_backward(img, tok)
        fb_ms = time_ms(fb, args.reps, warm=1)
        # algorithmic work: transformer part forward + dgrad + wgrad (3 x), convolutional front forward only
        gf_pair, gf_conv = (23.549, 2.379) if args.patch == 32 else (49.617, 4.330)
        gf_step = 3.0 * (gf_pair - gf_conv) + gf_conv
        pairs_s = B / step_ms * 1e3
        res = {"batch": B, "layers": args.layers, "forward_loss_ms": fwd_ms, "forward_loss_taped_ms": fwd_tape_ms,
               "forward_backward_ms": fb_ms, "step_ms": step_ms, "optimizer_and_repack_ms": step_ms - fb_ms,
               "pairs_per_s": pairs_s, "gflop_per_pair_step": gf_step, "tflops": pairs_s * gf_step / 1e3,
               "frac_sustained_tensor": pairs_s * gf_step / 1e3 / peaks["tflops_sustained"],
               "ms_each_step": each, "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
               "device_bytes": int(L.msclip_device_bytes(model._handle)), "torch_allocated": int(torch.cuda.memory_allocated()),
               "free_bytes": int(torch.cuda.mem_get_info()[0]), "total_bytes": int(torch.cuda.mem_get_info()[1]),
               "wall_s": wall}
        out["train_step"] = res
        print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "train_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
