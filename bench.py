#!/usr/bin/env python
"""Benchmark of the MS-CLIP-S encode-and-contrast hot path (BASELINE.json metric: image-text pairs/sec,
MS-CLIP-S ViT-B/32, forward + global-batch contrastive loss).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over one batch of synthetic pairs: encode_image + encode_text +
similarity + symmetric cross-entropy.  Weak scaling: every rank owns `--batch` (4096) pairs, the global
batch is N x 4096 (32 768 at N = 8, BASELINE.json configs[3]); at N = 1 the workload is configs[1].
Under torchrun one process drives one GPU; the only cross-rank traffic on the data path is the in-kernel
NVLink read of the peers' embeddings inside the fused loss kernel.

JSON keys: see the contract in the task statement; `value` is device-resident throughput, `e2e` goes
through the public API with pinned HOST buffers (H2D of images/tokens and D2H of the loss inside the
timed region), `roofline` times the dominant kernel (the shared-block fc1 GEMM) alone with CUDA events,
`cpu_baseline` times the reference's own CLIP.forward (staged copy under baseline/_ref; the oracle port if it is
absent) on this box's host cores, `comparators` holds the same reference forward run eagerly on this GPU
(fp32 / TF32 / autocast bf16) and, at N > 1, NCCL all-gather + logits + CE against the fused loss stage.
`--global-batch 32768` is BASELINE.json's metric configuration on any N (micro-batches, one loss).
`train_step` (N = 1, bf16 build; `--no-train` skips it) is a secondary, device-timed figure: the same batch through one
TRAINING step - taped forward + fused loss, loss backward, backward of both towers, fused AdamW, weight re-pack
 - with its own algorithmic FLOP count; the headline `value` / `e2e` stay the forward metric.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

from msclip_b200 import synth
from msclip_b200.config import MSCLIPConfig

METRIC = "image-text pairs/sec (forward + contrastive loss), MS-CLIP-S ViT-B/32"
UNIT = "pairs/s"
GF_PER_PAIR = {32: 23.549e9, 16: 49.617e9}        # BASELINE.md section 3 (2*m*n*k of every GEMM/bmm/conv)
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed
# `ncu --set full` capture (never measured inside a bench run): fc1 at the text-tower shape read 0.4895 GB and
# wrote 1.887 GB = its algorithmic bytes (A + W read once, bf16 output written once)
NCU_DRAM_BYTES_PER_LAUNCH = {(4096 * 77, 3072, 768): 2376500000}
NCU_TRAFFIC_SOURCE = "profiles/r02_gemm_ncu.md (ncu --set full, fc1 M=315392 N=3072 K=768, CTA pair)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------- reference arm
def cpu_oracle_throughput(cfg, sd_np, sample: int, steps: int, warmup: int, seed: int = 1234):
    """pairs/s of the CPU oracle (fp32 torch port of the reference forward + loss) with all host threads."""
    from oracle import msclip_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.to_torch(sd_np)
    img = torch.from_numpy(synth.synth_images(sample, seed, cfg.image_resolution))
    tok = torch.from_numpy(synth.synth_tokens(sample, seed, cfg.context_length, cfg.vocab_size))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            logits = O.forward(img, tok, sd, cfg)
            loss = float(O.contrastive_loss(logits))
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    assert math.isfinite(loss)
    total = sum(times)
    return sample * len(times) / total, total / len(times), torch.get_num_threads()


def reference_module_available() -> bool:
    """The real reference (staged copy under baseline/_ref on the GPU box, tools/stage_reference.py)."""
    try:
        from oracle import ref_shim
        return ref_shim.reference_available()
    except Exception:
        return False


def cpu_reference_throughput(cfg, sd_np, sample: int, steps: int, warmup: int, seed: int = 1234):
    """pairs/s of the REAL reference module - `CLIP.forward` (M.py:3126-3155), fp32, eval mode, all host threads -
    plus the symmetric cross-entropy the north star adds on its logits."""
    import torch.nn.functional as F
    from oracle import ref_shim
    torch.set_num_threads(os.cpu_count() or 1)
    model = ref_shim.build_reference_model(cfg, state_dict=sd_np)
    img = torch.from_numpy(synth.synth_images(sample, seed, cfg.image_resolution))
    tok = torch.from_numpy(synth.synth_tokens(sample, seed, cfg.context_length, cfg.vocab_size))
    target = torch.arange(sample)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            logits = model(img, tok)
            loss = float(0.5 * (F.cross_entropy(logits, target) + F.cross_entropy(logits.t(), target)))
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    assert math.isfinite(loss)
    total = sum(times)
    return sample * len(times) / total, total / len(times), torch.get_num_threads()


def cpu_baseline(cfg, sd_np, sample, steps, warmup):
    """(value, s/step, cores, kind, description): the real reference module when it is on the box, else the port."""
    if reference_module_available():
        v, sec, cores = cpu_reference_throughput(cfg, sd_np, sample, steps, warmup)
        return v, sec, cores, "reference", (f"the reference's own CLIP.forward (lib/models/clip_openai_pe_res_v1.py, unmodified, "
                                            f"staged under baseline/_ref) + symmetric CE, fp32, {sample} pairs x {steps} timed steps")
    v, sec, cores = cpu_oracle_throughput(cfg, sd_np, sample, steps, warmup)
    return v, sec, cores, "port", f"CPU oracle (torch fp32 port of the reference forward + loss), {sample} pairs x {steps} timed steps"


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    sd_np = synth.synth_state_dict(cfg, seed=0)
    sample = args.cpu_sample
    value, sec_per_step, cores, kind, desc = cpu_baseline(cfg, sd_np, sample, max(args.steps, 3), max(args.warmup, 1))
    desc += f", {cores} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------- comparators
def eager_reference_on_gpu(cfg, sd_np, dev, batch: int, seed: int = 1234):
    """the SAME reference PyTorch forward run eagerly on this B200 (cuBLAS / cuDNN / ATen),
    in true fp32, with TF32 tensor cores, and under autocast(bf16); forward + symmetric CE, pairs/s by CUDA events."""
    import torch.nn.functional as F
    from oracle import ref_shim
    model = ref_shim.build_reference_model(cfg, state_dict=sd_np).to(dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    out = {}
    while batch >= 256:
        try:
            img = torch.randn(batch, 3, cfg.image_resolution, cfg.image_resolution, device=dev, generator=g)
            tok = torch.from_numpy(synth.synth_tokens(batch, seed, cfg.context_length, cfg.vocab_size)).to(dev)
            target = torch.arange(batch, device=dev)

            def step():
                logits = model(img, tok)
                return 0.5 * (F.cross_entropy(logits.float(), target) + F.cross_entropy(logits.float().t(), target))

            for name, tf32, amp in (("fp32", False, False), ("tf32", True, False), ("autocast_bf16", True, True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    step()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    reps = 2
                    e0.record()
                    for _ in range(reps):
                        loss = step()
                    e1.record()
                    torch.cuda.synchronize()
                sec = e0.elapsed_time(e1) / 1e3 / reps
                out[name] = {"value": batch / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "loss": float(loss)}
            out["batch"] = batch
            out["what"] = ("the reference's unmodified CLIP.forward (baseline/_ref) + symmetric CE, PyTorch eager on this GPU, "
                           "same synthetic weights; 1 warm-up + 2 timed steps per mode")
            break
        except torch.OutOfMemoryError:
            out = {}
            batch //= 2
            torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False
    del model
    torch.cuda.empty_cache()
    return out


def nccl_loss_comparator(model, L, h, dev, B, world, sp, scale):
    """The reference's exchange (two NCCL all-gathers, lib/utils/comm.py:140-154) + logits (M.py:3141) + symmetric CE in
    PyTorch, against the fused loss kernel with the in-kernel NVLink gather, on the same per-rank features.  Device
    time, max over ranks."""
    import torch.nn.functional as F
    from msclip_b200 import _lib
    from msclip_b200.comm import gather_tensors
    g = torch.Generator(device=dev).manual_seed(99 + torch.distributed.get_rank())
    fi = F.normalize(torch.randn(B, 512, device=dev, generator=g), dim=-1)
    ft = F.normalize(fi + 0.3 * torch.randn(B, 512, device=dev, generator=g), dim=-1)
    parts = torch.zeros(2, device=dev)
    target = torch.arange(B * world, device=dev)

    def ours():
        _lib.check(L.msclip_contrastive_loss_features(h, C.c_void_p(fi.data_ptr()), C.c_void_p(ft.data_ptr()), B, scale,
                                                      C.c_void_p(parts.data_ptr()), None, sp), "contrastive_loss_features")

    def nccl(tf32):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        fa, ta = gather_tensors(fi), gather_tensors(ft)
        logits = scale * fa @ ta.t()
        return 0.5 * (F.cross_entropy(logits, target) + F.cross_entropy(logits.t(), target))

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms), r

    ms_ours, _ = timed(ours)
    p = parts.clone()
    torch.distributed.all_reduce(p)
    loss_ours = float(p.sum() / (2.0 * world * B))
    ms_fp32, loss_fp32 = timed(lambda: nccl(False))
    ms_tf32, _ = timed(lambda: nccl(True))
    torch.backends.cuda.matmul.allow_tf32 = False
    return {"fused_p2p_loss_ms": ms_ours, "nccl_allgather_matmul_ce_fp32_ms": ms_fp32, "nccl_allgather_matmul_ce_tf32_ms": ms_tf32,
            "loss_fused": loss_ours, "loss_nccl_fp32": float(loss_fp32), "global_batch": B * world,
            "what": "loss stage only, same per-rank features: ours = msclip_contrastive_loss_features (publish + in-kernel NVLink "
                    "gather + fused similarity/CE); comparator = 2 x dist.all_gather + [G,G] logits + 2 x F.cross_entropy"}


def train_step_block(model, img_dev, tok_dev, B, cfg, steps: int = 4, warmup: int = 2):
    """Device-timed training step (secondary figure, not the headline metric): pairs/s and the tensor-roofline fraction with
    3 x the transformer FLOPs of the forward (forward + dgrad + wgrad) + 1 x the frozen convolutional front."""
    try:
        from msclip_b200.optim import AdamW
        torch.cuda.empty_cache()
        model.enable_training()
        opt = AdamW(model, lr=1e-4, weight_decay=0.05, lr_share=1e-4, wd_share=0.2)      # b32.yaml:32-53, b32-yfcc-msclips.yaml:13-14
        losses = []

        def step():
            opt.zero_grad()
            losses.append(model.loss_and_backward(img_dev, tok_dev))
            opt.step()

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            step()
            evs[i + 1].record()
        torch.cuda.synchronize()
        per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        ms = evs[0].elapsed_time(evs[steps]) / steps
        gf = (3.0 * (23.549 - 2.379) + 2.379) if cfg.patch_size == 32 else None      # GFLOP: BASELINE.md section 3 (pair 23.549, convs 2.379)
        pk = peaks()
        pairs_s = B / ms * 1e3
        out = {"ms_per_step": ms, "ms_each_step": per_step, "value": pairs_s, "unit": "pairs/s", "steps": steps, "warmup": warmup,
               "what": "msclip_forward_loss (taped) + msclip_contrastive_loss_backward + msclip_backward + msclip_op_adamw + msclip_update_weight",
               "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
               "device_bytes": int(model._library().msclip_device_bytes(model._handle))}
        if gf:
            out.update({"gflop_per_pair": gf, "tflops": pairs_s * gf / 1e3,
                        "frac_of_sustained_tensor_peak": pairs_s * gf / 1e3 / pk["sustained"]})
        return out
    except Exception as exc:          # noqa: BLE001 - a secondary figure must never take the headline line down
        return {"error": repr(exc)[:300]}


def workload_config(args, cfg, world):
    if args.global_batch:
        b_local = args.global_batch // world
        name = (f"MS-CLIP-S ViT-B/{cfg.patch_size} full {cfg.layers}-layer shared encoder, global batch {args.global_batch} "
                f"synthetic 224^2 + 77-token pairs: {b_local} per GPU in micro-batches of {min(args.batch, b_local)} "
                f"(embeddings retained), one contrastive loss over all {args.global_batch}")
        return {"workload": name, "per_gpu_batch": b_local, "micro_batch": min(args.batch, b_local), "global_batch": args.global_batch,
                "image": "3x224x224 f32", "tokens": cfg.context_length, "parallelism": f"dp{world}",
                "l2": "inputs (2.5 GB per micro-batch) exceed the 126 MB L2"}
    name = (f"MS-CLIP-S ViT-B/{cfg.patch_size} full {cfg.layers}-layer shared encoder, batch {args.batch} synthetic "
            f"224^2 + 77-token pairs per GPU")
    return {"workload": name, "per_gpu_batch": args.batch, "global_batch": args.batch * world, "image": "3x224x224 f32",
            "tokens": cfg.context_length, "parallelism": f"dp{world}", "l2": "inputs (2.5 GB/step) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args, cfg, rank, world, local):
    from msclip_b200 import _lib
    from msclip_b200.model import CLIP

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the MS-CLIP-S path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    # weak scaling (default): B = --batch pairs per rank and step, one msclip_forward_loss.
    # --global-batch G (strong scaling): every rank owns G / world pairs, encoded in micro-batches of --batch pairs
    # whose embeddings are retained (msclip_encode_pairs), then ONE loss over all G.
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch must be divisible by the number of GPUs")
        B = args.global_batch // world
        micro = min(args.batch, B)
        if B % micro:
            raise SystemExit("--global-batch / gpus must be a multiple of --batch (the micro-batch)")
    else:
        B = micro = args.batch
    n_micro = B // micro
    sd_np = synth.synth_state_dict(cfg, seed=0)
    model = CLIP(cfg, precision=args.precision)
    model.load_state_dict({k: torch.as_tensor(v) for k, v in sd_np.items()})
    model = model.to(dev).eval()
    model._sync_weights()
    if world > 1 or n_micro > 1:
        model.setup_data_parallel(B)

    # synthetic shard of this rank (rows rank*B .. rank*B+B of the global batch)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    img_dev = torch.randn(B, 3, cfg.image_resolution, cfg.image_resolution, device=dev, generator=g)
    tok_np = synth.synth_tokens(B, 1234, cfg.context_length, cfg.vocab_size, offset=rank * B)
    tok_dev = torch.from_numpy(tok_np).to(dev)
    parts = torch.zeros(2, device=dev)
    loss_dev = torch.zeros((), device=dev)
    L = _lib.lib(args.precision)
    h = model._handle
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    scale = C.c_float()
    _lib.check(L.msclip_logit_scale_exp(h, C.byref(scale)), "msclip_logit_scale_exp")
    img_micro_bytes = micro * 3 * cfg.image_resolution * cfg.image_resolution * 4
    tok_micro_bytes = micro * cfg.context_length * 8

    def step_device():
        if n_micro == 1:
            _lib.check(L.msclip_forward_loss(h, C.c_void_p(img_dev.data_ptr()), _lib.F32, C.c_void_p(tok_dev.data_ptr()), B,
                                             C.c_void_p(parts.data_ptr()), C.c_void_p(loss_dev.data_ptr()) if world == 1 else None,
                                             sp), "msclip_forward_loss")
            return
        for m in range(n_micro):
            _lib.check(L.msclip_encode_pairs(h, C.c_void_p(img_dev.data_ptr() + m * img_micro_bytes), _lib.F32,
                                             C.c_void_p(tok_dev.data_ptr() + m * tok_micro_bytes), micro, m * micro, sp),
                       "msclip_encode_pairs")
        _lib.check(L.msclip_contrastive_loss(h, B, scale.value, C.c_void_p(parts.data_ptr()),
                                             C.c_void_p(loss_dev.data_ptr()) if world == 1 else None, sp), "msclip_contrastive_loss")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.msclip_launch_count(h)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) / 1e3, int(L.msclip_launch_count(h) - n0)

    warmup = max(args.warmup, args.min_warmup)
    for _ in range(warmup):
        step_device()
    with ClockSampler(local) as clocks:
        sec, launches = timed(step_device, args.steps)
    clock_summary = clocks.summary()
    pairs = B * world * args.steps
    value = pairs / sec
    if world > 1:
        p = parts.clone()
        torch.distributed.all_reduce(p)
        loss_value = float(p.sum() / (2.0 * world * B))
    else:
        loss_value = float(loss_dev)

    # ---- end to end through the public C ABI with pinned HOST buffers
    e2e = None
    if not args.no_e2e:
        # pinned host inputs: the whole step when it is one batch, else two micro-batch buffers used alternately
        # (every micro-batch is still copied host -> device inside the timed region)
        n_host = min(n_micro, 2)
        img_host = torch.empty((n_host * micro,) + tuple(img_dev.shape[1:]), dtype=torch.float32).pin_memory()
        img_host.copy_(img_dev[:n_host * micro])
        tok_host = torch.from_numpy(tok_np[:n_host * micro]).pin_memory()
        out_host = torch.zeros(3).pin_memory()

        def stage_host(m=0):
            _lib.check(L.msclip_stage_images(h, C.c_void_p(img_host.data_ptr() + m * img_micro_bytes), _lib.F32, micro, sp),
                       "msclip_stage_images")

        def step_host():
            # a training loop's input pipeline: the images of the NEXT (micro-)step start their H2D copy (second staging
            # slot) before this one is issued, so every step still moves all of its image bytes over PCIe, overlapped
            # with compute; tokens (2.5 MB per 4096) and the loss read-back ride the compute stream
            if n_micro == 1:
                stage_host()
                _lib.check(L.msclip_forward_loss(h, C.c_void_p(img_host.data_ptr()), _lib.F32, C.c_void_p(tok_host.data_ptr()), B,
                                                 C.c_void_p(out_host.data_ptr()),
                                                 C.c_void_p(out_host.data_ptr() + 8) if world == 1 else None, sp),
                           "msclip_forward_loss(host)")
                return
            for m in range(n_micro):
                stage_host((m + 1) % n_host)
                _lib.check(L.msclip_encode_pairs(h, C.c_void_p(img_host.data_ptr() + (m % n_host) * img_micro_bytes), _lib.F32,
                                                 C.c_void_p(tok_host.data_ptr() + (m % n_host) * tok_micro_bytes), micro, m * micro, sp),
                           "msclip_encode_pairs(host)")
            _lib.check(L.msclip_contrastive_loss(h, B, scale.value, C.c_void_p(out_host.data_ptr()),
                                                 C.c_void_p(out_host.data_ptr() + 8) if world == 1 else None, sp),
                       "msclip_contrastive_loss(host)")

        stage_host()
        for _ in range(3):
            step_host()
        e2e_steps = max(3, min(args.steps, 10))
        sec_h, _ = timed(step_host, e2e_steps)
        loss_f32_host = float(out_host[2])
        odt_img = _lib.torch_operand_dtype(args.precision)
        e2e = {"value": B * world * e2e_steps / sec_h, "unit": UNIT, "ms_per_step": sec_h / e2e_steps * 1e3,
               "h2d_bytes_per_step": int(n_micro * (img_micro_bytes + tok_micro_bytes)), "d2h_bytes_per_step": 12 if world == 1 else 8,
               "steps": e2e_steps, "api": ("msclip_stage_images (prefetch of the next step) + msclip_forward_loss, pinned host pointers"
                                           if n_micro == 1 else
                                           "per micro-batch: msclip_stage_images (prefetch of the next one) + msclip_encode_pairs; "
                                           "then msclip_contrastive_loss; pinned host pointers"),
               "host_image_dtype": "f32"}
        # what bounds it: the raw pinned host->device rate of this box
        cp0, cp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cp0.record(stream)
        img_dev[:n_host * micro].copy_(img_host, non_blocking=True)
        cp1.record(stream)
        torch.cuda.synchronize()
        e2e["h2d_gb_per_s_measured"] = img_host.numel() * 4 / (cp0.elapsed_time(cp1) / 1e3) / 1e9
        # secondary: the same call with the images already in the 16-bit operand type on the host (the reference
        # casts with image.type(self.dtype) before the first conv, M.py:2980; the first kernel rounds fp32 pixels
        # to this type anyway, so the result is bit-identical) - half the PCIe bytes
        if n_micro == 1:
            img_host16 = torch.empty(img_dev.shape, dtype=odt_img).pin_memory()
            img_host16.copy_(img_dev)
            code16 = _lib.F16 if args.precision == "fp16" else _lib.BF16

            def step_host16():
                _lib.check(L.msclip_stage_images(h, C.c_void_p(img_host16.data_ptr()), code16, B, sp), "msclip_stage_images")
                _lib.check(L.msclip_forward_loss(h, C.c_void_p(img_host16.data_ptr()), code16, C.c_void_p(tok_host.data_ptr()), B,
                                                 C.c_void_p(out_host.data_ptr()),
                                                 C.c_void_p(out_host.data_ptr() + 8) if world == 1 else None, sp),
                           "msclip_forward_loss(host, 16-bit images)")

            _lib.check(L.msclip_stage_images(h, C.c_void_p(img_host16.data_ptr()), code16, B, sp), "msclip_stage_images")
            for _ in range(3):
                step_host16()
            sec_h16, _ = timed(step_host16, e2e_steps)
            e2e["with_16bit_host_images"] = {"value": B * world * e2e_steps / sec_h16, "ms_per_step": sec_h16 / e2e_steps * 1e3,
                                             "h2d_bytes_per_step": int(img_host16.numel() * 2 + tok_host.numel() * 8),
                                             "loss_bit_identical_to_f32_images": bool(float(out_host[2]) == loss_f32_host) if world == 1 else None}
            del img_host16
        del img_host

    # ---- roofline of the dominant kernel: the shared-block fc1 GEMM (+bias+QuickGELU) at the text-tower M
    pk = peaks()
    M, N, K = micro * cfg.context_length, 4 * cfg.width, cfg.width
    odt = _lib.torch_operand_dtype(args.precision)
    a = torch.randn(M, K, device=dev).to(odt)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(odt)
    bias = torch.randn(N, device=dev)
    o = torch.empty(M, N, device=dev, dtype=odt)

    def gemm():
        _lib.check(L.msclip_op_gemm(C.c_void_p(a.data_ptr()), K, C.c_void_p(w.data_ptr()), K, M, N, K, 1.0,
                                    C.c_void_p(bias.data_ptr()), C.c_void_p(o.data_ptr()), N, None, 0, _lib.EPI_QGELU_BF16, sp))
    for _ in range(3):
        gemm()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(stream)
    for _ in range(reps):
        gemm()
    e1.record(stream)
    torch.cuda.synchronize()
    gemm_s = e0.elapsed_time(e1) / 1e3 / reps
    gemm_tf = 2.0 * M * N * K / gemm_s / 1e12
    step_tf = value / world * GF_PER_PAIR[cfg.patch_size] / 1e12
    roofline = {"bound": "tensor", "kernel": f"gemm_tcgen05_kernel<256, QGELU> fc1 M={M} N={N} K={K}",
                "achieved": gemm_tf, "peak": pk["burst"], "unit": "TFLOP/s", "frac": gemm_tf / pk["burst"],
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((M, N, K)), "traffic_source": NCU_TRAFFIC_SOURCE,
                "algorithmic_bytes": (M * K + N * K + M * N) * 2,
                "peak_source": pk["source"] + ", burst (kernel timed alone)",
                "us_per_launch": gemm_s * 1e6,
                "whole_step": {"achieved": step_tf, "peak": pk["sustained"], "frac": step_tf / pk["sustained"],
                               "note": f"per-GPU pairs/s x {GF_PER_PAIR[cfg.patch_size] / 1e9:.3f} GFLOP/pair against the sustained cuBLAS peak"}}
    del a, w, o

    # ---- where a multi-GPU step spends its time: towers vs loss stage (peer wait + fused kernel), per rank, device-timed
    phases = None
    if world > 1:
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        barrier()
        for k in range(args.steps):
            ev[k][0].record(stream)
            for m in range(n_micro):
                _lib.check(L.msclip_encode_pairs(h, C.c_void_p(img_dev.data_ptr() + m * img_micro_bytes), _lib.F32,
                                                 C.c_void_p(tok_dev.data_ptr() + m * tok_micro_bytes), micro, m * micro, sp),
                           "msclip_encode_pairs")
            ev[k][1].record(stream)
            _lib.check(L.msclip_contrastive_loss(h, B, scale.value, C.c_void_p(parts.data_ptr()), None, sp), "msclip_contrastive_loss")
            ev[k][2].record(stream)
        barrier()
        enc = torch.tensor([sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps], device=dev)
        los = torch.tensor([sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps], device=dev)
        enc_all = [torch.empty_like(enc) for _ in range(world)]
        los_all = [torch.empty_like(los) for _ in range(world)]
        torch.distributed.all_gather(enc_all, enc)
        torch.distributed.all_gather(los_all, los)
        phases = {"encode_ms_per_rank": [round(float(x), 3) for x in enc_all],
                  "loss_stage_ms_per_rank": [round(float(x), 3) for x in los_all],
                  "what": "mean over the timed steps, CUDA events on each rank's stream: encode = both towers; loss stage = publish + "
                          "stream wait for the peers' flags (absorbs the skew between ranks: a rank whose towers finish early "
                          "waits here for the slowest one) + fused similarity / cross-entropy kernels; compare with "
                          "comparators.loss_stage_vs_nccl.fused_p2p_loss_ms (the same stage with ranks aligned by a barrier)"}

    # ---- comparators: the reference forward run eagerly on this GPU; NCCL exchange + logits + CE
    comparators = {}
    if world > 1 and not args.no_comparators:
        comparators["loss_stage_vs_nccl"] = nccl_loss_comparator(model, L, h, dev, B, world, sp, scale.value)
    if rank != 0:
        return
    if world == 1 and not args.no_comparators and reference_module_available():
        free_before = torch.cuda.mem_get_info()[0]
        comparators["reference_eager_b200"] = eager_reference_on_gpu(cfg, sd_np, dev, min(micro, 4096))
        comparators["reference_eager_b200"]["free_hbm_gb_before"] = free_before / 1e9
    # ---- secondary: one TRAINING step of the same workload (forward + loss + backward of loss, heads,
    # all transformer blocks, adapter bottom paths and embeddings + fused AdamW + weight re-pack; convolutional front frozen)
    train = None
    if world == 1 and n_micro == 1 and args.precision == "bf16" and not args.no_train:
        train = train_step_block(model, img_dev, tok_dev, B, cfg)
    cpu = None
    if not args.no_cpu:
        v, sec_step, cores, kind, desc = cpu_baseline(cfg, sd_np, args.cpu_sample, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak",
        "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic", "config": workload_config(args, cfg, world), "impl": "ours",
        "loss": loss_value, "loss_expected_ln_G": math.log(B * world),
        "e2e": e2e, "gpu_launches": launches, "clocks": clock_summary, "roofline": roofline, "cpu_baseline": cpu,
        "device_bytes": int(L.msclip_device_bytes(h)), "comparators": comparators or None, "phases": phases,
        "train_step": train,
    }
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """The driver reads ONE JSON line from stdout: route everything else that writes to fd 1 (NCCL's version banner,
    library chatter of the ranks) to stderr and keep a private handle on the real stdout for that line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="pairs per GPU per step")
    ap.add_argument("--patch", type=int, default=32, choices=[16, 32])
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: total pairs per step over all GPUs (e.g. 32768), encoded per rank in micro-batches "
                         "of --batch pairs with ONE loss over the whole global batch; 0 = weak scaling (--batch pairs per GPU)")
    ap.add_argument("--cpu-sample", type=int, default=64, help="pairs per step of the CPU reference arm")
    ap.add_argument("--no-comparators", action="store_true", help="skip the eager-PyTorch-on-B200 / NCCL comparators")
    ap.add_argument("--min-warmup", type=int, default=3, help="timing rule: at least 3 warm-up steps (lower only for profiling)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"], help="MMA operand type (library build)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step figure")
    args = ap.parse_args()
    rank, world, local = dist_env()
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    cfg = MSCLIPConfig(patch_size=args.patch, layers=args.layers)
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
    else:
        run_ours(args, cfg, rank, world, local)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
