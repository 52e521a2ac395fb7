"""Does a pinned host->device copy on a side stream slow down concurrent kernels on this box?  (diagnostic)"""
import time
import torch
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
x = torch.randn(1 << 28, device="cuda")          # 1 GiB fp32 for a memory-bound kernel
host = torch.empty(2468773888 // 4, dtype=torch.float32).pin_memory()
dev = torch.empty_like(host, device="cuda")
side = torch.cuda.Stream()


def run(kind, with_copy, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if with_copy:
        with torch.cuda.stream(side):
            for _ in range(3):
                dev.copy_(host, non_blocking=True)
    e0.record()
    for _ in range(iters):
        if kind == "gemm":
            torch.matmul(a, b)
        else:
            x.mul_(1.0001)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for kind, iters in (("gemm", 150), ("stream", 300)):
    run(kind, False, 10)
    t0 = run(kind, False, iters)
    t1 = run(kind, True, iters)
    print(f"{kind}: alone {t0:.1f} ms, with 3 x 2.47 GB concurrent H2D copies {t1:.1f} ms")
torch.cuda.synchronize()
t = time.perf_counter()
dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
print(f"copy alone: {(time.perf_counter() - t) * 1e3:.1f} ms")
