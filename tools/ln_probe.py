"""Time the residual GEMMs with LayerNorm warps (msclip_op_gemm_resid_ln) against the two launches they replace, at the
text / image tower shapes.  python tools/ln_probe.py [--reps 10]"""
import argparse, ctypes as C, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from msclip_b200 import _lib

def ptr(t): return C.c_void_p(t.data_ptr())

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=10); ap.add_argument("--only", default="")
    a = ap.parse_args()
    L = _lib.lib(); sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for name, M, K in (("text/out_proj", 4096 * 77, 768), ("text/fc2", 4096 * 77, 3072), ("image/out_proj", 4096 * 50, 768), ("image/fc2", 4096 * 50, 3072)):
        if a.only not in name: continue
        N = 768
        A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device="cuda"); g = torch.rand(N, device="cuda") + 0.5; be = torch.randn(N, device="cuda")
        x = torch.randn(M, N, device="cuda"); h = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        cnt = torch.zeros(L.msclip_op_gemm_resid_ln_counters(M), device="cuda", dtype=torch.int32)
        def fused(): _lib.check(L.msclip_op_gemm_resid_ln(ptr(A), K, ptr(W), K, M, N, K, ptr(b), ptr(x), N, ptr(g), ptr(be), ptr(h), N, ptr(cnt), sp))
        def gemm(): _lib.check(L.msclip_op_gemm(ptr(A), K, ptr(W), K, M, N, K, 1.0, ptr(b), ptr(x), N, ptr(x), N, _lib.EPI_RESID_F32, sp))
        def ln(): _lib.check(L.msclip_op_layernorm(ptr(x), 1, ptr(g), ptr(be), ptr(h), M, sp))
        def t(fn):
            for _ in range(2): fn()
            torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps): fn()
            e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / a.reps
        tf, tg, tl = t(fused), t(gemm), t(ln)
        print(f"{name:16s} fused {tf:.3f} ms | gemm {tg:.3f} + layernorm {tl:.3f} = {tg + tl:.3f} ms", flush=True)

if __name__ == "__main__":
    main()
