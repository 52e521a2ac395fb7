"""Per-kernel parity on the B200, through the C ABI (include/msclip_b200_ops.h).

Each sm_100a kernel is compared with the operation of the reference it replaces, restated by the oracle
(oracle/msclip_oracle.py) or, for pure contractions, by a torch fp32 product of the *same* bf16-rounded
operands.  Tolerances (Frobenius-relative unless stated):
  * fp32-out GEMM / similarity:     2e-5   (fp32 accumulation, different summation order only)
  * bf16-out GEMM, attention, LN:   3e-3   (one bf16 rounding of the result = 2^-9 per element)
  * data movement (im2col):         bit-exact
"""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from msclip_b200 import _lib
from oracle import msclip_oracle as O

pytestmark = pytest.mark.gpu

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
_results = {}
PREC = "bf16"          # set per test by the autouse fixture below: every test runs against both library builds


class _L:
    """The library build under test (bf16 or fp16 MMA operands)."""
    def __getattr__(self, name):
        return getattr(_lib.lib(PREC), name)


LIB = _L()


def op_dtype():
    return _lib.torch_operand_dtype(PREC)


def tol16():
    """Frobenius-relative bound for a result rounded once to the 16-bit operand type."""
    return 3e-3 if PREC == "bf16" else 4e-4


def check(rc, what=""):
    _lib.check(rc, what, PREC)


@pytest.fixture(params=["bf16", "fp16"], autouse=True)
def precision(request):
    global PREC
    PREC = request.param
    yield request.param
    PREC = "bf16"


def _record(name, value):
    _results[f"{PREC}/{name}"] = value
    try:
        os.makedirs(OUT_DIR, exist_ok=True)
        with open(os.path.join(OUT_DIR, "parity_ops.json"), "w") as f:
            json.dump(_results, f, indent=1, sort_keys=True)
    except OSError:
        pass


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    assert _lib.device_count() >= 1, "no sm_100 device visible to libmsclip_b200"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    yield


def run_gemm(M, N, K, epi, alpha=1.0, bias=True, lda=None, a_off=0, ldo=None, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    lda = lda or K
    ldo = ldo or N
    a_full = (torch.randn(M, lda, device="cuda", generator=g)).to(op_dtype())
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(op_dtype())
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    a = a_full[:, a_off:a_off + K]
    f32_out = epi in (_lib.EPI_RESID_F32, _lib.EPI_F32)
    out = torch.full((M, ldo), 7.0, device="cuda", dtype=torch.float32 if f32_out else op_dtype())
    resid = torch.randn(M, ldo, device="cuda", generator=g) if epi == _lib.EPI_RESID_F32 else None
    if resid is not None:
        out.copy_(resid)
    ref = alpha * (a.float() @ w.float().t())
    if b is not None:
        ref = ref + b
    if epi == _lib.EPI_QGELU_BF16:
        ref = O.quick_gelu(ref)
    elif epi == _lib.EPI_RELU_BF16:
        ref = torch.relu(ref)
    elif epi == _lib.EPI_RESID_F32:
        ref = ref + resid[:, :N]
    a_ptr = C.c_void_p(a_full.data_ptr() + 2 * a_off)
    rc = LIB.msclip_op_gemm(a_ptr, lda, ptr(w), K, M, N, K, alpha, ptr(b), ptr(out), ldo,
                                   ptr(out) if resid is not None else None, ldo, epi, stream())
    check(rc, "msclip_op_gemm")
    torch.cuda.synchronize()
    got = out[:, :N].float()
    if ldo > N:     # columns beyond N must be untouched
        assert torch.all(out[:, N:].float() == (resid[:, N:] if resid is not None else 7.0))
    return rel(got, ref), float((got - ref).abs().max())


GEMM_CASES = [
    # name,            M,     N,    K,    epilogue,             kwargs
    ("qkv",            1000,  2304, 768,  _lib.EPI_BF16,        {}),
    ("out_proj",       300,   768,  768,  _lib.EPI_RESID_F32,   {}),
    ("fc1",            128,   3072, 768,  _lib.EPI_QGELU_BF16,  {}),
    ("fc2",            5173,  768,  3072, _lib.EPI_RESID_F32,   {}),
    ("many_tiles",     40000, 768,  768,  _lib.EPI_BF16,        {}),
    ("conv_first",     5000,  96,   32,   _lib.EPI_RELU_BF16,   {}),
    ("conv_stem0",     3136,  96,   432,  _lib.EPI_RELU_BF16,   {}),
    ("conv_stem1",     784,   192,  864,  _lib.EPI_RELU_BF16,   {}),
    ("conv_stem3",     49,    768,  3456, _lib.EPI_RELU_BF16,   {}),
    ("branch_1x1",     2000,  48,   48,   _lib.EPI_RELU_BF16,   dict(lda=96, a_off=48)),
    ("branch_cat_out", 777,   48,   432,  _lib.EPI_RELU_BF16,   dict(ldo=96)),
    ("branch_384",     196,   384,  384,  _lib.EPI_RELU_BF16,   {}),
    ("adapter_pw",     392,   768,  48,   _lib.EPI_F32,         dict(bias=False)),
    ("proj",           77,    512,  768,  _lib.EPI_F32,         dict(bias=False)),
    ("logits_small",   8,     8,    1536, _lib.EPI_F32,         dict(bias=False, alpha=14.2857)),
    ("logits_ragged",  1024,  1000, 1536, _lib.EPI_F32,         dict(bias=False, alpha=100.0)),
    ("ragged_bf16",    130,   200,  72,   _lib.EPI_BF16,        {}),
    ("one_row",        1,     64,   8,    _lib.EPI_F32,         {}),
    ("pair_edge",      257,   256,  64,   _lib.EPI_BF16,        {}),
    ("pair_big_k",     2048,  512,  3072, _lib.EPI_F32,         {}),
]


@pytest.mark.parametrize("name,M,N,K,epi,kw", GEMM_CASES, ids=[c[0] for c in GEMM_CASES])
def test_gemm(name, M, N, K, epi, kw):
    r, mx = run_gemm(M, N, K, epi, **kw)
    _record(f"gemm/{name}", {"rel": r, "max_abs": mx})
    tol = 2e-5 if epi in (_lib.EPI_RESID_F32, _lib.EPI_F32) else tol16()
    assert r < tol, (name, r, mx)


@pytest.mark.parametrize("mode", [0, 1, 2, 4])
@pytest.mark.parametrize("name", ["qkv", "fc1", "fc2", "many_tiles", "pair_big_k"])
def test_gemm_tile_modes_agree(name, mode):
    """256-wide tiles can run on one CTA (0), a CTA pair (1) or multicast clusters of 2 / 4 pairs: same result."""
    case = [c for c in GEMM_CASES if c[0] == name][0]
    LIB.msclip_op_set_gemm_pair_mode(mode)
    try:
        r, mx = run_gemm(*case[1:5], **case[5])
    finally:
        LIB.msclip_op_set_gemm_pair_mode(-1)
    _record(f"gemm_mode{mode}/{name}", {"rel": r, "max_abs": mx})
    assert r < (2e-5 if case[4] in (_lib.EPI_RESID_F32, _lib.EPI_F32) else tol16())


def test_gemm_rejects_bad_arguments():
    a = torch.zeros(8, 12, device="cuda", dtype=op_dtype())
    out = torch.zeros(8, 8, device="cuda")
    rc = LIB.msclip_op_gemm(ptr(a), 12, ptr(a), 12, 8, 8, 12, 1.0, None, ptr(out), 8, None, 0, _lib.EPI_F32, stream())
    assert rc != 0 and "multiples of 8" in _lib.last_error(PREC)


@pytest.mark.parametrize("M,K", [(4096 * 77, 768), (4096 + 77, 768), (600, 3072), (256, 768), (257, 768), (20000, 3072)])
def test_gemm_resid_ln_equals_gemm_then_layernorm(M, K):
    """out-proj / fc2 with the LayerNorm warps (gemm.cu, LN = 3) against the two launches they replace: the updated
    residual stream within fp32 summation-order noise of the plain residual GEMM (different epilogue chunking, same
    products), and h BIT-IDENTICAL to layernorm_kernel applied to the x the fused kernel wrote."""
    g = torch.Generator(device="cuda").manual_seed(12)
    N = 768
    a = torch.randn(M, K, device="cuda", generator=g).to(op_dtype())
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(op_dtype())
    b = 0.2 * torch.randn(N, device="cuda", generator=g)
    gamma = 1.0 + 0.3 * torch.randn(N, device="cuda", generator=g)
    beta = 0.3 * torch.randn(N, device="cuda", generator=g)
    x0 = torch.randn(M, N, device="cuda", generator=g) + 2.0 * torch.randn(M, 1, device="cuda", generator=g)
    x_f, x_p = x0.clone(), x0.clone()
    h_f = torch.full((M, N), float("nan"), device="cuda", dtype=op_dtype())
    cnt = torch.zeros(LIB.msclip_op_gemm_resid_ln_counters(M), device="cuda", dtype=torch.int32)
    for rep in range(2):                                 # twice: the counters must come back zero
        x_f.copy_(x0)
        check(LIB.msclip_op_gemm_resid_ln(ptr(a), K, ptr(w), K, M, N, K, ptr(b), ptr(x_f), N, ptr(gamma), ptr(beta), ptr(h_f), N,
                                          ptr(cnt), stream()))
        assert int(cnt.abs().sum()) == 0
    check(LIB.msclip_op_gemm(ptr(a), K, ptr(w), K, M, N, K, 1.0, ptr(b), ptr(x_p), N, ptr(x_p), N, _lib.EPI_RESID_F32, stream()))
    assert torch.equal(x_f, x_p)                         # same MMAs, same epilogue arithmetic per element
    h_ref = torch.empty(M, N, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_layernorm(ptr(x_f), 1, ptr(gamma), ptr(beta), ptr(h_ref), M, stream()))
    assert torch.equal(h_f.view(torch.int16), h_ref.view(torch.int16))
    ref = x0.double() + a.double() @ w.double().t() + b.double()
    r = rel(x_f, ref.float())
    _record(f"gemm_resid_ln/M{M}_K{K}", {"rel": r})
    assert r < 2e-5, r


def _ln_records(x):
    """Row records of the LN fold as the elementwise producers write them: shift = row mean, slice 0 = sums."""
    mean = x.mean(-1, keepdim=True)
    c = x - mean
    rec = torch.zeros(x.shape[0], 16, device=x.device)
    rec[:, 0] = mean[:, 0]
    rec[:, 4] = c.sum(-1)
    rec[:, 5] = (c * c).sum(-1)
    return rec, c


@pytest.mark.parametrize("M,N,epi", [(4096 + 77, 2304, "bf16"), (1000, 3072, "qgelu"), (77, 2304, "bf16")])
def test_gemm_ln_consume_matches_layernorm_linear(M, N, epi):
    """QKV / fc1 with the LayerNorm folded in (gemm_common.cuh) against LN (M.py:204-219) + F.linear in fp32; the rows
    carry a common-mode offset of several standard deviations, which the per-row shift has to absorb."""
    g = torch.Generator(device="cuda").manual_seed(11)
    K = 768
    x = torch.randn(M, K, device="cuda", generator=g) * (1 + torch.rand(M, 1, device="cuda", generator=g)) \
        + 5.0 * torch.randn(M, 1, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = 0.2 * torch.randn(N, device="cuda", generator=g)
    gamma = 1.0 + 0.3 * torch.randn(K, device="cuda", generator=g)
    beta = 0.3 * torch.randn(K, device="cuda", generator=g)
    rs = torch.ones(N, device="cuda")
    rs[:768] = 0.125
    wf = torch.empty(N, K, device="cuda", dtype=op_dtype())
    cs, bf = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    check(LIB.msclip_op_pack_ln_fold(ptr(w), ptr(rs), ptr(gamma), ptr(beta), ptr(b), ptr(wf), ptr(cs), ptr(bf), N, K, stream()))
    assert rel(wf.float(), w * rs[:, None] * gamma[None, :]) < tol16()
    assert torch.allclose(cs, wf.float().sum(-1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(bf, rs * (b + w @ beta), rtol=1e-4, atol=1e-4)
    rec, c = _ln_records(x)
    xc = c.to(op_dtype())
    out = torch.zeros(M, N, device="cuda", dtype=op_dtype())
    code = _lib.EPI_BF16 if epi == "bf16" else _lib.EPI_QGELU_BF16
    check(LIB.msclip_op_gemm_ln(ptr(xc), K, ptr(wf), K, M, N, K, ptr(bf), ptr(out), N, None, 0, code, 1, ptr(rec), None, None, 0,
                                ptr(cs), stream()))
    ref = F.linear(O.layer_norm(x, gamma, beta), w, b) * rs
    if epi == "qgelu":
        ref = O.quick_gelu(ref)
    r = rel(out.float(), ref)
    _record(f"gemm_ln/consume_M{M}_N{N}_{epi}", {"rel": r})
    assert r < 1.5 * tol16(), r          # operand rounding of x - shift and of W * gamma, then the output rounding


@pytest.mark.parametrize("M,K", [(4096 + 77, 768), (600, 3072), (50, 768)])
def test_gemm_ln_emit_matches_reference(M, K):
    """out-proj / fc2 with the residual epilogue that also emits the centred 16-bit copy and the row records."""
    g = torch.Generator(device="cuda").manual_seed(12)
    N = 768
    a = torch.randn(M, K, device="cuda", generator=g).to(op_dtype())
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(op_dtype())
    b = 0.2 * torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g) * 2 + 3.0 * torch.randn(M, 1, device="cuda", generator=g)
    rec_in, _ = _ln_records(x)
    rec_in[:, 0] += 0.25                                   # a stale shift: mean of x = shift + s1 / 768 must still come out
    rec_in[:, 4] -= 0.25 * N
    rec_out = torch.full((M, 16), 9.0, device="cuda")
    x_new = x.clone()
    xc = torch.zeros(M, N, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_gemm_ln(ptr(a), K, ptr(w), K, M, N, K, ptr(b), ptr(x_new), N, ptr(x_new), N, _lib.EPI_RESID_F32, 2,
                                ptr(rec_in), ptr(rec_out), ptr(xc), N, None, stream()))
    ref = x + a.float() @ w.float().t() + b
    assert rel(x_new, ref) < 2e-5
    shift = x.mean(-1)
    assert torch.allclose(rec_out[:, 0], shift, rtol=1e-4, atol=1e-4)
    cen = x_new - rec_out[:, :1]
    assert rel(xc.float(), cen) < tol16()
    s1 = rec_out[:, 4:16:2].sum(-1)
    s2 = rec_out[:, 5:16:2].sum(-1)
    assert torch.allclose(s1, cen.sum(-1), rtol=1e-3, atol=2e-2)
    assert torch.allclose(s2, (cen * cen).sum(-1), rtol=1e-4)
    # chained: the emitted copy + records reproduce LayerNorm(x_new) . W2^T through the consume epilogue
    w2 = torch.randn(256, N, device="cuda", generator=g) / math.sqrt(N)
    gamma, beta = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g), 0.1 * torch.randn(N, device="cuda", generator=g)
    wf = torch.empty(256, N, device="cuda", dtype=op_dtype())
    cs, bf = torch.empty(256, device="cuda"), torch.empty(256, device="cuda")
    check(LIB.msclip_op_pack_ln_fold(ptr(w2), None, ptr(gamma), ptr(beta), None, ptr(wf), ptr(cs), ptr(bf), 256, N, stream()))
    y = torch.zeros(M, 256, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_gemm_ln(ptr(xc), N, ptr(wf), N, M, 256, N, ptr(bf), ptr(y), 256, None, 0, _lib.EPI_BF16, 1, ptr(rec_out),
                                None, None, 0, ptr(cs), stream()))
    r = rel(y.float(), F.linear(O.layer_norm(x_new, gamma, beta), w2))
    _record(f"gemm_ln/emit_M{M}_K{K}", {"rel_chain": r})
    assert r < 1.5 * tol16(), r


@pytest.mark.parametrize("rows,stride", [(1, 1), (77, 1), (5000, 1), (64, 50)])
def test_layernorm(rows, stride):
    x = torch.randn(rows * stride, 768, device="cuda") * 3 + 0.5
    w, b = torch.randn(768, device="cuda"), torch.randn(768, device="cuda")
    y = torch.empty(rows, 768, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_layernorm(ptr(x), stride, ptr(w), ptr(b), ptr(y), rows, stream()))
    ref = O.layer_norm(x[::stride], w, b)
    r = rel(y.float(), ref)
    # compare against the bf16 rounding of the oracle too: must agree to the last bit almost everywhere
    exact = float((y == ref.to(op_dtype())).float().mean())
    _record(f"layernorm/{rows}x{stride}", {"rel": r, "bf16_exact_fraction": exact})
    assert r < tol16() and exact > 0.99


def attention_reference(qkv, B, L, H, causal):
    q, k, v = qkv.float().view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)      # [B,H,L,64]
    s = q @ k.transpose(-1, -2)
    if causal:
        s = s + O.causal_mask(L, s.device)
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, H * 64)


@pytest.mark.parametrize("B,L,causal", [(3, 50, 0), (2, 77, 1), (2, 197, 0), (1, 197, 1), (2, 1, 0), (2, 64, 1),
                                        (1, 80, 0), (5, 17, 1), (300, 50, 0)])
def test_attention(B, L, causal):
    H = 12
    qkv = (torch.randn(B * L, 3 * H * 64, device="cuda") * 0.7).to(op_dtype())
    out = torch.zeros(B * L, H * 64, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_attention(ptr(qkv), ptr(out), B, L, H, causal, stream()))
    ref = attention_reference(qkv, B, L, H, causal)
    r = rel(out.float(), ref)
    _record(f"attention/B{B}_L{L}_c{causal}", {"rel": r})
    assert r < 2 * tol16(), r      # P is rounded to 16 bits before P.V, the result once more


def test_im2col_first_bit_exact():
    B, R = 3, 224
    img = torch.randn(B, 3, R, R, device="cuda")
    out = torch.empty(B * 112 * 112, 32, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_im2col_first(ptr(img), _lib.F32, ptr(out), B, R, R, stream()))
    cols = F.unfold(img, 3, padding=1, stride=2)                  # [B, 27, 112*112], k = c*9 + ky*3 + kx
    ref = cols.transpose(1, 2).reshape(-1, 27).to(op_dtype())
    assert torch.equal(out[:, :27], ref)
    assert torch.all(out[:, 27:] == 0)
    img16 = img.to(op_dtype())
    code16 = _lib.F16 if PREC == "fp16" else _lib.BF16          # image dtype codes are independent of the operand type
    check(LIB.msclip_op_im2col_first(ptr(img16), code16, ptr(out), B, R, R, stream()))
    assert torch.equal(out[:, :27], F.unfold(img16.float(), 3, padding=1, stride=2).transpose(1, 2).reshape(-1, 27).to(op_dtype()))


@pytest.mark.parametrize("H,cpix,coff,Cc,k,s,p", [(112, 96, 0, 48, 3, 2, 1), (112, 96, 48, 48, 1, 2, 0), (28, 192, 0, 192, 3, 2, 1),
                                                (14, 384, 0, 384, 3, 1, 1), (14, 384, 0, 384, 1, 1, 0)])
def test_im2col_nhwc_bit_exact(H, cpix, coff, Cc, k, s, p):
    B = 2
    x = torch.randn(B, H, H, cpix, device="cuda").to(op_dtype())
    Ho = (H + 2 * p - k) // s + 1
    ld, off = k * k * Cc + 16, 8
    out = torch.full((B * Ho * Ho, ld), 5.0, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_im2col_nhwc(ptr(x), B, H, H, cpix, coff, Cc, k, s, p, ptr(out), ld, off, stream()))
    nchw = x[..., coff:coff + Cc].permute(0, 3, 1, 2).float()
    cols = F.unfold(nchw, k, padding=p, stride=s)                 # [B, C*k*k, Ho*Ho], index c*k*k + tap
    ref = cols.view(B, Cc, k * k, Ho * Ho).permute(0, 3, 2, 1).reshape(B * Ho * Ho, k * k * Cc).to(op_dtype())
    assert torch.equal(out[:, off:off + k * k * Cc], ref)
    assert torch.all(out[:, :off] == 5.0) and torch.all(out[:, off + k * k * Cc:] == 5.0)


CONV_CASES = [
    # name,        B, H,   cpix, coff, C,   k, s, p, N
    ("stem0",      3, 112, 96,   0,    48,  3, 2, 1, 96),
    ("stem1",      3, 56,  96,   0,    96,  3, 2, 1, 192),
    ("stem3",      5, 14,  384,  0,    384, 3, 2, 1, 768),
    ("b16_stem3",  2, 14,  384,  0,    384, 3, 1, 1, 768),
    ("branch2_48", 2, 112, 48,   0,    48,  3, 2, 1, 48),
    ("one_image",  1, 28,  192,  0,    192, 3, 2, 1, 384),
]


@pytest.mark.parametrize("name,B,H,cpix,coff,Cc,k,s,p,N", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_gemm_matches_conv2d(name, B, H, cpix, coff, Cc, k, s, p, N):
    """Implicit-GEMM conv (+bias+ReLU) against F.conv2d on the same bf16-rounded operands."""
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, H, H, cpix, device="cuda", generator=g).to(op_dtype())
    w = (torch.randn(N, Cc, k, k, device="cuda", generator=g) / math.sqrt(Cc * k * k)).to(op_dtype())
    bias = torch.randn(N, device="cuda", generator=g)
    Ho = (H + 2 * p - k) // s + 1
    wk = w.permute(0, 2, 3, 1).reshape(N, k * k * Cc).contiguous()         # K order (ky, kx, c)
    out = torch.zeros(B * Ho * Ho, N, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_conv_gemm(ptr(x), H, H, cpix, coff, Cc, k, s, p, None, 0, 0, 0, 0, 0, 0, 0, 0, B, Ho, Ho,
                                              ptr(wk), k * k * Cc, N, ptr(bias), ptr(out), N, _lib.EPI_RELU_BF16, stream()))
    nchw = x[..., coff:coff + Cc].permute(0, 3, 1, 2).float()
    ref = torch.relu(F.conv2d(nchw, w.float(), bias, stride=s, padding=p)).permute(0, 2, 3, 1).reshape(B * Ho * Ho, N)
    r = rel(out.float(), ref)
    _record(f"conv_gemm/{name}", {"rel": r})
    assert r < tol16(), r


@pytest.mark.parametrize("H,cin,stride", [(112, 48, 2), (28, 192, 2), (14, 384, 1)])
def test_conv_gemm_two_sources(H, cin, stride):
    """ConvResBlock tail (M.py:1855-1861): relu(conv1x1(y2) + conv1x1_stride(p)) as one GEMM over K = [y2 | p]."""
    B = 2
    Ho = H // stride
    g = torch.Generator(device="cuda").manual_seed(2)
    cpix, coff = (2 * cin, cin) if H == 112 else (cin, 0)
    pfeat = torch.randn(B, H, H, cpix, device="cuda", generator=g).to(op_dtype())
    y2 = torch.randn(B, Ho, Ho, cin, device="cuda", generator=g).to(op_dtype())
    w3 = (torch.randn(2 * cin, cin, device="cuda", generator=g) / math.sqrt(cin)).to(op_dtype())
    wr = (torch.randn(2 * cin, cin, device="cuda", generator=g) / math.sqrt(cin)).to(op_dtype())
    bias = torch.randn(2 * cin, device="cuda", generator=g)
    wk = torch.cat([w3, wr], dim=1).contiguous()
    out = torch.zeros(B * Ho * Ho, 2 * cin, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_conv_gemm(ptr(y2), Ho, Ho, cin, 0, cin, 1, 1, 0, ptr(pfeat), H, H, cpix, coff, cin, 1, stride, 0,
                                              B, Ho, Ho, ptr(wk), 2 * cin, 2 * cin, ptr(bias), ptr(out), 2 * cin,
                                              _lib.EPI_RELU_BF16, stream()))
    ps = pfeat[:, ::stride, ::stride, coff:coff + cin].float().reshape(-1, cin)
    ref = torch.relu(y2.float().reshape(-1, cin) @ w3.float().t() + ps @ wr.float().t() + bias)
    r = rel(out.float(), ref)
    _record(f"conv_gemm2/H{H}", {"rel": r})
    assert r < tol16(), r


TMA_CONV_CASES = CONV_CASES + [
    ("stem0_dense", 3, 112, 48,  0,    48,  3, 2, 1, 96),      # the engine's layout: dense 48-channel pixels
    ("win_of_96",   2, 56,  96,  48,   48,  3, 2, 1, 48),      # channel window with an offset
    ("b16_stride1", 2, 14,  384, 0,    384, 3, 1, 1, 768),
    ("ragged_tail", 7, 28,  192, 0,    192, 3, 2, 1, 384),     # 7 * 196 = 1372 pixels: last tile / pair half empty
    ("tiny",        1, 8,   64,  0,    64,  3, 2, 1, 96),      # fewer pixels than one tile
]


@pytest.mark.parametrize("name,B,H,cpix,coff,Cc,k,s,p,N", TMA_CONV_CASES, ids=[c[0] for c in TMA_CONV_CASES])
def test_conv_tma_matches_conv2d(name, B, H, cpix, coff, Cc, k, s, p, N):
    """Convolution with the A operand fetched by im2col-mode TMA (gemm_tcgen05_kernel<..., CONV = 1>) against F.conv2d on
    the same rounded operands, and bit-for-bit against the gather-fed implicit-GEMM kernel (same products, same
    fp32 accumulation order per k-block: the padded K columns only add exact zeros)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, H, H, cpix, device="cuda", generator=g).to(op_dtype())
    w = (torch.randn(N, Cc, k, k, device="cuda", generator=g) / math.sqrt(Cc * k * k)).to(op_dtype())
    bias = torch.randn(N, device="cuda", generator=g)
    Ho = (H + 2 * p - k) // s + 1
    wk = w.permute(0, 2, 3, 1).reshape(N, k * k * Cc).contiguous()         # K order (ky, kx, c)
    out = torch.full((B * Ho * Ho, N), float("nan"), device="cuda", dtype=op_dtype())
    kp = LIB.msclip_op_conv_tma_kpad(Cc, k, 0, 0)
    assert kp == k * k * ((Cc + 63) // 64) * 64
    scratch = torch.empty(N * kp, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_conv_tma(ptr(x), H, H, cpix, coff, Cc, k, s, p, None, 0, 0, 0, 0, 0, 0, 0, 0, B, Ho, Ho,
                                 ptr(wk), k * k * Cc, N, ptr(bias), ptr(out), N, _lib.EPI_RELU_BF16, ptr(scratch), stream()))
    nchw = x[..., coff:coff + Cc].permute(0, 3, 1, 2).float()
    ref = torch.relu(F.conv2d(nchw, w.float(), bias, stride=s, padding=p)).permute(0, 2, 3, 1).reshape(B * Ho * Ho, N)
    r = rel(out.float(), ref)
    _record(f"conv_tma/{name}", {"rel": r})
    assert torch.isfinite(out.float()).all()
    assert r < tol16(), r
    old = torch.zeros_like(out)
    check(LIB.msclip_op_conv_gemm(ptr(x), H, H, cpix, coff, Cc, k, s, p, None, 0, 0, 0, 0, 0, 0, 0, 0, B, Ho, Ho,
                                  ptr(wk), k * k * Cc, N, ptr(bias), ptr(old), N, _lib.EPI_RELU_BF16, stream()))
    assert rel(out.float(), old.float()) < 2e-3                         # different k-block boundaries -> fp32 sum order


@pytest.mark.parametrize("H,cin,stride", [(56, 96, 2), (28, 192, 2), (14, 384, 1), (14, 384, 2)])
def test_conv_tma_two_sources(H, cin, stride):
    """ConvResBlock tail (M.py:1855-1861) with both K segments fetched by im2col-mode TMA (1x1 / stride 1 and 1x1 / stride s)."""
    B = 3
    Ho = H // stride
    g = torch.Generator(device="cuda").manual_seed(2)
    pfeat = torch.randn(B, H, H, cin, device="cuda", generator=g).to(op_dtype())
    y2 = torch.randn(B, Ho, Ho, cin, device="cuda", generator=g).to(op_dtype())
    w3 = (torch.randn(2 * cin, cin, device="cuda", generator=g) / math.sqrt(cin)).to(op_dtype())
    wr = (torch.randn(2 * cin, cin, device="cuda", generator=g) / math.sqrt(cin)).to(op_dtype())
    bias = torch.randn(2 * cin, device="cuda", generator=g)
    wk = torch.cat([w3, wr], dim=1).contiguous()
    out = torch.full((B * Ho * Ho, 2 * cin), float("nan"), device="cuda", dtype=op_dtype())
    kp = LIB.msclip_op_conv_tma_kpad(cin, 1, cin, 1)
    scratch = torch.empty(2 * cin * kp, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_conv_tma(ptr(y2), Ho, Ho, cin, 0, cin, 1, 1, 0, ptr(pfeat), H, H, cin, 0, cin, 1, stride, 0,
                                 B, Ho, Ho, ptr(wk), 2 * cin, 2 * cin, ptr(bias), ptr(out), 2 * cin,
                                 _lib.EPI_RELU_BF16, ptr(scratch), stream()))
    ps = pfeat[:, ::stride, ::stride, :].float().reshape(-1, cin)
    ref = torch.relu(y2.float().reshape(-1, cin) @ w3.float().t() + ps @ wr.float().t() + bias)
    r = rel(out.float(), ref)
    _record(f"conv_tma2/H{H}_s{stride}", {"rel": r})
    assert r < tol16(), r


@pytest.mark.parametrize("H,cpix,coff,Cc,k", [(112, 96, 48, 48, 16), (56, 96, 0, 96, 8), (14, 384, 0, 384, 1), (7, 768, 0, 768, 1)])
def test_patch_pool(H, cpix, coff, Cc, k):
    B = 2
    x = torch.randn(B, H, H, cpix, device="cuda").to(op_dtype())
    wt = torch.randn(Cc, 1, k, k, device="cuda") / k
    bias = torch.randn(Cc, device="cuda")
    w_packed = wt.view(Cc, k * k).t().contiguous()                # [k*k][C]
    g = H // k
    out = torch.empty(B * g * g, Cc, device="cuda", dtype=op_dtype())
    check(LIB.msclip_op_patch_pool(ptr(x), B, H, H, cpix, coff, Cc, k, ptr(w_packed), ptr(bias), ptr(out), stream()))
    nchw = x[..., coff:coff + Cc].permute(0, 3, 1, 2).float()
    ref = F.conv2d(nchw, wt, stride=k, groups=Cc) + bias[None, :, None, None]
    ref = ref.flatten(2).transpose(1, 2).reshape(B * g * g, Cc)
    r = rel(out.float(), ref)
    _record(f"patch_pool/H{H}_k{k}", {"rel": r})
    assert r < tol16()


@pytest.mark.parametrize("B,H,W,k,img16", [(3, 224, 224, 16, False), (2, 224, 224, 8, False), (2, 64, 96, 16, False),
                                           (1, 96, 64, 8, True), (700, 32, 32, 16, False)])
def test_front_conv_matches_unfused_reference(B, H, W, k, img16):
    """Fused 112x112 stage (front.cu) against conv2d / 1x1 / depth-wise pooling on the same rounded operands:
    stem and branch first convs (M.py:1993, 2260-2273), bottleneck entry (M.py:1842-1846), strided copy of p0
    (M.py:1857) and adapter-0 patch pooling (M.py:1756)."""
    g = torch.Generator(device="cuda").manual_seed(7)
    dt = op_dtype()
    img = torch.randn(B, 3, H, W, device="cuda", generator=g)
    code = _lib.F32
    if img16:
        img = img.to(dt)
        code = _lib.F16 if PREC == "fp16" else _lib.BF16
    w0 = (torch.randn(96, 3, 3, 3, device="cuda", generator=g) / math.sqrt(27)).to(dt)
    b0 = 0.3 * torch.randn(96, device="cuda", generator=g)
    w1 = (torch.randn(48, 48, device="cuda", generator=g) / math.sqrt(48)).to(dt)
    b1 = 0.3 * torch.randn(48, device="cuda", generator=g)
    pw = torch.randn(48, 1, k, k, device="cuda", generator=g) / k
    pb = torch.randn(48, device="cuda", generator=g)
    w0p = torch.zeros(96, 32, device="cuda", dtype=dt)
    w0p[:, :27] = w0.reshape(96, 27)                                   # k = c*9 + ky*3 + kx
    pwp = pw.view(48, k * k).t().contiguous()                          # [k*k][C]
    Ho, Wo = H // 2, W // 2
    stem = torch.full((B, Ho, Wo, 48), 7.0, device="cuda", dtype=dt)
    y1 = torch.full((B, Ho, Wo, 48), 7.0, device="cuda", dtype=dt)
    # p0s lands in the right half of a 96-wide [y2 | p0s] operand when W is the model's 224 (engine layout), dense otherwise
    pitch = 96 if W == 224 else 48
    cat = torch.full((B, Ho // 2, Wo // 2, pitch), 7.0, device="cuda", dtype=dt)
    p0s = cat[..., pitch - 48:]
    pooled = torch.full((B * (Ho // k) * (Wo // k), 48), 7.0, device="cuda", dtype=dt)
    check(LIB.msclip_op_front_conv(ptr(img), code, B, H, W, ptr(w0p), ptr(b0), ptr(w1), ptr(b1), ptr(pwp), ptr(pb), k,
                                   ptr(stem), ptr(y1), C.c_void_p(cat.data_ptr() + 2 * (pitch - 48)), pitch, ptr(pooled), stream()))
    assert pitch == 48 or torch.all(cat[..., :48] == 7.0)
    a = torch.relu(F.conv2d(img.to(dt).float(), w0.float(), b0, stride=2, padding=1))    # operands rounded as the kernel does
    p0 = a[:, 48:]
    p0r = p0.to(dt).float()
    refs = {
        "stem": a[:, :48].permute(0, 2, 3, 1),
        "y1": torch.relu(F.conv2d(p0r, w1.float()[:, :, None, None], b1)).permute(0, 2, 3, 1),
        "p0s": p0r[:, :, ::2, ::2].permute(0, 2, 3, 1),
        "pooled": (F.conv2d(p0, pw, stride=k, groups=48) + pb[None, :, None, None]).permute(0, 2, 3, 1).reshape(-1, 48),
    }
    outs = {"stem": stem, "y1": y1, "p0s": p0s, "pooled": pooled}
    res = {}
    for name, ref in refs.items():
        res[name] = rel(outs[name].float().reshape(ref.shape), ref)
    _record(f"front_conv/B{B}_{H}x{W}_k{k}_{'i16' if img16 else 'f32'}", res)
    for name, r in res.items():
        assert r < tol16(), (name, r)
    # the strided copy is the rounded p0 itself: bit-equal wherever the fp32 accumulation order does not flip a rounding
    same = (p0s.float() == refs["p0s"].to(dt).float()).float().mean().item()
    assert same > 0.98, same


@pytest.mark.parametrize("g", [7, 14])
def test_adapter_tail_matches_oracle(g):
    """Lateral adapter (M.py:1752-1778) through the oracle: BN folded by hand here, exactly as the engine does."""
    B, D, L = 3, 768, g * g + 1
    x = torch.randn(B, L, D, device="cuda")
    t = torch.randn(B * g * g, D, device="cuda")
    dw = torch.randn(D, 1, 3, 3, device="cuda") / 3
    gamma, beta = torch.rand(D, device="cuda") + 0.5, torch.randn(D, device="cuda") * 0.1
    mean, var = torch.randn(D, device="cuda") * 0.1, torch.rand(D, device="cuda") + 0.5
    lw, lb = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    scale = gamma / torch.sqrt(var + 1e-5)
    w9 = (dw.view(D, 9) * scale[:, None]).t().contiguous()
    bias = beta - mean * scale
    out = torch.empty_like(x)
    check(LIB.msclip_op_adapter_fuse_ln(ptr(x), ptr(t), ptr(w9), ptr(bias), ptr(lw), ptr(lb), ptr(out), B, g, stream()))
    # oracle pieces
    cls, tok = x[:, :1], x[:, 1:]
    gmap = tok.transpose(1, 2).reshape(B, D, g, g)
    bconv = F.conv2d(gmap, dw, padding=1, groups=D)
    bconv = (bconv - mean[None, :, None, None]) / torch.sqrt(var + 1e-5)[None, :, None, None] * gamma[None, :, None, None] \
        + beta[None, :, None, None]
    y = torch.cat([cls + cls, bconv.flatten(2).transpose(1, 2) + t.view(B, g * g, D)], dim=1)
    ref = O.layer_norm(y, lw, lb)
    r = rel(out, ref)
    _record(f"adapter_tail/g{g}", {"rel": r})
    assert r < 1e-5


@pytest.mark.parametrize("b,scale", [(8, math.e), (128, 14.2857), (300, 100.0), (1000, 14.2857), (4096, 100.0)])
def test_contrastive_lse(b, scale):
    g = torch.Generator(device="cuda").manual_seed(b)
    fi = F.normalize(torch.randn(b, 512, device="cuda", generator=g), dim=-1)
    ft = F.normalize(fi * 0.5 + torch.randn(b, 512, device="cuda", generator=g) * 0.05, dim=-1)   # correlated pairs
    fi16, ft16 = fi.half().contiguous(), ft.half().contiguous()   # the exchanged embeddings are fp16 in both builds
    ws = torch.empty(LIB.msclip_op_contrastive_lse_workspace(b), device="cuda", dtype=torch.uint8)
    parts = torch.zeros(2, device="cuda")
    check(LIB.msclip_op_contrastive_lse(ptr(fi16), ptr(ft16), b, scale, ptr(ws), ptr(parts), stream()))
    logits = scale * fi16.double() @ ft16.double().t()
    d = logits.diag()
    ref0 = float((torch.logsumexp(logits, 1) - d).sum())
    ref1 = float((torch.logsumexp(logits, 0) - d).sum())
    got = parts.double().cpu()
    loss = float(got.sum()) / (2 * b)
    loss_ref = (ref0 + ref1) / (2 * b)
    loss_fp32_feats = float(O.contrastive_loss(scale * fi @ ft.t()))
    _record(f"contrastive_lse/b{b}", {"parts": got.tolist(), "ref": [ref0, ref1], "loss": loss, "loss_ref": loss_ref,
                                      "loss_fp32_features": loss_fp32_feats})
    # per row the kernel's fp32 (lse_i - s_ii) may differ from the float64 value by a few ulp of the score
    assert abs(got[0] - ref0) <= 2e-5 * abs(ref0) + 3e-6 * b
    assert abs(got[1] - ref1) <= 2e-5 * abs(ref1) + 3e-6 * b
    assert float(got.min()) >= 0.0        # lse_i >= s_ii exactly: both come from the same accumulator
    assert abs(loss - loss_fp32_feats) <= 1e-3 * abs(loss_fp32_feats) + 1e-4
