"""Backward pass on the B200, through the C ABI.

The reference ships no backward: the specification is torch.autograd on its forward.  Three layers of checks:
  * per kernel (msclip_b200_ops.h): weight-gradient GEMM, attention backward, LayerNorm backward, QuickGELU backward and
    the fused AdamW against torch fp32 autograd / torch.optim.AdamW on the same (bf16-rounded) operands;
  * per tower: msclip_backward against autograd through the CPU oracle (full tensors, every trainable key);
  * whole step: loss_and_backward against the gradients of the REAL reference (tests/golden/grad_*.npz; the oracle's own
    backward is pinned to them in tests/test_oracle_grads.py).
Tolerances (Frobenius-relative): fp32-accumulated contractions of identical 16-bit operands 1e-4; kernels that round
intermediates to bf16 (P and dS in attention, du in the MLP) 1e-2; end-to-end gradients of a bf16-operand pipeline 3e-2 per
tensor (same order as the forward's 0.3-1.7e-2 on logits), 1.5e-2 on the norm-weighted aggregate.
"""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest
import torch

from msclip_b200 import _lib, synth
from msclip_b200.config import MSCLIPConfig
from msclip_b200.model import CLIP
from oracle import msclip_oracle as O
from golden_util import GRAD_CASES, grad_sample, load_case, load_grad_case, rel_err, trainable_keys

pytestmark = pytest.mark.gpu

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
_results = {}
LIB = None


def _record(name, value):
    _results[name] = value
    try:
        os.makedirs(OUT_DIR, exist_ok=True)
        with open(os.path.join(OUT_DIR, "parity_backward.json"), "w") as f:
            json.dump(_results, f, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    global LIB
    assert torch.cuda.is_available(), "GPU tests need a B200"
    LIB = _lib.lib("bf16")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    yield


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, what=""):
    _lib.check(rc, what, "bf16")


# ------------------------------------------------------------------------------------------------ weight gradient
def run_wgrad(tokens, n, k, accumulate, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dy = torch.randn(tokens, n, device="cuda", generator=g).bfloat16()
    x = torch.randn(tokens, k, device="cuda", generator=g).bfloat16()
    dw0 = torch.randn(n, k, device="cuda", generator=g)
    dw = dw0.clone()
    ws = torch.empty(LIB.msclip_op_wgrad_workspace(tokens, n, k), dtype=torch.uint8, device="cuda")
    check(LIB.msclip_op_wgrad(ptr(dy), n, ptr(x), k, tokens, n, k, ptr(dw), int(accumulate), ptr(ws), stream()), "msclip_op_wgrad")
    torch.cuda.synchronize()
    ref = dy.float().t() @ x.float()
    if accumulate:
        ref = ref + dw0
    return rel(dw, ref)


WGRAD_CASES = [("out_proj", 1000, 768, 768), ("qkv", 616, 2304, 768), ("fc2", 4113, 768, 3072), ("fc1", 2500, 3072, 768),
               ("proj", 300, 768, 512), ("tiny", 8, 768, 512), ("long", 40000, 768, 768)]


@pytest.mark.parametrize("name,tokens,n,k", WGRAD_CASES)
def test_wgrad(name, tokens, n, k):
    e0 = run_wgrad(tokens, n, k, False)
    if e0 > 1e-4:
        # bring-up aid: which descriptor offsets of the MN-major operands would have worked?
        probe = {}
        for lbo, sbo in [(1024, 8192), (8192, 128), (128, 1024), (1024, 1024), (8192, 8192), (2048, 1024), (1024, 2048)]:
            LIB.msclip_op_set_wgrad_desc(lbo, sbo)
            try:
                probe[f"lbo{lbo}_sbo{sbo}"] = run_wgrad(tokens, n, k, False)
            except Exception as exc:      # noqa: BLE001
                probe[f"lbo{lbo}_sbo{sbo}"] = repr(exc)
        LIB.msclip_op_set_wgrad_desc(0, 0)
        _record(f"wgrad_probe/{name}", probe)
        raise AssertionError(f"wgrad {name}: rel err {e0}; descriptor probe {probe}")
    e1 = run_wgrad(tokens, n, k, True, seed=1)
    _record(f"wgrad/{name}", {"rel": e0, "rel_accumulate": e1})
    assert e1 < 1e-4


# ------------------------------------------------------------------------------------------------ attention backward
ATT_CASES = [(3, 50, False), (2, 77, True), (5, 64, False), (3, 80, True), (1, 7, True), (2, 33, False), (64, 77, True),
             (2, 197, False), (1, 197, True), (3, 130, False), (2, 96, True), (1, 208, False), (1, 81, True)]   # > 80: three-pass kernel


@pytest.mark.parametrize("batch,L,causal", ATT_CASES)
def test_attention_bwd(batch, L, causal):
    heads, hd = 12, 64
    W = heads * hd
    g = torch.Generator(device="cuda").manual_seed(batch * 100 + L)
    qkv = torch.randn(batch * L, 3 * W, device="cuda", generator=g)
    qkv[:, :W] *= 0.125 * 1.5          # the packed projection delivers q / 8
    qkv = qkv.bfloat16()
    dctx = (0.1 * torch.randn(batch * L, W, device="cuda", generator=g)).bfloat16()
    dqkv = torch.full((batch * L, 3 * W), 7.0, device="cuda").bfloat16()
    check(LIB.msclip_op_attention_bwd(ptr(qkv), ptr(dctx), ptr(dqkv), batch, L, heads, int(causal), stream()), "attention_bwd")
    torch.cuda.synchronize()
    leaf = qkv.float().requires_grad_(True)
    q, k, v = [t.view(batch, L, heads, hd).transpose(1, 2) for t in leaf.split(W, dim=-1)]
    s = q @ k.transpose(-1, -2)
    if causal:
        s = s + torch.full((L, L), float("-inf"), device="cuda").triu_(1)
    ctx = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(batch * L, W)
    (ctx * dctx.float()).sum().backward()
    ref = leaf.grad.clone()
    ref[:, :W] *= 0.125                # ours is the gradient of the UNSCALED q
    got = dqkv.float()
    errs = {n: rel(got[:, i * W:(i + 1) * W], ref[:, i * W:(i + 1) * W]) for i, n in enumerate("qkv")}
    _record(f"attention_bwd/b{batch}_l{L}_{'causal' if causal else 'full'}", errs)
    assert max(errs.values()) < 1e-2, errs


# ------------------------------------------------------------------------------------------------ LayerNorm / QuickGELU
@pytest.mark.parametrize("rows,accumulate", [(1, 1), (77, 0), (1000, 1), (20000, 1)])
def test_layernorm_bwd(rows, accumulate):
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = (2.0 * torch.randn(rows, 768, device="cuda", generator=g) + 0.5).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(768, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(768, device="cuda", generator=g)).requires_grad_(True)
    dy = torch.randn(rows, 768, device="cuda", generator=g)
    dx0 = torch.randn(rows, 768, device="cuda", generator=g)
    O.layer_norm(x, gamma, beta).backward(dy)
    ref_dx = x.grad + (dx0 if accumulate else 0)
    dx = dx0.clone()
    dx16 = torch.empty(rows, 768, device="cuda", dtype=torch.bfloat16)
    dgam, dbet, dcs = [torch.full((768,), 0.25, device="cuda") for _ in range(3)]
    ws = torch.empty(LIB.msclip_op_bwd_workspace(rows), dtype=torch.uint8, device="cuda")
    check(LIB.msclip_op_layernorm_bwd(ptr(x.detach()), ptr(dy), ptr(gamma.detach()), ptr(dx), ptr(dx16), ptr(dgam), ptr(dbet),
                                      ptr(dcs), rows, accumulate, ptr(ws), stream()), "layernorm_bwd")
    torch.cuda.synchronize()
    errs = {"dx": rel(dx, ref_dx), "dx16": rel(dx16.float(), ref_dx), "dgamma": rel(dgam - 0.25, gamma.grad),
            "dbeta": rel(dbet - 0.25, beta.grad), "colsum": rel(dcs - 0.25, ref_dx.sum(0))}
    _record(f"layernorm_bwd/{rows}_{accumulate}", errs)
    assert errs["dx"] < 2e-5 and errs["dx16"] < 3e-3 and errs["dgamma"] < 1e-4 and errs["dbeta"] < 1e-4 and errs["colsum"] < 1e-4, errs


@pytest.mark.parametrize("rows,width", [(3, 3072), (1000, 3072), (5000, 768)])
def test_qgelu_bwd(rows, width):
    g = torch.Generator(device="cuda").manual_seed(rows + width)
    u = (2.0 * torch.randn(rows, width, device="cuda", generator=g)).bfloat16()
    da = torch.randn(rows, width, device="cuda", generator=g).bfloat16()
    uf = u.float().requires_grad_(True)
    O.quick_gelu(uf).backward(da.float())
    ref = uf.grad
    db = torch.zeros(width, device="cuda")
    ws = torch.empty(LIB.msclip_op_bwd_workspace(rows), dtype=torch.uint8, device="cuda")
    out = da.clone()
    check(LIB.msclip_op_qgelu_bwd(ptr(out), ptr(u), ptr(db), rows, width, ptr(ws), stream()), "qgelu_bwd")
    torch.cuda.synchronize()
    errs = {"du": rel(out.float(), ref), "dbias": rel(db, ref.sum(0))}      # the bias gradient sums the UNROUNDED values
    _record(f"qgelu_bwd/{rows}x{width}", errs)
    assert errs["du"] < 3e-3 and errs["dbias"] < 1e-4, errs


def test_adamw_matches_torch():
    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(768, 768), (3072,), (1,), (65537,), (77, 768)]
    ours = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    ref = [torch.nn.Parameter(t.clone()) for t in ours]
    lrs = [1e-3, 1e-3, 5e-3, 2e-3, 1e-3]
    wds = [0.05, 0.0, 0.0, 0.2, 0.05]
    opt = torch.optim.AdamW([{"params": [p], "lr": lr, "weight_decay": wd} for p, lr, wd in zip(ref, lrs, wds)],
                            betas=(0.9, 0.98), eps=1e-6)
    m = [torch.zeros_like(t) for t in ours]
    v = [torch.zeros_like(t) for t in ours]
    n = len(ours)
    P = C.c_void_p
    for step in range(1, 4):
        grads = [torch.randn(s, device="cuda", generator=g) for s in shapes]
        for p, gr in zip(ref, grads):
            p.grad = gr.clone()
        opt.step()
        check(LIB.msclip_op_adamw(n, (P * n)(*[t.data_ptr() for t in ours]), (P * n)(*[t.data_ptr() for t in grads]),
                                  (P * n)(*[t.data_ptr() for t in m]), (P * n)(*[t.data_ptr() for t in v]),
                                  (C.c_int64 * n)(*[t.numel() for t in ours]), (C.c_float * n)(*lrs), (C.c_float * n)(*wds),
                                  0.9, 0.98, 1e-6, step, stream()), "adamw")
        torch.cuda.synchronize()
    worst = max(rel(a, b.detach()) for a, b in zip(ours, ref))
    _record("adamw", worst)
    assert worst < 1e-6


# ------------------------------------------------------------------------------------------------ towers vs oracle autograd
def build_train_model(cfg, sd_np):
    model = CLIP(cfg, precision="bf16")
    model.load_state_dict({k: torch.as_tensor(v) for k, v in sd_np.items()}, strict=True)
    model = model.cuda().eval()
    model.enable_training()
    return model


def oracle_leaves(sd_np):
    sd = O.to_torch(sd_np)
    seen = set()
    for t in sd.values():
        if t.dtype == torch.float32 and id(t) not in seen:
            t.requires_grad_(True)
            seen.add(id(t))
    return sd


def compare_grads(model, ref, keys, tag, tol=3e-2, agg_tol=1.5e-2):
    params = model.trainable_parameters()
    errs, num, den = {}, 0.0, 0.0
    for k in keys:
        r = ref[k]
        if r is None:
            continue
        got = params[k].grad.detach().cpu().double().reshape(r.shape)
        r = r.double()
        d = float((got - r).norm())
        errs[k] = d / max(float(r.norm()), 1e-30)
        num += d * d
        den += float(r.norm()) ** 2
    agg = math.sqrt(num / max(den, 1e-300))
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    _record(tag, {"aggregate": agg, "worst": worst, "n": len(errs)})
    bad = {k: e for k, e in errs.items() if e > tol}
    assert not bad, (tag, worst)
    assert agg < agg_tol, (tag, agg)
    return errs


@pytest.mark.parametrize("layers,batch,ragged", [(2, 6, False), (3, 5, True)])
def test_text_tower_backward_matches_oracle_autograd(layers, batch, ragged):
    cfg = MSCLIPConfig(patch_size=32, layers=layers)
    sd_np = synth.synth_state_dict(cfg, seed=5)
    tok = synth.synth_tokens(batch, 21, cfg.context_length, cfg.vocab_size, ragged=ragged)
    model = build_train_model(cfg, sd_np)
    g = torch.Generator().manual_seed(9)
    d_feat = torch.randn(batch, cfg.embed_dim, generator=g)
    model.zero_grad()
    model.encode_text(torch.from_numpy(tok).cuda())
    model.backward_features(None, d_feat.cuda())
    torch.cuda.synchronize()
    sd = oracle_leaves(sd_np)
    O.encode_text(torch.from_numpy(tok), sd, cfg).backward(d_feat)
    keys = [k for k in trainable_keys(cfg) if not k.startswith("visual.") or ".resblocks." in k]
    keys = [k for k in keys if k != "logit_scale" and sd[k].grad is not None]
    compare_grads(model, {k: sd[k].grad for k in keys}, keys, f"text_tower/l{layers}_b{batch}")


@pytest.mark.parametrize("layers,batch,patch", [(2, 3, 32), (3, 4, 32), (5, 2, 32), (3, 2, 16)])
def test_image_tower_backward_matches_oracle_autograd(layers, batch, patch):
    cfg = MSCLIPConfig(patch_size=patch, layers=layers)
    sd_np = synth.synth_state_dict(cfg, seed=6)
    img = synth.synth_images(batch, 22, cfg.image_resolution)
    model = build_train_model(cfg, sd_np)
    g = torch.Generator().manual_seed(10)
    d_feat = torch.randn(batch, cfg.embed_dim, generator=g)
    model.zero_grad()
    model.encode_image(torch.from_numpy(img).cuda())
    model.backward_features(d_feat.cuda(), None)
    torch.cuda.synchronize()
    sd = oracle_leaves(sd_np)
    O.encode_image(torch.from_numpy(img), sd, cfg).backward(d_feat)
    keys = [k for k in trainable_keys(cfg) if k.startswith("visual.")]
    compare_grads(model, {k: sd[k].grad for k in keys}, keys, f"image_tower/p{patch}_l{layers}_b{batch}")


# ------------------------------------------------------------------------------------------------ whole step vs the reference
@pytest.mark.parametrize("name", GRAD_CASES)
def test_training_step_matches_reference_gradients(name):
    cfg, sd_np, img, tok, z, meta = load_case(name)
    gz = load_grad_case(name)
    model = build_train_model(cfg, sd_np)
    model.zero_grad()
    loss = float(model.loss_and_backward(torch.from_numpy(img).cuda(), torch.from_numpy(tok).cuda()))
    torch.cuda.synchronize()
    ref_loss = json.loads(str(gz["meta"]))["loss"]
    assert abs(loss - ref_loss) <= 1e-3 * abs(ref_loss), (loss, ref_loss)
    params = model.trainable_parameters()
    errs, ac_errs, norm_errs, num, den = {}, {}, {}, 0.0, 0.0
    for key in trainable_keys(cfg):
        got = params[key].grad.detach().float().cpu().numpy()
        ref_s = gz["sample/" + key].astype(np.float64)
        got_s = grad_sample(got, key, tok).astype(np.float64)
        ref_norm = float(gz["norm/" + key])
        d = float(np.linalg.norm(got_s - ref_s))
        errs[key] = d / max(float(np.linalg.norm(ref_s)), 1e-30)
        ac_errs[key] = rel_err(gz["sample_autocast/" + key], ref_s)     # the reference's own autocast(bf16) backward
        norm_errs[key] = abs(float(np.linalg.norm(got.astype(np.float64))) - ref_norm) / max(ref_norm, 1e-30)
        num += d * d
        den += float(np.linalg.norm(ref_s)) ** 2
    agg = math.sqrt(num / max(den, 1e-300))
    agg_ac = json.loads(str(gz["meta"]))["autocast_aggregate"]
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    _record(f"training_step/{name}", {"loss": loss, "loss_reference": ref_loss, "aggregate": agg,
                                      "reference_autocast_aggregate": agg_ac, "worst": [(k, e, ac_errs[k]) for k, e in worst],
                                      "worst_norm": sorted(norm_errs.items(), key=lambda kv: -kv[1])[:3]})
    # The yardstick of the forward tests applied to the backward: gradients whose terms cancel over the
    # batch (biases, at a near-uniform softmax) cannot be better than their 16-bit operands allow; we hold ourselves to
    # 0.75 x the error of the reference's own autocast(bf16) backward in aggregate, and per tensor to 3e-2 or that tensor's
    # autocast error, whichever is larger.
    assert agg < 0.75 * agg_ac, (agg, agg_ac, worst)
    bad = {k: (e, ac_errs[k]) for k, e in errs.items() if e > max(3e-2, 1.25 * ac_errs[k])}
    assert not bad, bad
    bad_norm = {k: (e, ac_errs[k]) for k, e in norm_errs.items() if e > max(3e-2, ac_errs[k])}
    assert not bad_norm, bad_norm


def test_adamw_step_trains_and_refreshes_packed_weights():
    """Three fused AdamW steps on a 2-layer model lower the loss, and the packed weights follow the fp32 masters: the
    loss after the steps equals the loss of a FRESH handle loaded with the updated state dict."""
    from msclip_b200.optim import AdamW
    cfg = MSCLIPConfig(patch_size=32, layers=2)
    sd_np = synth.synth_state_dict(cfg, seed=11, logit_scale=math.log(20.0))
    img, tok = synth.correlated_pair_batch(cfg, 8, seed=3)
    timg, ttok = torch.from_numpy(img).cuda(), torch.from_numpy(tok).cuda()
    model = build_train_model(cfg, sd_np)
    opt = AdamW(model, lr=2e-4, weight_decay=0.05, lr_share=1e-4, wd_share=0.2)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        losses.append(float(model.loss_and_backward(timg, ttok)))
        opt.step()
    final = float(model.contrastive_loss(timg, ttok))
    fresh = CLIP(cfg, precision="bf16")
    fresh.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)
    fresh = fresh.cuda().eval()
    ref_final = float(fresh.contrastive_loss(timg, ttok))
    _record("adamw_training", {"losses": losses, "final": final, "fresh_handle": ref_final})
    assert final < losses[0], (losses, final)
    assert abs(final - ref_final) <= 1e-5 * abs(ref_final) + 1e-7, (final, ref_final)


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """Training-format checkpoint of the reference (lib/utils/utils.py:157-199, with the DDP `module.` prefix): save after two
    steps, load into a fresh model + optimiser, and the third step must produce bit-identical weights."""
    from msclip_b200.checkpoint import load_checkpoint, save_checkpoint
    from msclip_b200.optim import AdamW
    cfg = MSCLIPConfig(patch_size=32, layers=2)
    sd_np = synth.synth_state_dict(cfg, seed=12, logit_scale=math.log(10.0))
    img, tok = synth.correlated_pair_batch(cfg, 6, seed=4)
    timg, ttok = torch.from_numpy(img).cuda(), torch.from_numpy(tok).cuda()

    def one_step(model, opt):
        opt.zero_grad()
        loss = float(model.loss_and_backward(timg, ttok))
        opt.step()
        return loss

    a = build_train_model(cfg, sd_np)
    oa = AdamW(a, lr=3e-4, weight_decay=0.05, lr_share=1e-4, wd_share=0.2)
    for _ in range(2):
        one_step(a, oa)
    path = str(tmp_path / "checkpoint.pth")
    saved = save_checkpoint(a, path, oa, epoch_or_step=1, in_epoch=False, distributed=True)
    assert saved["step"] == 2 and all(k.startswith("module.") for k in saved["state_dict"])
    b = build_train_model(cfg, synth.synth_state_dict(cfg, seed=99))
    ob = AdamW(b, lr=1.0)
    meta = load_checkpoint(b, path, ob)
    assert meta["step"] == 2 and meta["model"] == "clip_openai_pe_res_v1"
    la, lb = one_step(a, oa), one_step(b, ob)
    assert la == lb, (la, lb)
    for (k, pa), (_k, pb) in zip(a.state_dict().items(), b.state_dict().items()):
        if pa.dtype == torch.float32 and "token_embedding" not in k:      # the embedding scatter uses fp32 atomics
            assert torch.equal(pa, pb), k


def test_training_trajectory_tracks_torch_autograd_plus_adamw():
    """End-to-end training dynamics: 12 optimiser steps of the drop-in (our forward, loss, backward, fused AdamW, re-packed
    weights) against the same 12 steps done with torch - autograd through the CPU-oracle forward (moved to the GPU, fp32) and
    torch.optim.AdamW with the same parameter groups, only our trainable keys receiving updates (the convolutional front is
    frozen on both sides).  The per-step losses must track each other; both must fall."""
    from msclip_b200.optim import AdamW, _is_shared, _no_decay
    cfg = MSCLIPConfig(patch_size=32, layers=3)
    sd_np = synth.synth_state_dict(cfg, seed=21, logit_scale=math.log(30.0))
    img, tok = synth.correlated_pair_batch(cfg, 16, seed=5)
    timg, ttok = torch.from_numpy(img).cuda(), torch.from_numpy(tok).cuda()
    steps, lr, wd, lr_s, wd_s = 12, 2e-4, 0.05, 1e-4, 0.2
    model = build_train_model(cfg, sd_np)
    opt = AdamW(model, lr=lr, weight_decay=wd, lr_share=lr_s, wd_share=wd_s)
    ours = []
    for _ in range(steps):
        opt.zero_grad()
        ours.append(float(model.loss_and_backward(timg, ttok)))
        opt.step()
    # torch side
    sd = O.to_torch(sd_np, device="cuda")
    keys = [k for k in trainable_keys(cfg)]
    groups, seen = [], set()
    for k in keys:
        t = sd[k]
        if id(t) in seen:
            continue
        seen.add(id(t))
        t.requires_grad_(True)
        shared = _is_shared(k)
        groups.append({"params": [t], "lr": lr_s if shared else lr,
                       "weight_decay": 0.0 if _no_decay(k) else (wd_s if shared else wd)})
    ref_opt = torch.optim.AdamW(groups, betas=(0.9, 0.999), eps=1e-8)
    ref = []
    for _ in range(steps):
        ref_opt.zero_grad()
        loss = O.contrastive_loss(O.forward(timg, ttok, sd, cfg))
        loss.backward()
        ref_opt.step()
        ref.append(float(loss.detach()))
    _record("training_trajectory", {"ours": ours, "torch_fp32": ref})
    assert ours[-1] < ours[0] and ref[-1] < ref[0], (ours, ref)
    for a, b in zip(ours, ref):
        assert abs(a - b) <= 2e-2 * abs(b) + 2e-3, (ours, ref)


def test_backward_error_paths():
    """Loud failures instead of silent wrong gradients: backward without training mode, without a taped forward, and for a
    batch larger than one tape holds."""
    cfg = MSCLIPConfig(patch_size=32, layers=2)
    sd_np = synth.synth_state_dict(cfg, seed=3)
    plain = CLIP(cfg, precision="bf16")
    plain.load_state_dict({k: torch.as_tensor(v) for k, v in sd_np.items()}, strict=True)
    plain = plain.cuda().eval()
    tok = torch.from_numpy(synth.synth_tokens(4, 1, cfg.context_length, cfg.vocab_size)).cuda()
    plain.encode_text(tok)
    with pytest.raises(_lib.MsclipError, match="training is not enabled"):
        plain.backward_features(None, torch.zeros(4, cfg.embed_dim, device="cuda"))
    with pytest.raises(_lib.MsclipError, match="enable_training"):
        plain.trainable_parameters()
    model = build_train_model(cfg, sd_np)
    with pytest.raises(_lib.MsclipError, match="no taped encode_text"):
        model.backward_features(None, torch.zeros(4, cfg.embed_dim, device="cuda"))
    model.encode_text(tok)
    with pytest.raises(_lib.MsclipError, match="no taped encode_image"):
        model.backward_features(torch.zeros(4, cfg.embed_dim, device="cuda"), None)
    fp16 = CLIP(cfg, precision="fp16")
    fp16.load_state_dict({k: torch.as_tensor(v) for k, v in sd_np.items()}, strict=True)
    fp16 = fp16.cuda().eval()
    with pytest.raises(_lib.MsclipError, match="bf16 build"):
        fp16.enable_training()


def test_micro_batched_training_step_equals_one_shot():
    """GradCache-style step (micro-batches encoded twice, one loss over all of them) against the one-shot step on the same
    16 pairs: same loss, same gradients up to summation order (the towers see 3 x 6-, 6-, 4-row batches instead of 16 rows)."""
    cfg = MSCLIPConfig(patch_size=32, layers=3)
    sd_np = synth.synth_state_dict(cfg, seed=31, logit_scale=math.log(25.0))
    img, tok = synth.correlated_pair_batch(cfg, 16, seed=6)
    timg, ttok = torch.from_numpy(img).cuda(), torch.from_numpy(tok).cuda()
    one = build_train_model(cfg, sd_np)
    one.zero_grad()
    loss_one = float(one.loss_and_backward(timg, ttok))
    mb = build_train_model(cfg, sd_np)
    mb.setup_data_parallel(16)
    mb.zero_grad()
    loss_mb = float(mb.loss_and_backward(timg, ttok, micro_batch=6))
    torch.cuda.synchronize()
    assert abs(loss_one - loss_mb) <= 2e-6 * abs(loss_one), (loss_one, loss_mb)
    ref, got = one.trainable_parameters(), mb.trainable_parameters()
    num = den = 0.0
    worst = ("", 0.0)
    for k, p in ref.items():
        d = float((got[k].grad - p.grad).double().norm())
        r = float(p.grad.double().norm())
        num += d * d
        den += r * r
        if r > 0 and d / r > worst[1]:
            worst = (k, d / r)
    agg = math.sqrt(num / max(den, 1e-300))
    _record("micro_batched_training", {"loss_one_shot": loss_one, "loss_micro": loss_mb, "aggregate": agg, "worst": list(worst)})
    assert agg < 2e-3 and worst[1] < 2e-2, (agg, worst)
