"""Model-level parity on the B200: the drop-in CLIP (msclip_b200.model, C ABI underneath) against

  (1) the golden vectors produced by the REAL reference (tests/golden/*.npz, oracle/make_golden.py),
  (2) the CPU oracle on fresh seeded inputs.

Three-way protocol of SURVEY.md 7.2(1): ours-vs-fp32 reference, the reference's own autocast(bf16)
forward-vs-fp32 reference (stored in the golden files), and the bound we hold ourselves to:
  * features, Frobenius-relative:  <= 6e-3 and strictly better than the reference's own bf16 mode
  * logits, Frobenius-relative:    <= 1.0 x the reference's own bf16 mode (same bf16-operand arithmetic: the
                                   error of near-orthogonal dot products is dominated by operand rounding)
                                   and max |error| <= 1.5e-3 x logit scale (i.e. 1.5e-3 in cosine units)
  * loss, relative:                <= 1e-3   (north star), every case, fused kernel and logits path alike
The same model built on the fp16-operand library (same kernels, -DMSCLIP_FP16) is held to the north star's
literal bar: features <= 1e-3, loss <= 1e-3 (<= 2e-4 observed), max |logit error| <= 3e-4 x scale.
The residual stream, LayerNorm, softmax and every accumulator are fp32 in both builds.
"""
import json
import math
import os

import numpy as np
import pytest
import torch

from msclip_b200 import synth
from msclip_b200.config import MSCLIPConfig
from msclip_b200.model import CLIP, get_clip_model
from oracle import msclip_oracle as O
from golden_util import CASES, load_case, rel_err

pytestmark = pytest.mark.gpu

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
_results = {}
FEAT_TOL, COS_TOL, LOSS_TOL = 6e-3, 1.5e-3, 1e-3


def _record(name, value):
    _results[name] = value
    try:
        os.makedirs(OUT_DIR, exist_ok=True)
        with open(os.path.join(OUT_DIR, "parity_model.json"), "w") as f:
            json.dump(_results, f, indent=1, sort_keys=True)
    except OSError:
        pass


def build_model(cfg, sd_np, precision="bf16"):
    model = CLIP(cfg, precision=precision)
    sd = {k: torch.as_tensor(v) for k, v in sd_np.items()}
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.cuda().eval()


def loss_of(logits):
    return float(O.contrastive_loss(torch.as_tensor(logits).double()))


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("name", CASES)
def test_matches_reference_golden(name, precision):
    cfg, sd_np, img, tok, z, meta = load_case(name)
    model = build_model(cfg, sd_np, precision)
    timg, ttok = torch.from_numpy(img).cuda(), torch.from_numpy(tok).cuda()
    fi = model.encode_image(timg).cpu().numpy()
    ft = model.encode_text(ttok).cpu().numpy()
    fi_raw = model.encode_image(timg, norm=False).cpu().numpy()
    logits = model(timg, ttok).cpu().numpy()
    loss_fused = float(model.contrastive_loss(timg, ttok))
    assert np.isfinite(fi).all() and np.isfinite(ft).all() and np.isfinite(logits).all()
    ref_loss = float(z["loss"])
    res = {
        "ours_vs_fp32": {
            "image_features": rel_err(fi, z["image_features"]), "text_features": rel_err(ft, z["text_features"]),
            "image_features_unnormalised": rel_err(fi_raw, z["image_features_unnormalised"]),
            "logits": rel_err(logits, z["logits"]), "logits_max_abs": float(np.abs(logits - z["logits"]).max()),
            "loss_from_logits": abs(loss_of(logits) - ref_loss) / abs(ref_loss),
            "loss_fused_kernel": abs(loss_fused - ref_loss) / abs(ref_loss),
        },
        "loss": {"reference_fp32": ref_loss, "ours_fused": loss_fused, "ours_from_logits": loss_of(logits)},
    }
    if "logits_autocast" in z.files:
        res["reference_autocast_vs_fp32"] = {
            "image_features": rel_err(z["image_features_autocast"], z["image_features"]),
            "text_features": rel_err(z["text_features_autocast"], z["text_features"]),
            "logits": rel_err(z["logits_autocast"], z["logits"]),
            "loss": abs(float(z["loss_autocast"]) - ref_loss) / abs(ref_loss),
        }
        res["ours_vs_reference_autocast"] = {"logits": rel_err(logits, z["logits_autocast"])}
    _record(f"{precision}/{name}", res)
    o = res["ours_vs_fp32"]
    a = res["reference_autocast_vs_fp32"]
    scale = math.exp(meta["logit_scale"])
    if precision == "fp16":
        assert max(o["image_features"], o["text_features"], o["image_features_unnormalised"]) < 1e-3, o
        assert o["logits_max_abs"] <= 3e-4 * scale, o
        assert o["loss_from_logits"] < LOSS_TOL and o["loss_fused_kernel"] < LOSS_TOL, o
        assert o["logits"] <= 0.3 * a["logits"], (o, a)
        return
    assert o["image_features"] < FEAT_TOL and o["text_features"] < FEAT_TOL and o["image_features_unnormalised"] < FEAT_TOL, o
    assert o["logits_max_abs"] <= COS_TOL * scale, o
    assert o["loss_from_logits"] < LOSS_TOL and o["loss_fused_kernel"] < LOSS_TOL, o      # north star: <= 1e-3 relative
    assert o["image_features"] < a["image_features"] and o["text_features"] < a["text_features"], (o, a)
    assert o["logits"] <= 1.0 * a["logits"], (o, a)           # no worse than the reference's own bf16


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_fresh_seed_against_cpu_oracle(precision):
    """Not in the golden set: ragged tokens, T = 100, both towers, vs the CPU oracle (fp32)."""
    cfg = MSCLIPConfig(layers=4)
    sd_np = synth.synth_state_dict(cfg, seed=41, logit_scale=math.log(100.0))
    img, tok = synth.synth_images(5, 77), synth.synth_tokens(5, 77, ragged=True)
    model = build_model(cfg, sd_np, precision)
    sd = O.to_torch(sd_np)
    with torch.no_grad():
        ref_i = O.encode_image(torch.from_numpy(img), sd, cfg)
        ref_t = O.encode_text(torch.from_numpy(tok), sd, cfg)
        ref_logits = O.similarity_logits(ref_i, ref_t, sd["logit_scale"])
    got = model(torch.from_numpy(img).cuda(), torch.from_numpy(tok).cuda()).cpu()
    r = rel_err(got.numpy(), ref_logits.numpy())
    mx = float((got - ref_logits).abs().max())
    ref_loss, got_loss = float(O.contrastive_loss(ref_logits)), float(O.contrastive_loss(got.double()))
    fused = float(model.contrastive_loss(torch.from_numpy(img).cuda(), torch.from_numpy(tok).cuda()))
    e_log, e_fused = abs(got_loss - ref_loss) / abs(ref_loss), abs(fused - ref_loss) / abs(ref_loss)
    _record(f"fresh_seed_l4/{precision}", {"logits": r, "logits_max_abs": mx, "loss_from_logits": e_log, "loss_fused_kernel": e_fused})
    # 5 x 5 near-orthogonal pairs at T = 100 (|logit| ~ 2): the worst case for operand rounding.  bf16: max error in
    # cosine units, Frobenius-relative <= 3e-2 (2.3e-2 observed), loss within 5e-3 (a 5-row loss at T = 100 amplifies
    # single logit errors; the north star's 1e-3 is held on the golden cases and, here, by the fp16-operand build)
    if precision == "fp16":
        assert mx <= 3e-4 * 100.0 and r <= 5e-3, (mx, r)
        assert e_log <= LOSS_TOL and e_fused <= LOSS_TOL, (e_log, e_fused)
    else:
        assert mx <= COS_TOL * 100.0 and r <= 3e-2, (mx, r)
        assert e_log <= 5e-3 and e_fused <= 5e-3, (e_log, e_fused)


def test_host_buffers_equal_device_buffers():
    """The C ABI accepts host pointers (pageable and pinned): results are bit-identical to device input."""
    cfg, sd_np, img, tok, z, meta = load_case("b32_l2_b8")
    model = build_model(cfg, sd_np)
    timg, ttok = torch.from_numpy(img), torch.from_numpy(tok)
    dev_i = model.encode_image(timg.cuda()).cpu()
    dev_t = model.encode_text(ttok.cuda()).cpu()
    assert torch.equal(model.encode_image(timg), dev_i)
    assert torch.equal(model.encode_image(timg.pin_memory()), dev_i)
    assert torch.equal(model.encode_text(ttok), dev_t)
    l_dev = float(model.contrastive_loss(timg.cuda(), ttok.cuda()))
    l_host = float(model.contrastive_loss(timg.pin_memory(), ttok.pin_memory()))
    assert l_dev == l_host


def test_determinism_and_batch_independence():
    """Same input twice -> same bits; a sample's embedding does not depend on its batch neighbours."""
    cfg, sd_np, img, tok, z, meta = load_case("b32_l3_b4")
    model = build_model(cfg, sd_np)
    timg, ttok = torch.from_numpy(img).cuda(), torch.from_numpy(tok).cuda()
    a, b = model.encode_image(timg), model.encode_image(timg)
    assert torch.equal(a, b)
    assert torch.equal(model.encode_image(timg[1:3]), a[1:3])
    t = model.encode_text(ttok)
    assert torch.equal(model.encode_text(ttok[2:]), t[2:])


def test_text_ignores_tokens_after_eot():
    cfg, sd_np, img, tok, z, meta = load_case("b32_l3_b4")     # ragged tokens
    model = build_model(cfg, sd_np)
    tok2 = tok.copy()
    for i in range(tok2.shape[0]):
        e = int(tok2[i].argmax())
        tok2[i, e + 1:] = (np.arange(76 - e) * 7 + 3) % 1000
    a = model.encode_text(torch.from_numpy(tok).cuda())
    b = model.encode_text(torch.from_numpy(tok2).cuda())
    assert torch.equal(a, b)


def test_load_state_dict_refreshes_packed_weights():
    cfg = MSCLIPConfig(layers=2)
    model = build_model(cfg, synth.synth_state_dict(cfg, seed=1))
    tok = torch.from_numpy(synth.synth_tokens(2, 3)).cuda()
    a = model.encode_text(tok)
    model.load_state_dict({k: torch.as_tensor(v) for k, v in synth.synth_state_dict(cfg, seed=2).items()})
    b = model.encode_text(tok)
    assert not torch.equal(a, b)
    model.load_state_dict({k: torch.as_tensor(v) for k, v in synth.synth_state_dict(cfg, seed=1).items()})
    assert torch.equal(model.encode_text(tok), a)


def test_error_behaviour():
    cfg = MSCLIPConfig(layers=2)
    model = build_model(cfg, synth.synth_state_dict(cfg, seed=1))
    with pytest.raises(AssertionError):
        model.encode_text(torch.zeros(1, 77, dtype=torch.long).cuda(), action="x")      # M.py:942
    with pytest.raises(ValueError):
        model.encode_image(torch.zeros(1, 3, 32, 32).cuda())
    bad = torch.full((1, 77), 60000, dtype=torch.long)
    with pytest.raises(RuntimeError):
        model.encode_text(bad)                      # host output path reports out-of-range ids (nn.Embedding raises)
    assert model.encode_image(torch.zeros(0, 3, 224, 224).cuda()).shape == (0, 512)


def test_zero_shot_path_config5_small():
    """tools/zero_shot.py:122-134, 265-266 on a reduced problem: class-mean text embeddings, renormalised,
    100 * image @ W, compared with the oracle (top-1 must agree wherever the oracle's margin is clear)."""
    cfg = MSCLIPConfig(layers=3)
    sd_np = synth.synth_state_dict(cfg, seed=9)
    model = build_model(cfg, sd_np)
    sd = O.to_torch(sd_np)
    n_cls, n_tpl, n_img = 6, 4, 16
    toks = np.stack([synth.synth_tokens(n_tpl, 100 + c, ragged=True) for c in range(n_cls)])
    img = synth.synth_images(n_img, 5)
    with torch.no_grad():
        w_ref = O.zeroshot_classifier(torch.from_numpy(toks), sd, cfg)
        l_ref = O.zeroshot_logits(torch.from_numpy(img), w_ref, sd, cfg)
    ws = []
    for c in range(n_cls):
        e = model.encode_text(torch.from_numpy(toks[c]).cuda()).mean(dim=0)
        ws.append(e / e.norm())
    w = torch.stack(ws, dim=0)                                        # [n_cls, 512]
    logits = model.similarity_logits(model.encode_image(torch.from_numpy(img).cuda()), w, 100.0).cpu()
    r = rel_err(logits.numpy(), l_ref.numpy())
    mx = float((logits - l_ref).abs().max())
    top2 = l_ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 0.05
    agree = (logits.argmax(1) == l_ref.argmax(1))[clear]
    _record("zero_shot_small", {"logits": r, "logits_max_abs": mx, "clear": int(clear.sum()), "agree": int(agree.sum())})
    assert mx <= COS_TOL * 100.0 and bool(agree.all())


def test_fused_zeroshot_classifier_and_predict_match_the_tool_formulas():
    """msclip_zeroshot_classifier (all prompts in one call, sorted by live length, per-class mean +
    renormalisation in one kernel) against the tool's per-class loop (tools/zero_shot.py:125-131) built from
    encode_text, and msclip_zeroshot_predict (logits + top-k on the device) against `100 * feats @ W` + torch.topk
    (tools/zero_shot.py:266, 150-163)."""
    cfg = MSCLIPConfig(layers=3)
    sd_np = synth.synth_state_dict(cfg, seed=9)
    model = build_model(cfg, sd_np)
    n_cls, n_tpl, n_img = 37, 5, 19                     # 185 prompts with ragged lengths, odd sizes everywhere
    toks = torch.from_numpy(np.stack([synth.synth_tokens(n_tpl, 300 + c, ragged=True) for c in range(n_cls)])).cuda()
    ws = []
    for c in range(n_cls):                              # the tool's loop
        e = model.encode_text(toks[c]).mean(dim=0)
        e /= e.norm()
        ws.append(e)
    w_loop = torch.stack(ws, dim=1)                     # [512, n_cls], tools/zero_shot.py:132
    w_fused = model.zeroshot_classifier(toks)
    assert w_fused.shape == w_loop.shape == (512, n_cls)
    # identical embeddings (the towers are row-independent, trimming is exact); only the fp32 order of the mean differs
    assert float((w_fused - w_loop).abs().max()) <= 2e-7
    w_host = model.zeroshot_classifier(toks.cpu())      # host tokens: same result
    assert torch.equal(w_host, w_fused)
    feats = model.encode_image(torch.from_numpy(synth.synth_images(n_img, 5)).cuda())
    idx, logits = model.zeroshot_predict(feats, w_fused, topk=5, return_logits=True)
    ref = 100.0 * feats.double() @ w_fused.double()
    assert float((logits.double() - ref).abs().max()) <= 2e-4           # split-operand tensor-core product: fp32 grade
    tv, ti = logits.topk(5, dim=1)
    assert torch.equal(idx.long(), ti) or torch.equal(torch.gather(logits, 1, idx.long()), tv)
    assert torch.equal(model.zeroshot_predict(feats, w_fused, topk=1)[:, 0].long(), logits.argmax(dim=1))
    with pytest.raises(Exception):
        model.zeroshot_predict(feats, w_fused, topk=9)
    _record("zero_shot_fused", {"classes": n_cls, "templates": n_tpl, "max_abs_w_diff": float((w_fused - w_loop).abs().max())})


def test_get_clip_model_accepts_reference_config():
    ns = lambda **k: type("N", (), k)()
    cu = dict(CUSTOM_ATTN=True, SHARE_MODULES=["attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj", "mlp"],
              PARALLEL_IN_V=True, PARALLEL_N_LAYERS=5, PARALLEL_LATERAL_LAYER=[2, 4, 6, 8, 10],
              PRALLEL_T2B_KERNELS=[16, 8, 4, 2, 1], PRALLEL_T2B_PADDINGS=[0] * 5, PRALLEL_T2B_STRIDES=[16, 8, 4, 2, 1],
              PRALLEL_T2B_USECLS=True, PARALLEL_RESNET=True, PARALLEL_RESNET_LAYERS=[0, 1, 1, 1, 1], EARLY_CONV=True,
              EARLY_CONV_NEW_IMPLEMENT=True, N_LAYERS=1, VISUAL_LAYER_MINUS1=False, EARLY_CONV_RES=True,
              EARLY_CONV_RES_FIRSTCONV_KERNEL=3, EARLY_CONV_RES_BLOCK="basic_v0", EARLY_CONV_RES_LAYERS=[1, 1, 1, 1])
    config = ns(MODEL=ns(SPEC=ns(EMBED_DIM=512, GATHER_TENSORS=True,
                                 VISION=ns(MODEL="vit", PATCH_SIZE=32, WIDTH=768, LAYERS=2),
                                 TEXT=ns(CONTEXT_LENGTH=77, VOCAB_SIZE=49408, WIDTH=768, HEADS=12, LAYERS=2, STYLE="clip",
                                         TOKENIZER="clip"))),
                TRAIN=ns(IMAGE_SIZE=[224, 224]), CUSTOM=ns(**cu), OUTPUT_DIR=".")
    model = get_clip_model(config).cuda().eval()
    out = model.encode_image(torch.randn(2, 3, 224, 224).cuda())
    assert out.shape == (2, 512) and torch.allclose(out.norm(dim=-1), torch.ones(2, device="cuda"), atol=1e-5)


def test_prefetched_images_equal_direct_transfer():
    """msclip_stage_images (input prefetch, double buffered, FIFO) feeds the same bits as a direct host call."""
    cfg, sd_np, img, tok, z, meta = load_case("b32_l2_b8")
    model = build_model(cfg, sd_np)
    a = torch.from_numpy(img).pin_memory()
    b = torch.from_numpy(img[::-1].copy()).pin_memory()
    ref_a, ref_b = model.encode_image(a.cuda()).cpu(), model.encode_image(b.cuda()).cpu()
    model.prefetch_images(a)
    model.prefetch_images(b)                       # two slots in flight
    assert torch.equal(model.encode_image(a), ref_a)
    model.prefetch_images(a)                       # slot of `a` is reused while `b` is still pending
    assert torch.equal(model.encode_image(b), ref_b)
    assert torch.equal(model.encode_image(a), ref_a)
    ttok = torch.from_numpy(tok).pin_memory()
    l_direct = float(model.contrastive_loss(a, ttok))
    model.prefetch_images(a)
    assert float(model.contrastive_loss(a, ttok)) == l_direct


def test_internal_chunk_boundaries():
    """Batches that cross the library's internal chunking (512 images per conv pass, 4096 sequences per tower
    pass) give the same embeddings as the same samples encoded in small batches (bit-exact: every kernel is
    batch-independent and deterministic)."""
    cfg = MSCLIPConfig(layers=3)
    model = build_model(cfg, synth.synth_state_dict(cfg, seed=6))
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.randn(1030, 3, 224, 224, device="cuda", generator=g)
    big = model.encode_image(img)
    idx = [0, 511, 512, 1023, 1024, 1029]
    small = torch.cat([model.encode_image(img[i:i + 1]) for i in idx])
    assert torch.equal(big[idx], small)
    tok = torch.from_numpy(synth.synth_tokens(4100, 5, ragged=True)).cuda()
    tbig = model.encode_text(tok)
    tidx = [0, 4095, 4096, 4099]
    tsmall = torch.cat([model.encode_text(tok[i:i + 1]) for i in tidx])
    assert torch.equal(tbig[tidx], tsmall)
    assert torch.isfinite(big).all() and torch.isfinite(tbig).all()


def test_text_trim_is_bit_identical():
    """Positions after the EOT token cannot reach the pooled row of the causal text tower (M.py:2965-2971, 3057-3060):
    running the tower over the longest live prefix of the batch gives bit-identical embeddings."""
    cfg = MSCLIPConfig(layers=4)
    model = build_model(cfg, synth.synth_state_dict(cfg, seed=9))
    tok = synth.synth_tokens(96, 21, ragged=True)
    short = tok.copy()
    short[:, 20:] = 0                       # every prompt ends within 20 positions
    short[:, 19] = np.where(tok.argmax(1) >= 19, 49407, short[:, 19])
    full = tok.copy()
    full[0, :] = np.maximum(full[0, :], 1)
    full[0, 76] = 49407                     # one sequence uses the whole context
    for name, t in (("ragged", tok), ("short", short), ("full", full)):
        tt = torch.from_numpy(t).cuda()
        model.set_text_trim(True)
        a = model.encode_text(tt)
        model.set_text_trim(False)
        b = model.encode_text(tt)
        assert torch.isfinite(a).all(), name
        assert torch.equal(a, b), name
    model.set_text_trim(True)


def test_loss_matches_logits_path_at_odd_batch_sizes():
    """Fused loss kernel vs cross-entropy of the materialised logits for batches that are not multiples of the
    128-row / 128-column tiles (1, 3, 129, 257)."""
    cfg = MSCLIPConfig(layers=2)
    sd_np = synth.synth_state_dict(cfg, seed=8, logit_scale=math.log(20.0))
    model = build_model(cfg, sd_np)
    for b in (1, 3, 129, 257):
        img = torch.from_numpy(synth.synth_images(b, 40 + b)).cuda()
        tok = torch.from_numpy(synth.synth_tokens(b, 40 + b, ragged=True)).cuda()
        fused = float(model.contrastive_loss(img, tok))
        ref = loss_of(model(img, tok).cpu())
        assert abs(fused - ref) <= 1e-3 * abs(ref) + 2e-4, (b, fused, ref)


def test_micro_batched_loss_equals_one_shot():
    """SURVEY.md 8(d) config 4 on fewer GPUs than shards: the local batch goes through the towers in micro-batches whose
    embeddings are retained (msclip_encode_pairs), then ONE loss over all rows.  Every kernel is row-independent, so
    the result is bit-identical to the one-shot msclip_forward_loss."""
    cfg = MSCLIPConfig(layers=3)
    sd_np = synth.synth_state_dict(cfg, seed=9, logit_scale=math.log(1 / 0.07))
    b = 40
    img = torch.from_numpy(synth.synth_images(b, 5)).cuda()
    tok = torch.from_numpy(synth.synth_tokens(b, 5, ragged=True)).cuda()
    model = build_model(cfg, sd_np)
    one = float(model.contrastive_loss(img, tok))
    model.setup_data_parallel(b)                       # world 1: sizes the exchange buffer for the whole shard
    for micro in (8, 16, 40):                          # 40: degenerate (one micro-batch); 16: ragged last piece
        got = float(model.contrastive_loss(img, tok, micro_batch=micro))
        assert got == one, (micro, got, one)
    # out-of-order / overflowing micro-batches are refused, not mis-computed
    model.encode_pairs(img[:8], tok[:8], 0)
    with pytest.raises(Exception):
        model.encode_pairs(img[:8], tok[:8], 16)
    with pytest.raises(Exception):
        model.loss_of_encoded(b)
    ref = float(O.contrastive_loss(O.forward(img.cpu(), tok.cpu(), O.to_torch(sd_np), cfg)))
    assert abs(one - ref) <= LOSS_TOL * abs(ref)


def test_two_ranks_on_one_gpu_run_the_p2p_protocol():
    """The sharded loss (publish flags, stream waits, rank-ordered shard tables, double-buffered epochs) with world = 2
    on ONE GPU: two handles, each on its own stream, exchange buffers handed over as same-process pointers
    (msclip_comm_import_pointers).  Must equal the single-handle loss over the concatenated batch - the same check
    tests/test_multigpu.py makes across real GPUs, runnable on the single-GPU box."""
    import ctypes as C
    from msclip_b200 import _lib
    cfg = MSCLIPConfig(layers=2)
    sd_np = synth.synth_state_dict(cfg, seed=13, logit_scale=math.log(1 / 0.07))
    b = 24
    img = torch.from_numpy(synth.synth_images(2 * b, 31)).cuda()
    tok = torch.from_numpy(synth.synth_tokens(2 * b, 31, ragged=True)).cuda()
    solo = build_model(cfg, sd_np)
    truth = float(solo.contrastive_loss(img, tok))
    gi_solo, gt_solo = solo.contrastive_loss_backward()
    L = _lib.lib("bf16")
    ranks = [build_model(cfg, sd_np) for _ in range(2)]
    bases = (C.c_void_p * 2)()
    for r, m in enumerate(ranks):
        m._sync_weights()
        _lib.check(L.msclip_comm_init(m._handle, r, 2, b), "comm_init")
        base = C.c_void_p()
        _lib.check(L.msclip_comm_buffer(m._handle, C.byref(base)), "comm_buffer")
        bases[r] = base.value
    for m in ranks:
        _lib.check(L.msclip_comm_import_pointers(m._handle, bases), "comm_import_pointers")
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    parts = torch.zeros(2, 2, device="cuda")
    torch.cuda.synchronize()
    for epoch in range(3):                             # both parities of the exchange buffer, then reuse
        for r, m in enumerate(ranks):                  # nothing blocks the host: rank 0's loss is queued behind a
            sp = C.c_void_p(streams[r].cuda_stream)    # stream wait on rank 1's flag, which rank 1's stream raises
            lo = r * b
            _lib.check(L.msclip_forward_loss(m._handle, C.c_void_p(img[lo:lo + b].data_ptr()), _lib.F32,
                                             C.c_void_p(tok[lo:lo + b].data_ptr()), b, C.c_void_p(parts[r].data_ptr()),
                                             None, sp), "forward_loss")
        torch.cuda.synchronize()
        got = float(parts.sum()) / (2.0 * 2 * b)
        assert abs(got - truth) <= 2e-6 * abs(truth), (epoch, got, truth)
        # backward of the same step: each rank gets the gradient rows of ITS shard (second peer read: the peers' row lse)
        grads = [(torch.empty(b, 512, device="cuda"), torch.empty(b, 512, device="cuda")) for _ in range(2)]
        for r, m in enumerate(ranks):
            sp = C.c_void_p(streams[r].cuda_stream)
            _lib.check(L.msclip_contrastive_loss_backward(m._handle, C.c_void_p(grads[r][0].data_ptr()),
                                                          C.c_void_p(grads[r][1].data_ptr()), sp), "loss_backward")
        torch.cuda.synchronize()
        gi = torch.cat([grads[0][0], grads[1][0]])
        gt = torch.cat([grads[0][1], grads[1][1]])
        assert rel_err(gi.cpu().numpy(), gi_solo.cpu().numpy()) <= 1e-5 and rel_err(gt.cpu().numpy(), gt_solo.cpu().numpy()) <= 1e-5


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("b,scale", [(40, 1 / 0.07), (300, 100.0), (129, math.e)])
def test_contrastive_loss_backward_matches_autograd(b, scale, precision):
    """SURVEY.md 8f-1, first piece: d loss / d (normalised embeddings) from msclip_contrastive_loss_backward against
    torch.autograd (float64) on the symmetric cross-entropy of the same fp16-rounded embeddings.  The probabilities
    and the transposed embeddings enter the gradient GEMM as 16-bit operands: <= 1e-2 (bf16) / 2e-3 (fp16) relative."""
    import ctypes as C
    from msclip_b200 import _lib
    cfg = MSCLIPConfig(layers=2)
    model = build_model(cfg, synth.synth_state_dict(cfg, seed=3), precision)
    model._sync_weights()
    g = torch.Generator(device="cuda").manual_seed(b)
    fi = torch.nn.functional.normalize(torch.randn(b, 512, device="cuda", generator=g), dim=-1)
    ft = torch.nn.functional.normalize(fi * 0.6 + 0.1 * torch.randn(b, 512, device="cuda", generator=g), dim=-1)
    L = _lib.lib(precision)
    parts, loss = torch.zeros(2, device="cuda"), torch.zeros((), device="cuda")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.msclip_contrastive_loss_features(model._handle, C.c_void_p(fi.data_ptr()), C.c_void_p(ft.data_ptr()), b, scale,
                                                  C.c_void_p(parts.data_ptr()), C.c_void_p(loss.data_ptr()), sp), "loss", precision)
    model._last_loss_b = b
    gi, gt = model.contrastive_loss_backward()
    a = fi.half().double().requires_grad_(True)
    t = ft.half().double().requires_grad_(True)
    logits = scale * a @ t.t()
    tgt = torch.arange(b, device="cuda")
    ref = 0.5 * (torch.nn.functional.cross_entropy(logits, tgt) + torch.nn.functional.cross_entropy(logits.t(), tgt))
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-6
    ei, et = rel_err(gi.cpu().numpy(), a.grad.float().cpu().numpy()), rel_err(gt.cpu().numpy(), t.grad.float().cpu().numpy())
    _record(f"loss_backward/{precision}/b{b}", {"d_image_features": ei, "d_text_features": et})
    tol = 1e-2 if precision == "bf16" else 2e-3
    assert ei <= tol and et <= tol, (ei, et)
