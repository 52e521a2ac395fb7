"""Pin the CPU oracle (oracle/msclip_oracle.py) against the reference's golden vectors.

The golden files were produced by the real reference (oracle/make_golden.py).  fp32 vs fp32, same
torch build: the only difference is operation order, so the tolerance is 2e-5 Frobenius-relative on
features/logits/activations and 1e-6 relative on the loss.
"""
import json
import os

import numpy as np
import pytest
import torch

from msclip_b200 import synth
from msclip_b200.config import MSCLIPConfig
from oracle import msclip_oracle as O
from oracle import ref_shim
from golden_util import CASES, GOLDEN_DIR, load_case, rel_err

TOL = 2e-5


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    if name == "b16_l12_b4" and os.environ.get("MSCLIP_FAST_TESTS"):
        pytest.skip("fast mode")
    cfg, sd_np, img, tok, z, meta = load_case(name)
    sd = O.to_torch(sd_np)
    timg, ttok = torch.from_numpy(img), torch.from_numpy(tok)
    taps = {}
    with torch.no_grad():
        fi = O.encode_image(timg, sd, cfg, taps=taps)
        ft = O.encode_text(ttok, sd, cfg, taps=taps)
        fi_raw = O.encode_image(timg, sd, cfg, norm=False)
        logits = O.similarity_logits(fi, ft, sd["logit_scale"])
        loss = O.contrastive_loss(logits)
    assert rel_err(fi.numpy(), z["image_features"]) < TOL
    assert rel_err(ft.numpy(), z["text_features"]) < TOL
    assert rel_err(fi_raw.numpy(), z["image_features_unnormalised"]) < TOL
    assert rel_err(logits.numpy(), z["logits"]) < TOL
    assert abs(loss.item() - float(z["loss"])) <= 1e-6 * abs(float(z["loss"])) + 1e-6
    sb, st, sc = meta["tap_stride"]
    n = 0
    for key in z.files:
        if key.startswith("tap_"):
            got = taps[key[4:]][::sb, ::st, ::sc].numpy()
            assert rel_err(got, z[key]) < TOL, key
            n += 1
    assert n == (cfg.layers - 1) + len(cfg.active_laterals()) + cfg.layers


@pytest.mark.parametrize("tag", ["b32_l2", "b32_l3", "b32_l12", "b16_l3", "b16_l12"])
def test_state_dict_contract(tag):
    """Our key/shape list equals the one exported from the reference's own state_dict()."""
    with open(os.path.join(GOLDEN_DIR, f"state_dict_keys_{tag}.json")) as f:
        ref = json.load(f)
    p, l = tag.split("_")
    cfg = MSCLIPConfig(patch_size=int(p[1:]), layers=int(l[1:]))
    spec = synth.state_dict_spec(cfg)
    assert set(spec) == set(ref)
    for k, shape in spec.items():
        assert list(shape) == ref[k], k
    n_alias = sum(1 for k in spec if synth.alias_of(cfg, k))
    assert n_alias == 8 * (cfg.layers - 1)
    if cfg.layers == 12:
        assert len(spec) == 521 and n_alias == 88


def test_eot_pooling_ignores_tokens_after_eot():
    """encode_text is independent of what follows EOT (causal mask + argmax pooling, M.py:3059)."""
    cfg = MSCLIPConfig(layers=2)
    sd = O.to_torch(synth.synth_state_dict(cfg, seed=5))
    tok = synth.synth_tokens(3, 11, ragged=True)
    tok2 = tok.copy()
    for i in range(3):
        e = tok2[i].argmax()
        tok2[i, e + 1:] = np.arange(1, 77 - e)[: 76 - e] * 3 % 1000
    with torch.no_grad():
        a = O.encode_text(torch.from_numpy(tok), sd, cfg)
        b = O.encode_text(torch.from_numpy(tok2), sd, cfg)
    assert torch.equal(a, b)


def test_loss_properties():
    """Loss oracle: equals ln(B) for constant logits, symmetric under transpose, matches the
    explicit log-sum-exp formula of SURVEY.md Appendix A."""
    g = torch.Generator().manual_seed(0)
    s = torch.randn(37, 37, generator=g) * 3
    l = O.contrastive_loss(s)
    assert abs(O.contrastive_loss(torch.zeros(16, 16)).item() - np.log(16)) < 1e-6
    assert abs(l.item() - O.contrastive_loss(s.t()).item()) < 1e-6
    d = s.diag()
    ref = 0.5 * ((torch.logsumexp(s, 1) - d).mean() + (torch.logsumexp(s, 0) - d).mean())
    assert abs(l.item() - ref.item()) < 1e-5


def test_gather_rank_order_equals_single_process_logits():
    """Rank-ordered concat of per-rank shards (comm.py:150-153) reproduces the un-gathered logits."""
    g = torch.Generator().manual_seed(1)
    fi, ft = torch.randn(8, 16, generator=g), torch.randn(8, 16, generator=g)
    full = O.similarity_logits(fi, ft, 0.5)
    fi_all = O.gather_rank_order([fi[0:4], fi[4:8]])
    ft_all = O.gather_rank_order([ft[0:4], ft[4:8]])
    assert torch.equal(O.similarity_logits(fi_all, ft_all, 0.5), full)


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.reference_available(), reason="needs /root/reference")
def test_oracle_matches_live_reference_fresh_seed():
    """Container only: a case that is NOT in the golden set, reference executed live."""
    cfg = MSCLIPConfig(layers=3)
    sd_np = synth.synth_state_dict(cfg, seed=77, logit_scale=2.0)
    img, tok = synth.synth_images(3, 8), synth.synth_tokens(3, 8, ragged=True)
    model = ref_shim.build_reference_model(cfg, sd_np)
    sd = O.to_torch(sd_np)
    with torch.no_grad():
        ref = model(torch.from_numpy(img), torch.from_numpy(tok)).numpy()
        got = O.forward(torch.from_numpy(img), torch.from_numpy(tok), sd, cfg).numpy()
    assert rel_err(got, ref) < TOL
