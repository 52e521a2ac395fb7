"""Deterministic synthetic weights and inputs.

There is no network for checkpoints or datasets, and ``/root/reference`` does not exist on the GPU
box, so every test/bench input is regenerated from a seed with numpy's MT19937 ``RandomState``
(bit-stable across numpy versions and machines).  ``state_dict_spec`` is the reference's state-dict
contract (521 keys for 12 layers, of which the text blocks 1..11 alias the vision
tensors, lib/models/clip_openai_pe_res_v1.py:2786-2830); tests/test_oracle.py checks it against the
key list exported from the reference itself (tests/golden/state_dict_keys_*.json).

Unlike the reference's own init (trunc-normal 0.02 everywhere, BN identity, zero biases — under
which the conv stem contributes ~1e-5 to the tokens and a parity test has no teeth), the synthetic
weights keep every stage at O(1) activation scale and give every bias / BN statistic / LN affine a
non-trivial value.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

from .config import MSCLIPConfig

SOT, EOT = 49406, 49407     # lib/dataset/languages/simple_tokenizer.py:141-145


def _bn(spec, prefix, c):
    spec[prefix + ".weight"] = (c,)
    spec[prefix + ".bias"] = (c,)
    spec[prefix + ".running_mean"] = (c,)
    spec[prefix + ".running_var"] = (c,)
    spec[prefix + ".num_batches_tracked"] = ()


def _ln(spec, prefix, c):
    spec[prefix + ".weight"] = (c,)
    spec[prefix + ".bias"] = (c,)


def _block(spec, prefix, w):
    # module registration order of ResidualAttentionBlock (M.py:789-799)
    spec[prefix + ".attn.in_proj_weight"] = (3 * w, w)
    spec[prefix + ".attn.in_proj_bias"] = (3 * w,)
    spec[prefix + ".attn.out_proj.weight"] = (w, w)
    spec[prefix + ".attn.out_proj.bias"] = (w,)
    _ln(spec, prefix + ".ln_1", w)
    spec[prefix + ".mlp.c_fc.weight"] = (4 * w, w)
    spec[prefix + ".mlp.c_fc.bias"] = (4 * w,)
    spec[prefix + ".mlp.c_proj.weight"] = (w, 4 * w)
    spec[prefix + ".mlp.c_proj.bias"] = (w,)
    _ln(spec, prefix + ".ln_2", w)


SHARED_SUFFIXES = (
    ".attn.in_proj_weight", ".attn.in_proj_bias", ".attn.out_proj.weight", ".attn.out_proj.bias",
    ".mlp.c_fc.weight", ".mlp.c_fc.bias", ".mlp.c_proj.weight", ".mlp.c_proj.bias",
)


def state_dict_spec(cfg: MSCLIPConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape for every key of the reference ``CLIP.state_dict()`` (order not significant)."""
    w, e = cfg.width, cfg.embed_dim
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    spec["positional_embedding"] = (cfg.context_length, w)
    spec["text_projection"] = (w, e)
    spec["logit_scale"] = ()
    v = "visual."
    spec[v + "class_embedding"] = (w,)
    spec[v + "positional_embedding"] = (cfg.image_tokens, w)
    spec[v + "proj"] = (w, e)
    _ln(spec, v + "ln_pre", w)
    # vision "block 0" = early-conv residual stem (M.py:1939-2000)
    s = v + "transformer.resblocks.0."
    c0 = w // 16
    spec[s + "conv1.weight"] = (c0, 3, 3, 3)
    _bn(spec, s + "bn1", c0)
    c = c0
    for i in range(4):
        p = s + f"resnet_stage.conv_{i}."
        spec[p + "conv1.weight"] = (2 * c, c, 3, 3)
        _bn(spec, p + "bn1", 2 * c)
        spec[p + "downsample.0.weight"] = (2 * c, c, 1, 1)
        _bn(spec, p + "downsample.1", 2 * c)
        c *= 2
    spec[s + "last_conv.weight"] = (w, w, 1, 1)
    for i in range(1, cfg.layers):
        _block(spec, v + f"transformer.resblocks.{i}", w)
    # parallel branch (M.py:2128-2159): stage 0 conv-bn-relu, stages 1..4 one bottleneck each
    pb = v + "transformer.parallel_branch_v."
    spec[pb + "0.conv.weight"] = (c0, 3, 3, 3)
    _bn(spec, pb + "0.bn", c0)
    dims = cfg.branch_dims
    for j in range(1, 5):
        cin, cout = dims[j - 1], dims[j]
        mid = cout // 2
        p = pb + f"{j}.resnet_stage.conv_0."
        spec[p + "conv1.weight"] = (mid, cin, 1, 1)
        _bn(spec, p + "bn1", mid)
        spec[p + "conv2.weight"] = (mid, mid, 3, 3)
        _bn(spec, p + "bn2", mid)
        spec[p + "conv3.weight"] = (cout, mid, 1, 1)
        _bn(spec, p + "bn3", cout)
        spec[p + "residual_conv.weight"] = (cout, cin, 1, 1)
        _bn(spec, p + "residual_bn", cout)
    # lateral adapters (M.py:1556-1637)
    la = v + "transformer.parallel_lateral_adapter."
    for j in range(5):
        cj, k = dims[j], cfg.t2b_kernels[j]
        spec[la + f"{j}.top2bottom_dw_conv.conv.weight"] = (cj, 1, k, k)
        _bn(spec, la + f"{j}.top2bottom_dw_conv.bn", cj)
        spec[la + f"{j}.top2bottom_pw_conv.conv.weight"] = (w, cj, 1, 1)
        spec[la + f"{j}.bottom_dw_conv.conv.weight"] = (w, 1, 3, 3)
        _bn(spec, la + f"{j}.bottom_dw_conv.bn", w)
        _ln(spec, la + f"{j}.ln_adapt", w)
    _ln(spec, v + "ln_post", w)
    for i in range(cfg.layers):
        _block(spec, f"transformer.resblocks.{i}", w)
    spec["token_embedding.weight"] = (cfg.vocab_size, w)
    _ln(spec, "ln_final", w)
    return spec


def alias_of(cfg: MSCLIPConfig, key: str):
    """Text-tower keys of blocks 1..layers-1 that are the *same tensor* as a vision key."""
    if not key.startswith("transformer.resblocks."):
        return None
    idx = int(key.split(".")[2])
    if idx == 0 or idx >= cfg.layers:
        return None
    for suf in SHARED_SUFFIXES:
        if key.endswith(suf):
            return "visual." + key
    return None


def _rng(seed: int, key: str) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF)


def synth_state_dict(cfg: MSCLIPConfig, seed: int = 0, logit_scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded fp32 weights for every key of ``state_dict_spec`` (aliases share one array)."""
    out: Dict[str, np.ndarray] = {}
    for key, shape in state_dict_spec(cfg).items():
        src = alias_of(cfg, key)
        if src is not None:
            out[key] = out[src]
            continue
        r = _rng(seed, key)
        if key == "logit_scale":
            a = np.array(logit_scale, dtype=np.float32)
        elif key.endswith("num_batches_tracked"):
            a = np.array(0, dtype=np.int64)
        elif key.endswith("running_var"):
            a = (0.6 + 0.8 * r.random_sample(shape)).astype(np.float32)
        elif key.endswith("running_mean"):
            a = (0.1 * r.standard_normal(shape)).astype(np.float32)
        elif ".bn" in key or "downsample.1" in key or "residual_bn" in key or ".ln_" in key or key.startswith("ln_final") \
                or "ln_pre" in key or "ln_post" in key or "ln_adapt" in key:
            if key.endswith(".weight"):
                a = (1.0 + 0.1 * r.standard_normal(shape)).astype(np.float32)
            else:
                a = (0.1 * r.standard_normal(shape)).astype(np.float32)
        elif key.endswith("bias"):
            a = (0.05 * r.standard_normal(shape)).astype(np.float32)
        elif key.endswith("conv.weight") or "conv1.weight" in key or "conv2.weight" in key or "conv3.weight" in key \
                or "downsample.0.weight" in key or "residual_conv.weight" in key or "last_conv.weight" in key:
            fan_in = int(np.prod(shape[1:]))
            gain = 1.0 if "last_conv" in key or "pw_conv" in key else 1.3
            a = (gain / np.sqrt(fan_in) * r.standard_normal(shape)).astype(np.float32)
        elif key.endswith("in_proj_weight") or key.endswith("c_fc.weight"):
            a = (1.0 / np.sqrt(shape[1]) * r.standard_normal(shape)).astype(np.float32)
        elif key.endswith("out_proj.weight") or key.endswith("c_proj.weight"):
            a = (0.5 / np.sqrt(shape[1]) * r.standard_normal(shape)).astype(np.float32)
        elif key.endswith("proj") or key == "text_projection":
            a = (1.0 / np.sqrt(shape[0]) * r.standard_normal(shape)).astype(np.float32)
        elif key == "token_embedding.weight":
            a = (0.5 * r.standard_normal(shape)).astype(np.float32)
        elif "positional_embedding" in key or "class_embedding" in key:
            a = (0.3 * r.standard_normal(shape)).astype(np.float32)
        else:
            raise KeyError(f"no synthetic rule for {key}")
        out[key] = a
    return out


def synth_images(batch: int, seed: int = 1234, resolution: int = 224, offset: int = 0) -> np.ndarray:
    """``randn(B,3,R,R)`` fp32 NCHW (≈ mean/std-normalised pixels).

    Row ``i`` depends only on ``(seed, offset+i)`` so shards of a global batch can be generated
    independently on each rank.
    """
    out = np.empty((batch, 3, resolution, resolution), dtype=np.float32)
    for i in range(batch):
        out[i] = np.random.RandomState((seed * 1000003 + offset + i) & 0xFFFFFFFF).standard_normal(
            (3, resolution, resolution)).astype(np.float32)
    return out


def synth_tokens(batch: int, seed: int = 1234, context_length: int = 77, vocab_size: int = 49408,
                 ragged: bool = False, offset: int = 0) -> np.ndarray:
    """int64 ``[B, ctx]`` token ids: SOT, random body, EOT (= the largest id, so argmax finds it).

    ``ragged=False`` puts EOT at the last position (all 77 positions live); ``ragged=True`` puts it at
    a random position in [5, ctx-1] and zero-pads after it (real-data shape, M.py:3057-3060).
    """
    sot, eot = vocab_size - 2, vocab_size - 1
    out = np.zeros((batch, context_length), dtype=np.int64)
    for i in range(batch):
        r = np.random.RandomState((seed * 7919 + 17 + offset + i) & 0xFFFFFFFF)
        body = r.randint(1, sot, size=context_length)
        pos = context_length - 1 if not ragged else int(r.randint(min(5, context_length - 1), context_length))
        out[i, :pos] = body[:pos]
        out[i, 0] = sot
        out[i, pos] = eot
    return out


def correlated_pair_batch(cfg: MSCLIPConfig, batch: int, seed: int = 7):
    """A batch whose image and text embeddings are *not* independent: tokens are derived from a
    coarse signature of the image so the contrastive loss departs from ln(B)."""
    imgs = synth_images(batch, seed, cfg.image_resolution)
    toks = synth_tokens(batch, seed, cfg.context_length, cfg.vocab_size)
    sig = (imgs.reshape(batch, 3, -1)[:, :, :64] > 0).astype(np.int64)
    sig = sig.reshape(batch, -1)[:, : cfg.context_length - 2]
    toks[:, 1:1 + sig.shape[1]] = 1 + (sig * 997 + np.arange(sig.shape[1])[None, :] * 31) % (cfg.vocab_size - 3)
    return imgs, toks
