"""Shared helpers: load a golden case and regenerate its seeded weights/inputs."""
import json
import os

import numpy as np

from msclip_b200.config import MSCLIPConfig
from msclip_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["b32_l2_b8", "b32_l3_b4", "b32_l12_b8", "b16_l3_b2", "b16_l12_b4"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    c = meta["cfg"]
    cfg = MSCLIPConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in c.items()})
    sd = synth.synth_state_dict(cfg, seed=meta["weight_seed"], logit_scale=meta["logit_scale"])
    if meta["correlated"]:
        img, tok = synth.correlated_pair_batch(cfg, meta["batch"], seed=meta["input_seed"])
    else:
        img = synth.synth_images(meta["batch"], meta["input_seed"], cfg.image_resolution)
        tok = synth.synth_tokens(meta["batch"], meta["input_seed"], cfg.context_length, cfg.vocab_size,
                                 ragged=meta["ragged"])
    return cfg, sd, img, tok, z, meta


def rel_err(a, b):
    """Frobenius-norm relative error ||a-b|| / ||b||."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# ---- gradient fixtures (oracle/make_golden_grads.py): backward of the symmetric CE through the REAL reference ----------
GRAD_CASES = ["b32_l2_b8", "b32_l3_b4", "b32_l12_b8", "b16_l3_b2"]
GRAD_SAMPLE = 2048


def trainable_keys(cfg):
    """State-dict keys that receive a gradient from msclip_backward (include/msclip_b200.h): everything but the frozen
    convolutional front.  Aliased text keys (blocks >= 1) are listed under their vision name only."""
    block = ["attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias", "mlp.c_fc.weight",
             "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias"]
    lns = ["ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias"]
    keys = ["logit_scale", "positional_embedding", "text_projection", "token_embedding.weight", "ln_final.weight", "ln_final.bias",
            "visual.class_embedding", "visual.positional_embedding", "visual.proj", "visual.ln_pre.weight", "visual.ln_pre.bias",
            "visual.ln_post.weight", "visual.ln_post.bias"]
    for i in range(1, cfg.layers):
        keys += [f"visual.transformer.resblocks.{i}.{s}" for s in block + lns]
    keys += [f"transformer.resblocks.0.{s}" for s in block]
    for i in range(cfg.layers):
        keys += [f"transformer.resblocks.{i}.{s}" for s in lns]
    for j in cfg.active_laterals():
        keys += [f"visual.transformer.parallel_lateral_adapter.{j}.ln_adapt.{s}" for s in ("weight", "bias")]
    return keys


def grad_sample(g, key, tok=None):
    """A small, deterministic sample of a gradient tensor: strided over the flat tensor; for the token embedding the
    rows of the tokens that occur (all other rows are exactly zero)."""
    g = np.asarray(g, dtype=np.float32)
    if key == "token_embedding.weight" and tok is not None:
        rows = np.unique(np.asarray(tok))[:96]
        return g[rows][:, ::11].reshape(-1)
    flat = g.reshape(-1)
    stride = max(1, flat.size // GRAD_SAMPLE)
    return flat[::stride][:GRAD_SAMPLE].copy()


def load_grad_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, "grad_" + name + ".npz"))
    return z
