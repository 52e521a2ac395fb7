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
