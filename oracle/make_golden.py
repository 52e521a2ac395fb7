"""Generate tests/golden/*.npz by running the REAL reference (authoring container only).

TEST INFRASTRUCTURE.  Usage:  python -m oracle.make_golden   (needs /root/reference)

Every case = (config, weight seed, input recipe).  Weights and inputs are regenerated from seeds by
msclip_b200.synth, so only the reference's *outputs* are committed: normalised image/text features,
logits, the symmetric-CE loss of those logits, strided samples of the per-block activations
(captured with forward hooks on the reference modules) and the reference's state-dict key list.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from msclip_b200.config import MSCLIPConfig           # noqa: E402
from msclip_b200 import synth                          # noqa: E402
from oracle import ref_shim                            # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (config kwargs, batch, weight seed, input seed, ragged tokens, logit_scale, correlated)
CASES = {
    "b32_l2_b8":   (dict(patch_size=32, layers=2), 8, 0, 1234, False, 1.0, False),
    "b32_l3_b4":   (dict(patch_size=32, layers=3), 4, 1, 99, True, float(np.log(1 / 0.07)), False),
    "b32_l12_b8":  (dict(patch_size=32, layers=12), 8, 2, 4321, False, float(np.log(100.0)), True),
    "b16_l3_b2":   (dict(patch_size=16, layers=3), 2, 3, 5, True, 1.0, False),
    "b16_l12_b4":  (dict(patch_size=16, layers=12), 4, 4, 6, False, float(np.log(1 / 0.07)), False),
}

TAP_STRIDE = (1, 7, 13)     # batch, token, channel strides of the activation samples


def case_inputs(name):
    kw, batch, wseed, iseed, ragged, ls, corr = CASES[name]
    cfg = MSCLIPConfig(**kw)
    sd = synth.synth_state_dict(cfg, seed=wseed, logit_scale=ls)
    if corr:
        img, tok = synth.correlated_pair_batch(cfg, batch, seed=iseed)
    else:
        img = synth.synth_images(batch, iseed, cfg.image_resolution)
        tok = synth.synth_tokens(batch, iseed, cfg.context_length, cfg.vocab_size, ragged=ragged)
    return cfg, sd, img, tok


def sample(x):
    sb, st, sc = TAP_STRIDE
    return x[::sb, ::st, ::sc].contiguous().numpy()


def run_case(name):
    cfg, sd, img, tok = case_inputs(name)
    model = ref_shim.build_reference_model(cfg, sd)
    taps = {}

    def hook(tag, lnd=True):
        def fn(_m, _inp, out):
            o = out[1] if isinstance(out, tuple) else out       # Lateral_Adapter returns (top, bottom)
            if tag not in taps:      # first (fp32) forward only; later forwards (norm=False, autocast) must not overwrite
                taps[tag] = sample(o.detach().permute(1, 0, 2) if lnd else o.detach())
        return fn

    vt = model.visual.transformer
    for i in range(1, cfg.layers):
        vt.resblocks[i].register_forward_hook(hook(f"v_block{i}"))
    for j in cfg.active_laterals():
        vt.parallel_lateral_adapter[j].register_forward_hook(hook(f"v_adapter{j}"))
    for i in range(cfg.layers):
        model.transformer.resblocks[i].register_forward_hook(hook(f"t_block{i}"))
    with torch.no_grad():
        timg, ttok = torch.from_numpy(img), torch.from_numpy(tok)
        fi = model.encode_image(timg)
        ft = model.encode_text(ttok)
        fi_raw = model.encode_image(timg, norm=False)
        logits = model(timg, ttok)
        tgt = torch.arange(logits.shape[0])
        loss = 0.5 * (F.cross_entropy(logits, tgt) + F.cross_entropy(logits.t(), tgt))
        # the reference's own bf16 mode (torch.autocast; model.bfloat16() is unusable, M.py:689-691): the
        # yardstick for "how far may a bf16-operand pipeline be from the fp32 forward" 
        with torch.autocast("cpu", dtype=torch.bfloat16):
            fi_ac = model.encode_image(timg).float()
            ft_ac = model.encode_text(ttok).float()
            logits_ac = model(timg, ttok).float()
        loss_ac = 0.5 * (F.cross_entropy(logits_ac, tgt) + F.cross_entropy(logits_ac.t(), tgt))
    keys = {k: list(v.shape) for k, v in model.state_dict().items()}
    out = dict(
        meta=json.dumps(dict(case=name, cfg=cfg.to_dict(), batch=CASES[name][1], weight_seed=CASES[name][2],
                             input_seed=CASES[name][3], ragged=CASES[name][4], logit_scale=CASES[name][5],
                             correlated=CASES[name][6], tap_stride=TAP_STRIDE, torch=torch.__version__)),
        image_features=fi.numpy(), text_features=ft.numpy(), image_features_unnormalised=fi_raw.numpy(),
        logits=logits.numpy(), loss=np.float64(loss.item()),
        image_features_autocast=fi_ac.numpy(), text_features_autocast=ft_ac.numpy(),
        logits_autocast=logits_ac.numpy(), loss_autocast=np.float64(loss_ac.item()),
    )
    out.update({"tap_" + k: v for k, v in taps.items()})
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    tag = f"b{cfg.patch_size}_l{cfg.layers}"
    with open(os.path.join(GOLDEN_DIR, f"state_dict_keys_{tag}.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print(f"{name}: loss {loss.item():.6f}  ln(B) {np.log(CASES[name][1]):.6f}  |logits| max {logits.abs().max():.3f}"
          f"  keys {len(keys)}  taps {len(taps)}")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name in (sys.argv[1:] or CASES):
        run_case(name)


if __name__ == "__main__":
    main()
