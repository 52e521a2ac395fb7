"""Golden GRADIENTS: backward of the symmetric cross-entropy through the REAL reference module (eval mode, fp32, CPU).

TEST INFRASTRUCTURE (authoring container only - needs /root/reference).  For the seeded cases of oracle/make_golden.py it
runs ``loss = 0.5 * (CE(model(img, tok)) + CE(model(img, tok).T))`` on the unmodified reference (oracle/ref_shim.py),
calls ``loss.backward()`` and stores, for every key msclip_backward produces (tests/golden_util.trainable_keys), the
Frobenius norm of the gradient and a deterministic sample of it -> tests/golden/grad_<case>.npz.  The reference ships no
backward of its own: torch.autograd on its forward IS the specification.

    python oracle/make_golden_grads.py [case ...]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim                      # noqa: E402
from oracle.make_golden import case_inputs       # noqa: E402
from golden_util import GOLDEN_DIR, GRAD_CASES, grad_sample, trainable_keys   # noqa: E402


def reference_grads(cfg, sd, img, tok, autocast):
    model = ref_shim.build_reference_model(cfg, sd)
    model.eval()
    for p in model.parameters():
        p.requires_grad_(True)
    if autocast:
        # the reference's own bf16 mode (as in make_golden.py): the yardstick for a bf16-operand backward
        with torch.autocast("cpu", dtype=torch.bfloat16):
            logits = model(torch.from_numpy(img), torch.from_numpy(tok)).float()
    else:
        logits = model(torch.from_numpy(img), torch.from_numpy(tok))
    tgt = torch.arange(logits.shape[0])
    loss = 0.5 * (F.cross_entropy(logits, tgt) + F.cross_entropy(logits.t(), tgt))
    loss.backward()
    params = dict(model.named_parameters(remove_duplicate=False))
    return {k: params[k].grad.detach().numpy().copy() for k in trainable_keys(cfg)}, float(loss.detach())


def run_case(name):
    cfg, sd, img, tok = case_inputs(name)
    grads, loss = reference_grads(cfg, sd, img, tok, False)
    grads_ac, _ = reference_grads(cfg, sd, img, tok, True)
    out = {}
    num = den = 0.0
    for key in trainable_keys(cfg):
        g = grads[key]
        out["norm/" + key] = np.float64(np.linalg.norm(g.astype(np.float64)))
        out["sample/" + key] = grad_sample(g, key, tok)
        out["sample_autocast/" + key] = grad_sample(grads_ac[key], key, tok)
        num += float(np.linalg.norm(out["sample_autocast/" + key].astype(np.float64) - out["sample/" + key]) ** 2)
        den += float(np.linalg.norm(out["sample/" + key].astype(np.float64)) ** 2)
    out["meta"] = json.dumps({"case": name, "loss": loss, "torch": torch.__version__,
                              "autocast_aggregate": (num / den) ** 0.5})
    print(f"  reference autocast(bf16) backward vs fp32, sample aggregate: {(num / den) ** 0.5:.4f}")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "grad_" + name + ".npz"), **out)
    print(f"{name}: loss {float(loss):.6f}, {len(trainable_keys(cfg))} gradient tensors")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for case in (sys.argv[1:] or GRAD_CASES):
        run_case(case)
