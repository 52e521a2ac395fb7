"""The CPU oracle's BACKWARD (torch.autograd through oracle/msclip_oracle.py) against the gradients of the REAL reference
(tests/golden/grad_*.npz, oracle/make_golden_grads.py): pins the checker the GPU backward is held to
(tests/test_backward_gpu.py).  fp32 on both sides: 2e-4 Frobenius-relative on the samples, 1e-4 on the norms."""
import numpy as np
import pytest
import torch

from oracle import msclip_oracle as O
from golden_util import GRAD_CASES, grad_sample, load_case, load_grad_case, rel_err, trainable_keys


def oracle_grads(cfg, sd_np, img, tok):
    sd = O.to_torch(sd_np)
    leaves = {}
    for k, t in sd.items():
        if t.dtype == torch.float32 and id(t) not in leaves:
            t.requires_grad_(True)
            leaves[id(t)] = t
    logits = O.forward(torch.from_numpy(img), torch.from_numpy(tok), sd, cfg)
    loss = O.contrastive_loss(logits)
    loss.backward()
    return {k: sd[k].grad for k in trainable_keys(cfg)}, float(loss.detach())


@pytest.mark.parametrize("name", GRAD_CASES[:2])
def test_oracle_backward_matches_reference_gradients(name):
    cfg, sd_np, img, tok, _z, _meta = load_case(name)
    gz = load_grad_case(name)
    grads, loss = oracle_grads(cfg, sd_np, img, tok)
    worst = 0.0
    for key in trainable_keys(cfg):
        g = grads[key]
        assert g is not None, key
        g = g.numpy()
        ref_norm = float(gz["norm/" + key])
        assert abs(np.linalg.norm(g.astype(np.float64)) - ref_norm) <= 1e-4 * ref_norm + 1e-12, key
        e = rel_err(grad_sample(g, key, tok), gz["sample/" + key])
        worst = max(worst, e)
        assert e < 2e-4, (key, e)
    assert worst < 2e-4
