"""Fused AdamW for the trainable parameters of the drop-in CLIP.

The reference ships no training loop; its configuration names the optimiser: AdamW, lr 1e-4, weight decay 0.05 with
none on biases / LayerNorms / BatchNorms (experiments/model/b32.yaml:32-53, WITHOUT_WD_LIST) and a separate learning
rate / weight decay for the modality-shared modules (b32-yfcc-msclips.yaml:10-14: SHARE_MODULES, LR_SHARE, WD_SHARE).
This class mirrors ``torch.optim.AdamW`` (same update rule, decoupled decay) with those parameter groups and applies
the step to every tensor with ONE kernel launch (msclip_op_adamw), then re-packs the 16-bit weight copies the forward
consumes (msclip_update_weight) - no host round trip, no re-upload of the state dict.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib

SHARE_MODULES = ("attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj", "mlp")     # b32-yfcc-msclips.yaml:10
WITHOUT_WD = ("bn", "bias", "ln")                                                        # b32.yaml:49


def _is_shared(key: str) -> bool:
    """Block parameters of layers >= 1 that both towers use (M.py:2786-2830)."""
    parts = key.split(".")
    try:
        idx = int(parts[parts.index("resblocks") + 1])
    except (ValueError, IndexError):
        return False
    return idx >= 1 and any(m in key for m in SHARE_MODULES)


def _no_decay(key: str) -> bool:
    return any(tag in key for tag in WITHOUT_WD) or key in ("logit_scale",) or "embedding" in key


class AdamW:
    def __init__(self, model, lr: float = 1e-4, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.05, lr_share: Optional[float] = None, wd_share: Optional[float] = None,
                 keys: Optional[Sequence[str]] = None):
        self.model = model
        self.betas, self.eps = betas, eps
        named: Dict[str, torch.nn.Parameter] = model.trainable_parameters()
        if keys is not None:
            named = {k: named[k] for k in keys}
        self.entries = []          # (key, param, lr, wd), one per distinct Parameter
        seen = set()
        for key, p in named.items():
            if p.data_ptr() in seen:
                continue
            seen.add(p.data_ptr())
            shared = _is_shared(key)
            plr = lr_share if (shared and lr_share is not None) else lr
            pwd = 0.0 if _no_decay(key) else (wd_share if (shared and wd_share is not None) else weight_decay)
            self.entries.append([key, p, float(plr), float(pwd)])
        self.state = {e[0]: (torch.zeros_like(e[1]), torch.zeros_like(e[1])) for e in self.entries}
        self.steps = 0

    def set_lr(self, lr: float, lr_share: Optional[float] = None) -> None:
        for e in self.entries:
            e[2] = float(lr_share) if (lr_share is not None and _is_shared(e[0])) else float(lr)

    def zero_grad(self) -> None:
        self.model.zero_grad()

    def state_dict(self):
        """{'steps', 'betas', 'eps', 'state': {key: (exp_avg, exp_avg_sq)}, 'groups': {key: (lr, weight_decay)}} on the CPU."""
        return {"steps": self.steps, "betas": tuple(self.betas), "eps": self.eps,
                "state": {k: (m.detach().cpu().clone(), v.detach().cpu().clone()) for k, (m, v) in self.state.items()},
                "groups": {e[0]: (e[2], e[3]) for e in self.entries}}

    def load_state_dict(self, sd) -> None:
        self.steps = int(sd["steps"])
        self.betas, self.eps = tuple(sd["betas"]), float(sd["eps"])
        for e in self.entries:
            m, v = sd["state"][e[0]]
            self.state[e[0]][0].copy_(m)
            self.state[e[0]][1].copy_(v)
            if e[0] in sd.get("groups", {}):
                e[2], e[3] = (float(x) for x in sd["groups"][e[0]])

    @torch.no_grad()
    def step(self) -> None:
        self.steps += 1
        n = len(self.entries)
        P = C.c_void_p
        params = (P * n)(*[e[1].data_ptr() for e in self.entries])
        grads = (P * n)(*[e[1].grad.data_ptr() for e in self.entries])
        m = (P * n)(*[self.state[e[0]][0].data_ptr() for e in self.entries])
        v = (P * n)(*[self.state[e[0]][1].data_ptr() for e in self.entries])
        numel = (C.c_int64 * n)(*[e[1].numel() for e in self.entries])
        lrs = (C.c_float * n)(*[e[2] for e in self.entries])
        wds = (C.c_float * n)(*[e[3] for e in self.entries])
        mdl = self.model
        with mdl._on_device():
            mdl._check(mdl._library().msclip_op_adamw(n, params, grads, m, v, numel, lrs, wds, float(self.betas[0]),
                                                     float(self.betas[1]), float(self.eps), int(self.steps), mdl._stream()),
                       "msclip_op_adamw")
        mdl.refresh_weights([e[0] for e in self.entries])


class CosineSchedule:
    """Learning-rate schedule named by the reference's configuration (experiments/model/b32.yaml:37-46: timm 'cosine' with
    warmup_epochs 5, warmup_lr 1e-6, min_lr 1e-5, cooldown_epochs 10) - the reference ships no loop that runs it, so this restates
    the published rule of timm's CosineLRScheduler (t_in_epochs, single cycle): linear warm-up from ``warmup_lr`` to the base rate
    over ``warmup`` epochs, then ``min_lr + 0.5 (base - min_lr)(1 + cos(pi t / t_initial))`` with t counted from epoch 0 as timm
    does (``warmup_prefix`` False), ``min_lr`` from ``t_initial`` on (cool-down epochs run at ``min_lr``).  Every parameter
    group keeps its own base rate (LR / LR_SHARE)."""

    def __init__(self, optimizer: AdamW, epochs: int, warmup_epochs: int = 5, warmup_lr: float = 1e-6, min_lr: float = 1e-5,
                 cooldown_epochs: int = 10):
        self.opt = optimizer
        self.t_initial = int(epochs)
        self.warmup, self.warmup_lr, self.min_lr, self.cooldown = int(warmup_epochs), float(warmup_lr), float(min_lr), int(cooldown_epochs)
        self.base = [e[2] for e in optimizer.entries]

    def total_epochs(self) -> int:
        return self.t_initial + self.cooldown

    def lr_at(self, epoch: float, base: float) -> float:
        import math
        if epoch < self.warmup:
            return self.warmup_lr + epoch * (base - self.warmup_lr) / self.warmup
        if epoch < self.t_initial:
            return self.min_lr + 0.5 * (base - self.min_lr) * (1.0 + math.cos(math.pi * epoch / self.t_initial))
        return self.min_lr

    def step(self, epoch: float) -> None:
        """Set every group's rate for ``epoch`` (fractional epochs allowed: per-iteration updates)."""
        for e, b in zip(self.opt.entries, self.base):
            e[2] = self.lr_at(epoch, b)
