"""Checkpoint formats of the reference.

The reference reads two layouts: a BARE state dict (released checkpoints, tools/zero_shot.py:223-224) and its TRAINING
format (lib/utils/utils.py:157-199: ``{'epoch' | 'step', 'model', 'state_dict', 'perf', 'optimizer', ...}``, with a
``module.`` prefix on every key when the model was wrapped by DistributedDataParallel, utils.py:172-173).  This module
loads either into the drop-in CLIP and writes the training format back, including the state of ``msclip_b200.optim.AdamW``
so that a run resumes bit-identically.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch


def extract_state_dict(obj: Any) -> Dict[str, torch.Tensor]:
    """Bare state dict or training-format checkpoint -> state dict with any DDP ``module.`` prefix removed."""
    sd = obj["state_dict"] if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict) else obj
    if not isinstance(sd, dict):
        raise TypeError("checkpoint holds no state dict")
    if sd and all(k.startswith("module.") for k in sd):
        sd = {k[len("module."):]: v for k, v in sd.items()}
    return sd


def load_checkpoint(model, path_or_obj, optimizer=None, strict: bool = True) -> Dict[str, Any]:
    """Load weights (and, from a training-format checkpoint, the optimiser state) and return the bookkeeping fields
    (``epoch`` / ``step`` = the NEXT epoch / step as the reference stores it, ``perf``, ``model``)."""
    obj = torch.load(path_or_obj, map_location="cpu") if isinstance(path_or_obj, (str, bytes)) or hasattr(path_or_obj, "read") \
        else path_or_obj
    model.load_state_dict(extract_state_dict(obj), strict=strict)
    if getattr(model, "_train_keys", None):          # a training handle keeps packed copies: refresh them from the new masters
        model.refresh_weights()
    meta = {k: obj[k] for k in ("epoch", "step", "model", "perf") if isinstance(obj, dict) and k in obj}
    if optimizer is not None and isinstance(obj, dict) and "optimizer" in obj:
        optimizer.load_state_dict(obj["optimizer"])
    return meta


def save_checkpoint(model, path, optimizer=None, *, epoch_or_step: int = 0, in_epoch: bool = True, best_perf: float = 0.0,
                    model_name: str = "clip_openai_pe_res_v1", distributed: bool = False) -> Dict[str, Any]:
    """Write the reference's training format (lib/utils/utils.py:174-199); returns the dict that was saved."""
    states = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    if distributed:
        states = {"module." + k: v for k, v in states.items()}
    save_dict = {("epoch" if in_epoch else "step"): epoch_or_step + 1, "model": model_name, "state_dict": states, "perf": best_perf}
    if optimizer is not None:
        save_dict["optimizer"] = optimizer.state_dict()
    torch.save(save_dict, path)
    return save_dict
