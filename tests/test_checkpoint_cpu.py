"""Checkpoint layouts of the reference on the CPU (no library call): bare state dict, training format, DDP prefix."""
import io

import torch

from msclip_b200.checkpoint import extract_state_dict, load_checkpoint
from msclip_b200.config import MSCLIPConfig
from msclip_b200.model import CLIP


def test_training_format_and_ddp_prefix_load_like_a_bare_state_dict():
    cfg = MSCLIPConfig(patch_size=32, layers=2)
    src = CLIP(cfg)
    fresh = {}          # aliased keys (text blocks >= 1 share the vision tensors, M.py:2786-2830) must carry the same values
    sd = {}
    for k, v in src.state_dict().items():
        if v.dtype == torch.float32:
            sd[k] = fresh.setdefault(v.data_ptr(), torch.randn_like(v))
        else:
            sd[k] = v
    ckpt = {"epoch": 7, "model": "clip_openai_pe_res_v1", "perf": 0.5, "optimizer": {},
            "state_dict": {"module." + k: v for k, v in sd.items()}}
    assert set(extract_state_dict(ckpt)) == set(sd) == set(extract_state_dict(sd))
    buf = io.BytesIO()
    torch.save(ckpt, buf)
    buf.seek(0)
    dst = CLIP(cfg)
    meta = load_checkpoint(dst, buf)
    assert meta == {"epoch": 7, "model": "clip_openai_pe_res_v1", "perf": 0.5}
    for k, v in dst.state_dict().items():
        assert torch.equal(v, sd[k]), k
