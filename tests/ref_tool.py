"""Harness that runs the reference's UNMODIFIED ``tools/zero_shot.py`` (tools/zero_shot.py:185-310) as a subprocess,
once with the reference's own model and once with the B200-native drop-in behind ``get_clip_model``.

TEST INFRASTRUCTURE.  The tool is executed from the staged copy of the reference (``baseline/_ref``, see
tools/stage_reference.py) or from ``/root/reference``; ``tests/ref_env`` supplies the absent third-party packages
and the start-up hook.  Inputs are synthetic: an ``ImageFolder`` of random PNGs and a checkpoint written with
``torch.save(reference_model.state_dict())``."""
from __future__ import annotations

import os
import re
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ENV = os.path.join(REPO, "tests", "ref_env")


def reference_root():
    from oracle import ref_shim
    return ref_shim.REFERENCE_ROOT if ref_shim.reference_available() else None


def make_image_folder(root: str, n_classes: int = 4, per_class: int = 8, seed: int = 0, size=(256, 240)) -> int:
    """DATASET.ROOT/val/<class>/<i>.png - what torchvision.datasets.ImageFolder (zero_shot.py:214-216) reads."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    for c in range(n_classes):
        d = os.path.join(root, "val", f"class_{c:03d}")
        os.makedirs(d, exist_ok=True)
        for i in range(per_class):
            # smooth random fields (low-frequency noise up-sampled) so bicubic resize + crop is well behaved
            small = rng.integers(0, 256, size=(size[1] // 16, size[0] // 16, 3), dtype=np.uint8)
            Image.fromarray(small).resize(size, Image.BILINEAR).save(os.path.join(d, f"{i:03d}.png"))
    return n_classes * per_class


def make_checkpoint(path: str, layers: int, patch: int = 32, seed: int = 0) -> None:
    """Random-init weights of the REAL reference model, saved the way released checkpoints are (bare state_dict,
    zero_shot.py:223-224).  BatchNorm statistics / affine terms and biases are randomised so that every folded
    constant matters."""
    import torch
    from msclip_b200.config import MSCLIPConfig
    from oracle import ref_shim
    cfg = MSCLIPConfig(layers=layers, patch_size=patch)
    torch.manual_seed(seed)
    model = ref_shim.build_reference_model(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, t in model.state_dict().items():
            if k.endswith("running_mean"):
                t.copy_(0.1 * torch.randn(t.shape, generator=g))
            elif k.endswith("running_var"):
                t.copy_(1.0 + 0.5 * torch.rand(t.shape, generator=g))
            elif k.endswith(".bias") and t.dim() == 1:
                t.copy_(0.02 * torch.randn(t.shape, generator=g))
        model.logit_scale.fill_(float(np.log(1 / 0.07)))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(model.state_dict(), path)


def run_tool(ref_root: str, work: str, ckpt: str, data_root: str, layers: int, dropin: bool, dump: str,
             model_yaml: str = "experiments/model/b32-yfcc-msclips.yaml", timeout: int = 3000, extra_env=None):
    """python tools/zero_shot.py --ds ... --model ... <opts>, cwd = reference root; returns (top1 %, log text)."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([REF_ENV, REPO] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env["MSCLIP_TOOL_DUMP"] = dump
    env["MSCLIP_DROPIN"] = "1" if dropin else "0"
    env["MSCLIP_DROPIN_TOKENIZER"] = "1" if dropin else "0"   # the drop-in run also tokenises with the native tokenizer
    env.update(extra_env or {})
    # user-side configuration files (the tool itself and the shipped YAMLs stay untouched): the shipped model config
    # as BASE plus the checkpoint path / depth, and the ImageNet dataset config pointed at the synthetic folder
    os.makedirs(work, exist_ok=True)
    model_cfg = os.path.join(work, "model.yaml")
    with open(model_cfg, "w") as f:
        f.write(f"BASE: ['{os.path.join(ref_root, model_yaml)}']\nOUTPUT_DIR: '{os.path.join(work, 'out')}'\n"
                f"MODEL:\n  PRETRAINED_MODEL: '{ckpt}'\n  SPEC:\n    VISION:\n      LAYERS: {layers}\n"
                f"    TEXT:\n      LAYERS: {layers}\n")
    ds_cfg = os.path.join(work, "dataset.yaml")
    with open(ds_cfg, "w") as f:
        f.write(f"BASE: ['{os.path.join(ref_root, 'experiments/dataset/imagenet.yaml')}']\nDATASET:\n  ROOT: '{data_root}'\n")
    cmd = [sys.executable, os.path.join("tools", "zero_shot.py"), "--ds", ds_cfg, "--model", model_cfg]
    r = subprocess.run(cmd, cwd=ref_root, env=env, capture_output=True, text=True, timeout=timeout)
    log = r.stdout + "\n" + r.stderr
    if r.returncode != 0:
        raise RuntimeError(f"tools/zero_shot.py failed ({r.returncode}):\n{log[-4000:]}")
    m = re.search(r"accuracy@1\s+([0-9.]+)%", log)
    if not m:
        raise RuntimeError("no accuracy line in the tool's log:\n" + log[-2000:])
    return float(m.group(1)), log


def load_dump(path: str):
    z = np.load(path)
    return z["image"], z["text_class"], str(z["kind"])
