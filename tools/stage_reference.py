"""Stage the unmodified reference next to the repo so it travels to the GPU box.

    python tools/stage_reference.py [--src /root/reference] [--dst baseline/_ref]

`gpurun` (and the round-end driver) ship the working tree but not `/root/reference`.  This copies the reference's
`lib/`, `tools/` and `experiments/` (1.8 MB of Python / YAML, byte for byte) into the git-ignored `baseline/_ref/`,
which DOES travel, so that on the GPU box (i) the real `tools/zero_shot.py` can be run on top of the drop-in
(tests/test_reference_tool_gpu.py), (ii) `bench.py --impl reference` times the real `CLIP.forward`, and (iii) the
eager-PyTorch-on-B200 comparators run the real module.  Nothing under `baseline/_ref/` is product source and nothing
under `msclip_b200/` reads it; `__graft_entry__.build()` calls this when `/root/reference` is present."""
from __future__ import annotations

import argparse
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARTS = ("lib", "tools", "experiments")


def stage(src: str = "/root/reference", dst: str = os.path.join(ROOT, "baseline", "_ref")) -> bool:
    """Copy PARTS of the reference tree; returns False when there is no reference to copy."""
    if not os.path.isfile(os.path.join(src, "lib", "models", "clip_openai_pe_res_v1.py")):
        return False
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    for part in PARTS:
        s, d = os.path.join(src, part), os.path.join(dst, part)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d, ignore=ignore)
    # byte-for-byte check of the two files everything else hangs on
    for rel in (os.path.join("tools", "zero_shot.py"), os.path.join("lib", "models", "clip_openai_pe_res_v1.py")):
        if not filecmp.cmp(os.path.join(src, rel), os.path.join(dst, rel), shallow=False):
            raise RuntimeError(f"staged copy of {rel} differs from the reference")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=os.path.join(ROOT, "baseline", "_ref"))
    a = ap.parse_args()
    ok = stage(a.src, a.dst)
    print(("staged " + a.src + " -> " + a.dst) if ok else ("no reference under " + a.src))
    sys.exit(0 if ok else 1)
