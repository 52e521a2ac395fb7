"""CPU-side checks of the environment that runs the reference's unmodified tools (tests/ref_env, tests/ref_tool.py):
the yacs stand-in behaves like yacs where the reference relies on it, and the real tools/zero_shot.py gets through
its config system, model construction and checkpoint loading on this box - up to its hard-coded `.cuda()`
(tools/zero_shot.py:127), which is where a CPU-only box must stop."""
import os
import sys
import textwrap

import pytest

import ref_tool

sys.path.insert(0, ref_tool.REF_ENV)
from yacs.config import CfgNode as CN  # noqa: E402

sys.path.remove(ref_tool.REF_ENV)


def test_yacs_stand_in(tmp_path):
    c = CN()
    c.A = CN()
    c.A.X = 1
    c.A.LIST = [1, 2]
    c.OPEN = CN(new_allowed=True)
    c.NAME = ""
    y = tmp_path / "c.yaml"
    y.write_text(textwrap.dedent("""
        A:
          X: 3
        OPEN:
          NEW_FLAG: true
          NESTED:
            K: [1, 2]
    """))
    c.merge_from_file(str(y))
    assert c.A.X == 3 and c.OPEN.NEW_FLAG is True and c.OPEN.NESTED.K == [1, 2]
    assert getattr(c.OPEN, "MISSING", "dflt") == "dflt"          # the model reads CUSTOM flags this way
    c.merge_from_list(["A.X", "7", "NAME", "run1", "A.LIST", "[4, 5]"])
    assert c.A.X == 7 and c.NAME == "run1" and c.A.LIST == [4, 5]
    with pytest.raises(KeyError):
        c.merge_from_list(["A.NOPE", "1"])                       # what forbids passing LAYERS through `opts`
    bad = tmp_path / "bad.yaml"
    bad.write_text("A:\n  UNKNOWN: 1\n")
    with pytest.raises(KeyError):
        c.merge_from_file(str(bad))
    c.freeze()
    with pytest.raises(AttributeError):
        c.A.X = 9
    c.defrost()
    c.A.X = 9
    assert "X: 9" in c.dump() and c.clone().A.X == 9


def test_unmodified_tool_reaches_the_gpu_call_on_cpu(tmp_path):
    import torch
    root = ref_tool.reference_root()
    if root is None:
        pytest.skip("reference not present")
    if torch.cuda.is_available():
        pytest.skip("GPU box: tests/test_reference_tool_gpu.py runs the tool to completion")
    data = os.path.join(str(tmp_path), "data")
    ref_tool.make_image_folder(data, 2, 1)
    ckpt = os.path.join(str(tmp_path), "ckpt", "x", "model.pth")
    ref_tool.make_checkpoint(ckpt, 2)
    with pytest.raises(RuntimeError) as e:
        ref_tool.run_tool(root, str(tmp_path), ckpt, data, 2, dropin=True, dump=os.path.join(str(tmp_path), "d.npz"),
                          timeout=600)
    log = str(e.value)
    assert "=> load model file" in log and "Start to build zeroshot classifier" in log, log[-1500:]
    assert "tokenizer(texts).cuda()" in log                      # tools/zero_shot.py:127
