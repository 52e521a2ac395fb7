"""The post-import hook that swaps the reference's get_clip_model (tools/zero_shot.py:220) for ours."""
import os
import sys
import textwrap

from msclip_b200 import dropin
from msclip_b200.config import MSCLIPConfig


def test_hook_patches_reference_module_on_import(tmp_path, monkeypatch):
    pkg = tmp_path / "models"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    (pkg / "clip_openai_pe_res_v1.py").write_text(textwrap.dedent("""
        def get_clip_model(config, vocab_size=None, eot_token=None, **kwargs):
            return "reference model"
    """))
    monkeypatch.syspath_prepend(str(tmp_path))          # what tools/_init_paths.py does with lib/
    for name in ("models", dropin.TARGET):
        sys.modules.pop(name, None)
    dropin.install()
    try:
        from models import clip_openai_pe_res_v1 as M
        assert M._msclip_b200_patched and M._reference_get_clip_model(None) == "reference model"
        model = M.get_clip_model(MSCLIPConfig(layers=2))
        assert type(model).__name__ == "CLIP" and hasattr(model, "encode_image") and hasattr(model, "encode_text")
        assert "visual.transformer.resblocks.0.conv1.weight" in model.state_dict()
        dropin.install()                                 # idempotent
        assert M._reference_get_clip_model(None) == "reference model"
    finally:
        for name in ("models", dropin.TARGET):
            sys.modules.pop(name, None)
