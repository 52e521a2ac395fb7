"""Import the *real* reference (Hxyou/MSCLIP, /root/reference) in the authoring container.

TEST INFRASTRUCTURE — used by oracle/make_golden.py and by container-only tests.  The reference
needs four shims under torch 2.11 / this image:
  1. ``timm.models.layers`` (DropPath, trunc_normal_)   — imported at M.py:22
  2. ``matplotlib.pyplot`` exporting ``get``            — imported at M.py:5
  3. ``torch.nn.modules.linear._LinearWithBias``        — imported at M.py:17 (removed in torch 1.9)
  4. yacs is absent: ``CLIP(...)`` is built directly with the ctor arguments ``get_clip_model``
     derives (M.py:3214-3225) and a CUSTOM namespace that raises AttributeError on missing flags.
Nothing here is copied from the reference; it is executed where it lies.
"""
from __future__ import annotations

import logging
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root() -> str:
    """/root/reference in the authoring container; on the GPU box the byte-identical copy that
    tools/stage_reference.py placed in the git-ignored baseline/_ref/ (it travels with the working tree)."""
    cands = [os.environ.get("MSCLIP_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "lib", "models", "clip_openai_pe_res_v1.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "models", "clip_openai_pe_res_v1.py"))


class CustomNamespace:
    """Attribute bag whose missing attributes raise (so ``getattr(cfg, FLAG, default)`` works)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def custom_flags(cfg) -> CustomNamespace:
    """CUSTOM.* of b32-yfcc-msclips.yaml / b16-yfcc-msclips.yaml."""
    kw = dict(
        CUSTOM_ATTN=True,
        SHARE_MODULES=["attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj", "mlp"],
        LR_SHARE=1e-4, WD_SHARE=0.2,
        PARALLEL_IN_V=True, PARALLEL_N_LAYERS=5, PARALLEL_LATERAL_LAYER=[2, 4, 6, 8, 10],
        PRALLEL_T2B_KERNELS=list(cfg.t2b_kernels), PRALLEL_T2B_PADDINGS=[0] * 5,
        PRALLEL_T2B_STRIDES=list(cfg.t2b_kernels), PRALLEL_T2B_USECLS=True,
        PARALLEL_RESNET=True, PARALLEL_RESNET_LAYERS=[0, 1, 1, 1, 1],
        EARLY_CONV=True, EARLY_CONV_NEW_IMPLEMENT=True, N_LAYERS=1, VISUAL_LAYER_MINUS1=False,
        EARLY_CONV_RES=True, EARLY_CONV_RES_FIRSTCONV_KERNEL=3, EARLY_CONV_RES_BLOCK="basic_v0",
        EARLY_CONV_RES_LAYERS=[1, 1, 1, 1],
    )
    if cfg.patch_size == 16:
        kw.update(PARALLEL_KERNELS=[3] * 5, PARALLEL_PADDINGS=[1] * 5,
                  PARALLEL_STRIDES=list(cfg.parallel_strides),
                  EARLY_CONV_RES_STRIDES=list(cfg.early_strides))
    return CustomNamespace(**kw)


_MODULE = None


def import_reference():
    """Return the reference module ``models.clip_openai_pe_res_v1`` (imported once)."""
    global _MODULE
    if _MODULE is not None:
        return _MODULE
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    import torch
    import transformers  # noqa: F401  (must be imported before the timm stub exists)

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm_layers = types.ModuleType("timm.models.layers")

        class DropPath(torch.nn.Module):
            def __init__(self, p=0.0):
                super().__init__()
                assert p == 0.0
            def forward(self, x):
                return x

        timm_layers.DropPath = DropPath
        timm_layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models, timm_models.layers = timm_models, timm_layers
        sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.get = lambda *a, **k: None
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    lin = torch.nn.modules.linear
    if not hasattr(lin, "_LinearWithBias"):
        class _LinearWithBias(torch.nn.Linear):
            def __init__(self, in_features, out_features):
                super().__init__(in_features, out_features, bias=True)
        lin._LinearWithBias = _LinearWithBias
    lib = os.path.join(REFERENCE_ROOT, "lib")
    if lib not in sys.path:
        sys.path.insert(0, lib)
    prev = logging.root.manager.disable
    logging.disable(logging.CRITICAL)
    try:
        from models import clip_openai_pe_res_v1 as M
    finally:
        logging.disable(prev)
    _MODULE = M
    return M


def build_reference_model(cfg, state_dict=None, gather_tensors=False):
    """Reference ``CLIP`` in eval mode, fp32, optionally loaded (strict) with ``state_dict``."""
    import torch
    M = import_reference()
    prev = logging.root.manager.disable
    logging.disable(logging.CRITICAL)
    try:
        model = M.CLIP(
            embed_dim=cfg.embed_dim, image_resolution=cfg.image_resolution, vision_layers=cfg.layers,
            vision_width=cfg.width, vision_patch_size=cfg.patch_size, vision_drop_path=0.0,
            context_length=cfg.context_length, vocab_size=cfg.vocab_size, transformer_width=cfg.width,
            transformer_heads=cfg.heads, transformer_layers=cfg.layers, transformer_style="clip",
            gather_tensors=gather_tensors, tokenizer_style="clip", pool_type="default", skip_cls=False,
            custom_config=custom_flags(cfg), output_dir=".")
    finally:
        logging.disable(prev)
    if state_dict is not None:
        sd = {k: torch.as_tensor(v) for k, v in state_dict.items()}
        model.load_state_dict(sd, strict=True)
    return model.eval()
