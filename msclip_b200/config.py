"""MS-CLIP-S configuration envelope.

The reference builds its model from a yacs ``CfgNode`` (``get_clip_model``,
lib/models/clip_openai_pe_res_v1.py:3182-3227) and ~90 ``CUSTOM.*`` flags read with
``getattr(cfg, FLAG, default)``.  Only the flag combination shipped in
experiments/model/b32-yfcc-msclips.yaml / b32-laion-msclips.yaml / b16-yfcc-msclips.yaml is live
(SURVEY.md Appendix B).  ``MSCLIPConfig`` is that envelope as plain data; ``from_reference_config``
accepts the reference's config object (or any duck-typed namespace / nested dict) and refuses
anything outside the envelope instead of silently mis-computing.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Any, List, Sequence

LATERAL_LAYERS = (2, 4, 6, 8, 10)


@dataclass(frozen=True)
class MSCLIPConfig:
    embed_dim: int = 512
    image_resolution: int = 224
    patch_size: int = 32            # 32 (B/32) or 16 (B/16)
    width: int = 768                # both towers (b32-yfcc-msclips.yaml:6-8)
    layers: int = 12                # "layers" counts the conv stem as vision layer 0
    context_length: int = 77
    vocab_size: int = 49408
    gather_tensors: bool = False
    # derived from patch size unless overridden (b16-yfcc-msclips.yaml:21-27,43)
    early_strides: Sequence[int] = field(default=())
    parallel_strides: Sequence[int] = field(default=())
    t2b_kernels: Sequence[int] = field(default=())

    def __post_init__(self):
        if self.patch_size not in (16, 32):
            raise ValueError("MS-CLIP-S envelope: patch_size must be 16 or 32")
        if self.width % 64 or self.width % 16:
            raise ValueError("width must be a multiple of 64 (head_dim is fixed at 64)")
        if self.layers < 1:
            raise ValueError("layers must be >= 1")
        p16 = self.patch_size == 16
        if not self.early_strides:
            object.__setattr__(self, "early_strides", (2, 2, 2, 1) if p16 else (2, 2, 2, 2))
        if not self.parallel_strides:
            object.__setattr__(self, "parallel_strides", (2, 2, 2, 2, 1) if p16 else (2, 2, 2, 2, 2))
        if not self.t2b_kernels:
            object.__setattr__(self, "t2b_kernels", (8, 4, 2, 1, 1) if p16 else (16, 8, 4, 2, 1))
        object.__setattr__(self, "early_strides", tuple(int(s) for s in self.early_strides))
        object.__setattr__(self, "parallel_strides", tuple(int(s) for s in self.parallel_strides))
        object.__setattr__(self, "t2b_kernels", tuple(int(s) for s in self.t2b_kernels))
        g = self.image_resolution // 2
        for s in self.early_strides:
            g //= s
        if g != self.grid:
            raise ValueError(f"early-conv strides give a {g}x{g} grid, patch size implies {self.grid}")
        # every lateral adapter must land exactly on the token grid
        r = self.image_resolution
        for j in range(5):
            r //= self.parallel_strides[j]
            if r // self.t2b_kernels[j] != self.grid or r % self.t2b_kernels[j]:
                raise ValueError(f"lateral adapter {j}: {r}/{self.t2b_kernels[j]} != grid {self.grid}")

    # ---- derived sizes -------------------------------------------------------------------------
    @property
    def heads(self) -> int:
        return self.width // 64

    @property
    def grid(self) -> int:
        return self.image_resolution // self.patch_size

    @property
    def image_tokens(self) -> int:          # L for the vision tower (CLS + grid^2)
        return self.grid * self.grid + 1

    @property
    def n_shared_blocks(self) -> int:       # vision blocks 1..layers-1 (vision block 0 is the stem)
        return self.layers - 1

    @property
    def branch_dims(self) -> List[int]:     # output channels of the 5 parallel-branch stages
        w = self.width
        return [w // 16, w // 8, w // 4, w // 2, w]

    def active_laterals(self) -> List[int]:
        """Indices j of the lateral adapters that actually run (block index < layers)."""
        return [j for j, idx in enumerate(LATERAL_LAYERS) if idx < self.layers]

    def to_dict(self) -> dict:
        return asdict(self)


def _get(obj: Any, name: str, default: Any = None) -> Any:
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


_DEAD_FLAGS = (
    "GUMBEL_SELECT", "GUMBEL_ADDTWO", "SHARE_BOTTOM_LAYER", "SAVE_GRADIENT", "GET_GRADIENT_FROMCKPT",
    "LORA_OPEN", "CONVIT_IN_V", "CVT_IN_V", "ADAPTER_FLAG", "PERCEIVER_IN_V", "PERCEIVER_IN_T",
    "PARALLEL_REUSE_EARLYCONV_FIRSTLAYER", "PARALLEL_REUSE_EARLYCONV_ALLLAYER", "PARALLEL_B2T",
    "PARALLEL_T2B_POOL_SIZE", "PRALLEL_T2B_ADD_BN_RELU", "PRALLEL_T2B_ADD_BN_LN_RELU",
    "PRALLEL_T2B_NOLN_ADD", "VISUAL_LAYER_MINUS1", "LOAD_SEARCHED_ARCH", "CONTAINER_IN_V",
    "OUTPUT_ATTN_RAW", "OUTPUT_LAST_LN", "LORA_INIT", "PARALLEL_T2B_WINDOWATTN",
)


def from_reference_config(config: Any) -> MSCLIPConfig:
    """Translate the object ``get_clip_model`` receives (M.py:3182-3225) into an MSCLIPConfig."""
    spec = _get(_get(config, "MODEL"), "SPEC")
    vis, txt = _get(spec, "VISION"), _get(spec, "TEXT")
    cu = _get(config, "CUSTOM")
    if _get(vis, "MODEL", "vit") != "vit":
        raise NotImplementedError("only the ViT vision tower is on the MS-CLIP-S path")
    for flag in _DEAD_FLAGS:
        if _get(cu, flag, False):
            raise NotImplementedError(f"CUSTOM.{flag} is outside the MS-CLIP-S envelope")
    required = {"CUSTOM_ATTN": True, "PARALLEL_IN_V": True, "PARALLEL_RESNET": True, "EARLY_CONV": True,
                "EARLY_CONV_NEW_IMPLEMENT": True, "EARLY_CONV_RES": True, "PRALLEL_T2B_USECLS": True}
    for flag, want in required.items():
        if bool(_get(cu, flag, False)) != want:
            raise NotImplementedError(f"CUSTOM.{flag} must be {want} for MS-CLIP-S")
    if list(_get(cu, "SHARE_MODULES", [])) != ["attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj", "mlp"]:
        raise NotImplementedError("CUSTOM.SHARE_MODULES must be the MS-CLIP-S list")
    if _get(cu, "N_LAYERS", -1) != 1:
        raise NotImplementedError("CUSTOM.N_LAYERS must be 1")
    if _get(cu, "PARALLEL_N_LAYERS", 0) != 5 or list(_get(cu, "PARALLEL_LATERAL_LAYER", [])) != list(LATERAL_LAYERS):
        raise NotImplementedError("parallel branch must have 5 stages at layers [2,4,6,8,10]")
    if list(_get(cu, "PARALLEL_RESNET_LAYERS", [])) != [0, 1, 1, 1, 1]:
        raise NotImplementedError("PARALLEL_RESNET_LAYERS must be [0,1,1,1,1]")
    if _get(cu, "EARLY_CONV_RES_FIRSTCONV_KERNEL", 3) != 3 or _get(cu, "EARLY_CONV_RES_BLOCK", "basic_v0") != "basic_v0":
        raise NotImplementedError("early-conv stem must be 3x3 / basic_v0")
    if list(_get(cu, "EARLY_CONV_RES_LAYERS", [1, 1, 1, 1])) != [1, 1, 1, 1]:
        raise NotImplementedError("EARLY_CONV_RES_LAYERS must be [1,1,1,1]")
    if list(_get(cu, "PARALLEL_KERNELS", [3] * 5)) != [3] * 5 or list(_get(cu, "PARALLEL_PADDINGS", [1] * 5)) != [1] * 5:
        raise NotImplementedError("parallel branch convs must be 3x3 pad 1")
    if list(_get(cu, "PRALLEL_T2B_PADDINGS", [0] * 5)) != [0] * 5:
        raise NotImplementedError("lateral adapter padding must be 0")
    k = list(_get(cu, "PRALLEL_T2B_KERNELS", []))
    if k != list(_get(cu, "PRALLEL_T2B_STRIDES", [])):
        raise NotImplementedError("lateral adapter kernel must equal its stride")
    width = int(_get(vis, "WIDTH"))
    if int(_get(txt, "WIDTH")) != width or int(_get(txt, "HEADS")) != width // 64:
        raise NotImplementedError("text tower must share the vision width/heads")
    if int(_get(txt, "LAYERS")) != int(_get(vis, "LAYERS")):
        raise NotImplementedError("both towers must have the same depth")
    if _get(txt, "STYLE", "clip") != "clip" or _get(txt, "TOKENIZER", "clip") != "clip":
        raise NotImplementedError("text tower style/tokenizer must be 'clip'")
    if _get(spec, "POOL_TYPE", "default") != "default" or _get(spec, "SKIP_CLS", False):
        raise NotImplementedError("only default (CLS / EOT) pooling is on the path")
    return MSCLIPConfig(
        embed_dim=int(_get(spec, "EMBED_DIM")),
        image_resolution=int(_get(_get(config, "TRAIN"), "IMAGE_SIZE")[0]),
        patch_size=int(_get(vis, "PATCH_SIZE")),
        width=width,
        layers=int(_get(vis, "LAYERS")),
        context_length=int(_get(txt, "CONTEXT_LENGTH")),
        vocab_size=int(_get(txt, "VOCAB_SIZE")),
        gather_tensors=bool(_get(spec, "GATHER_TENSORS", False)),
        early_strides=tuple(_get(cu, "EARLY_CONV_RES_STRIDES", ())),
        parallel_strides=tuple(_get(cu, "PARALLEL_STRIDES", ())),
        t2b_kernels=tuple(k),
    )
