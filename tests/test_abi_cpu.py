"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU driver, exports every
symbol the public headers declare, mirrors the reference's state-dict contract, validates the MS-CLIP-S
envelope, and fails loudly (no fallback) when asked to compute without a GPU."""
import ctypes as C
import json
import os
import re

import pytest
import torch

from msclip_b200 import _lib, synth
from msclip_b200.config import MSCLIPConfig, from_reference_config
from golden_util import GOLDEN_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("msclip_b200.h", "msclip_b200_ops.h"):
        with open(os.path.join(ROOT, "include", hdr)) as f:
            text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
        names |= set(re.findall(r"\b(msclip_[a-z0-9_]+)\s*\(", text))
    return names


def make_handle(**kw):
    cfg = MSCLIPConfig(**kw)
    cc = _lib.Config(cfg.patch_size, cfg.layers, cfg.width, cfg.embed_dim, cfg.image_resolution, cfg.context_length,
                     cfg.vocab_size, (C.c_int32 * 4)(*cfg.early_strides), (C.c_int32 * 5)(*cfg.parallel_strides),
                     (C.c_int32 * 5)(*cfg.t2b_kernels))
    h = C.c_void_p()
    rc = _lib.lib().msclip_create(C.byref(cc), C.byref(h))
    return rc, h, cc


def test_library_exports_every_declared_symbol():
    decl = declared_symbols()
    assert len(decl) >= 28
    lib = _lib.lib()
    for name in decl:
        assert hasattr(lib, name), f"{name} is declared in include/ but not exported"
    assert decl == set(_lib.exported_symbols()), decl ^ set(_lib.exported_symbols())
    assert b"sm_100a" in lib.msclip_version() and b"bf16" in lib.msclip_version()
    lib16 = _lib.lib("fp16")                                   # the fp16-operand build exports the same ABI
    for name in decl:
        assert hasattr(lib16, name), name
    assert b"fp16" in lib16.msclip_version()


@pytest.mark.parametrize("tag", ["b32_l2", "b32_l12", "b16_l3", "b16_l12"])
def test_c_side_state_dict_contract_equals_reference(tag):
    p, l = tag.split("_")
    rc, h, _ = make_handle(patch_size=int(p[1:]), layers=int(l[1:]))
    assert rc == 0
    lib = _lib.lib()
    with open(os.path.join(GOLDEN_DIR, f"state_dict_keys_{tag}.json")) as f:
        ref = json.load(f)
    n = lib.msclip_num_keys(h)
    got = {}
    for i in range(n):
        key, nd, shape = C.c_char_p(), C.c_int(), (C.c_int64 * 4)()
        assert lib.msclip_key_info(h, i, C.byref(key), C.byref(nd), shape) == 0
        got[key.value.decode()] = list(shape[: nd.value])
    assert got == ref
    lib.msclip_destroy(h)


def test_create_validates_the_envelope():
    lib = _lib.lib()
    rc, h, cc = make_handle()
    assert rc == 0
    lib.msclip_destroy(h)
    for field, value, msg in [("patch_size", 14, "patch_size"), ("width", 512, "width"), ("embed_dim", 256, "embed_dim"),
                              ("image_resolution", 336, "image_resolution")]:
        bad = _lib.Config.from_buffer_copy(cc)
        setattr(bad, field, value)
        h2 = C.c_void_p()
        assert lib.msclip_create(C.byref(bad), C.byref(h2)) != 0
        assert msg in _lib.last_error()
    bad = _lib.Config.from_buffer_copy(cc)
    bad.t2b_kernels[0] = 7
    assert lib.msclip_create(C.byref(bad), C.byref(C.c_void_p())) != 0 and "token grid" in _lib.last_error()


def test_set_weight_is_strict():
    lib = _lib.lib()
    rc, h, _ = make_handle(layers=2)
    shape = (C.c_int64 * 1)(768)
    assert lib.msclip_set_weight(h, b"no.such.key", None, _lib.F32, 1, shape) != 0
    assert "unexpected state-dict key" in _lib.last_error()
    shape = (C.c_int64 * 1)(100)
    assert lib.msclip_set_weight(h, b"ln_final.weight", None, _lib.F32, 1, shape) != 0
    assert "size mismatch" in _lib.last_error()
    assert lib.msclip_finalize_weights(h, None) != 0 and "missing state-dict key" in _lib.last_error()
    out = (C.c_float * 512)()
    assert lib.msclip_encode_text(h, None, 1, out, 1, None) != 0 and "not finalized" in _lib.last_error()
    lib.msclip_destroy(h)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly instead of computing somewhere else."""
    from msclip_b200.model import CLIP
    assert _lib.device_count() == 0
    model = CLIP(MSCLIPConfig(layers=2))
    with pytest.raises(_lib.MsclipError, match="no CPU fallback"):
        model.encode_text(torch.zeros(1, 77, dtype=torch.long))
    rc, h, _ = make_handle(layers=2)
    w = torch.ones(768)
    shape = (C.c_int64 * 1)(768)
    assert _lib.lib().msclip_set_weight(h, b"ln_final.weight", C.c_void_p(w.data_ptr()), _lib.F32, 1, shape) != 0
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    o = torch.zeros(8, 8)
    assert _lib.lib().msclip_op_gemm(C.c_void_p(a.data_ptr()), 8, C.c_void_p(a.data_ptr()), 8, 8, 8, 8, 1.0, None,
                                     C.c_void_p(o.data_ptr()), 8, None, 0, _lib.EPI_F32, None) != 0
    _lib.lib().msclip_destroy(h)


def test_product_path_does_not_import_the_oracle():
    for fn in os.listdir(os.path.join(ROOT, "msclip_b200")):
        if fn.endswith(".py"):
            with open(os.path.join(ROOT, "msclip_b200", fn)) as f:
                src = f.read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_python_module_mirrors_reference_state_dict_and_aliases():
    from msclip_b200.model import CLIP
    cfg = MSCLIPConfig(layers=3)
    m = CLIP(cfg)
    sd = m.state_dict()
    with open(os.path.join(GOLDEN_DIR, "state_dict_keys_b32_l3.json")) as f:
        ref = json.load(f)
    assert {k: list(v.shape) for k, v in sd.items()} == ref
    for k in sd:
        src = synth.alias_of(cfg, k)
        if src is not None:
            assert sd[k].data_ptr() == sd[src].data_ptr(), k       # same Parameter object (M.py:2808-2830)
    weights = synth.synth_state_dict(cfg, seed=3)
    res = m.load_state_dict({k: torch.as_tensor(v) for k, v in weights.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert float(m.logit_scale) == 1.0 and m.dtype == torch.float32


def test_config_translation_rejects_flags_outside_the_envelope():
    ns = lambda **k: type("N", (), k)()
    cu = dict(CUSTOM_ATTN=True, SHARE_MODULES=["attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj", "mlp"],
              PARALLEL_IN_V=True, PARALLEL_N_LAYERS=5, PARALLEL_LATERAL_LAYER=[2, 4, 6, 8, 10],
              PRALLEL_T2B_KERNELS=[8, 4, 2, 1, 1], PRALLEL_T2B_PADDINGS=[0] * 5, PRALLEL_T2B_STRIDES=[8, 4, 2, 1, 1],
              PRALLEL_T2B_USECLS=True, PARALLEL_RESNET=True, PARALLEL_RESNET_LAYERS=[0, 1, 1, 1, 1], EARLY_CONV=True,
              EARLY_CONV_NEW_IMPLEMENT=True, N_LAYERS=1, EARLY_CONV_RES=True, PARALLEL_STRIDES=[2, 2, 2, 2, 1],
              EARLY_CONV_RES_STRIDES=[2, 2, 2, 1])

    def cfg(**over):
        c = dict(cu)
        c.update(over)
        return ns(MODEL=ns(SPEC=ns(EMBED_DIM=512, VISION=ns(MODEL="vit", PATCH_SIZE=16, WIDTH=768, LAYERS=12),
                                   TEXT=ns(CONTEXT_LENGTH=77, VOCAB_SIZE=49408, WIDTH=768, HEADS=12, LAYERS=12,
                                           STYLE="clip", TOKENIZER="clip"))),
                  TRAIN=ns(IMAGE_SIZE=[224, 224]), CUSTOM=ns(**c), OUTPUT_DIR=".")
    ok = from_reference_config(cfg())
    assert ok.patch_size == 16 and ok.grid == 14 and ok.image_tokens == 197 and ok.early_strides == (2, 2, 2, 1)
    for over in (dict(LORA_OPEN=True), dict(GUMBEL_SELECT=True), dict(N_LAYERS=2), dict(PARALLEL_IN_V=False)):
        with pytest.raises(NotImplementedError):
            from_reference_config(cfg(**over))
