"""Swap the reference's model builder for the B200-native one without touching ``tools/``.

``tools/zero_shot.py`` does ``import _init_paths`` (which force-inserts ``lib/`` at ``sys.path[0]``,
tools/_init_paths.py:9-17) and then ``from models import clip_openai_pe_res_v1`` and calls
``clip_openai_pe_res_v1.get_clip_model(config)`` (tools/zero_shot.py:40, 220), so PYTHONPATH precedence
cannot replace the module.  ``install()`` registers a post-import hook instead: as soon as
``models.clip_openai_pe_res_v1`` has been imported, its ``get_clip_model`` attribute is replaced by
``msclip_b200.model.get_clip_model`` (same signature, same state-dict keys, same methods).  The hook is
inherited by the subprocess that ``tools/eval_zeroshot.py`` spawns when it is installed from a
``sitecustomize.py`` on PYTHONPATH:

    # sitecustomize.py
    import msclip_b200.dropin; msclip_b200.dropin.install()

``install(tokenizer=True)`` (or MSCLIP_DROPIN_TOKENIZER=1) additionally swaps ``dataset.languages.SimpleTokenizer``
(tools/zero_shot.py:34, 232) for the native tokenizer of the library (msclip_b200.tokenizer; identical ids, the reference's
own merges file); methods the zero-shot path never calls (decode, *_with_idx) fall through to the reference's Python class.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys

import os

TARGET = "models.clip_openai_pe_res_v1"
TOKENIZER_TARGET = "dataset.languages.simple_tokenizer"
_installed = False
_patchers = {}


def patch_module(module) -> None:
    """Replace ``module.get_clip_model`` (M.py:3182) by the B200-native builder; keep the original."""
    from .model import get_clip_model
    if getattr(module, "_msclip_b200_patched", False):
        return
    module._reference_get_clip_model = getattr(module, "get_clip_model", None)
    module.get_clip_model = get_clip_model
    module._msclip_b200_patched = True


def patch_tokenizer(module) -> None:
    """Replace ``module.SimpleTokenizer`` (simple_tokenizer.py:64) by the native tokenizer with the same default merges file."""
    from .tokenizer import SimpleTokenizer as Native
    if getattr(module, "_msclip_b200_patched", False):
        return
    reference_cls = module.SimpleTokenizer
    default_path = module.default_bpe()

    class SimpleTokenizer(Native):
        def __init__(self, bpe_path: str = default_path):
            super().__init__(bpe_path)
            self._bpe_path = bpe_path

        def __getattr__(self, name):                      # decode / encode_with_idx / ...: the reference's own implementation
            if name.startswith("_"):
                raise AttributeError(name)
            ref = self.__dict__.get("_ref")
            if ref is None:
                ref = self.__dict__["_ref"] = reference_cls(self.__dict__.get("_bpe_path", default_path))
            return getattr(ref, name)

    module._reference_SimpleTokenizer = reference_cls
    module.SimpleTokenizer = SimpleTokenizer
    module._msclip_b200_patched = True


class _PatchingLoader(importlib.abc.Loader):
    def __init__(self, wrapped, patch):
        self._wrapped = wrapped
        self._patch = patch

    def create_module(self, spec):
        return self._wrapped.create_module(spec)

    def exec_module(self, module):
        self._wrapped.exec_module(module)
        self._patch(module)


class _Finder(importlib.abc.MetaPathFinder):
    _busy = False

    def find_spec(self, fullname, path, target=None):
        if fullname not in _patchers or self._busy:
            return None
        self._busy = True          # other wrapping finders on sys.meta_path delegate back to us: answer only once
        try:
            for finder in sys.meta_path:
                if finder is self or not hasattr(finder, "find_spec"):
                    continue
                spec = finder.find_spec(fullname, path, target)
                if spec is not None and spec.loader is not None:
                    spec.loader = _PatchingLoader(spec.loader, _patchers[fullname])
                    return spec
            return None
        finally:
            self._busy = False


def install(tokenizer=None) -> None:
    """Idempotent.  Patches the module(s) now if already imported, otherwise on import."""
    global _installed
    _patchers[TARGET] = patch_module
    if tokenizer is None:
        tokenizer = os.environ.get("MSCLIP_DROPIN_TOKENIZER") == "1"
    if tokenizer:
        _patchers[TOKENIZER_TARGET] = patch_tokenizer
    for name, patch in _patchers.items():
        if name in sys.modules:
            patch(sys.modules[name])
    if not _installed:
        sys.meta_path.insert(0, _Finder())
        _installed = True
