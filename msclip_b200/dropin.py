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
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys

TARGET = "models.clip_openai_pe_res_v1"
_installed = False


def patch_module(module) -> None:
    """Replace ``module.get_clip_model`` (M.py:3182) by the B200-native builder; keep the original."""
    from .model import get_clip_model
    if getattr(module, "_msclip_b200_patched", False):
        return
    module._reference_get_clip_model = getattr(module, "get_clip_model", None)
    module.get_clip_model = get_clip_model
    module._msclip_b200_patched = True


class _PatchingLoader(importlib.abc.Loader):
    def __init__(self, wrapped):
        self._wrapped = wrapped

    def create_module(self, spec):
        return self._wrapped.create_module(spec)

    def exec_module(self, module):
        self._wrapped.exec_module(module)
        patch_module(module)


class _Finder(importlib.abc.MetaPathFinder):
    _busy = False

    def find_spec(self, fullname, path, target=None):
        if fullname != TARGET or self._busy:
            return None
        self._busy = True          # other wrapping finders on sys.meta_path delegate back to us: answer only once
        try:
            for finder in sys.meta_path:
                if finder is self or not hasattr(finder, "find_spec"):
                    continue
                spec = finder.find_spec(fullname, path, target)
                if spec is not None and spec.loader is not None:
                    spec.loader = _PatchingLoader(spec.loader)
                    return spec
            return None
        finally:
            self._busy = False


def install() -> None:
    """Idempotent.  Patches the module now if it is already imported, otherwise on import."""
    global _installed
    if TARGET in sys.modules:
        patch_module(sys.modules[TARGET])
    if not _installed:
        sys.meta_path.insert(0, _Finder())
        _installed = True
