"""Process start-up hook for running the reference's *unmodified* tools in this image (test infrastructure).

Put ``tests/ref_env`` on PYTHONPATH (the reference's tools spawn / are spawned as subprocesses, which inherit it):

* always: re-create ``torch.nn.modules.linear._LinearWithBias`` (removed in torch 1.9, imported at M.py:17);
* ``MSCLIP_DROPIN=1``: install ``msclip_b200.dropin`` so ``clip_openai_pe_res_v1.get_clip_model`` builds the
  B200-native model - without it the reference's own model runs;
* ``MSCLIP_TOOL_DUMP=<file.npz>``: record what the tool computes with the model it was given, whichever it is:
  every ``encode_image`` output and, per ``encode_text`` call, the renormalised class mean the tool derives from it
  (tools/zero_shot.py:128-130).  Written at interpreter exit.
"""
import atexit
import importlib.abc
import os
import sys

TARGET = "models.clip_openai_pe_res_v1"


def _restore_linear_with_bias():
    import torch

    lin = torch.nn.modules.linear
    if not hasattr(lin, "_LinearWithBias"):
        class _LinearWithBias(torch.nn.Linear):
            def __init__(self, in_features, out_features):
                super().__init__(in_features, out_features, bias=True)

        lin._LinearWithBias = _LinearWithBias


def _record(module, path):
    import numpy as np
    import torch

    rec = {"image": [], "text_class": [], "kind": ""}
    builder = module.get_clip_model

    def get_clip_model(*args, **kwargs):
        model = builder(*args, **kwargs)
        rec["kind"] = type(model).__module__ + "." + type(model).__name__
        enc_i, enc_t = model.encode_image, model.encode_text

        def encode_image(*a, **k):
            out = enc_i(*a, **k)
            rec["image"].append(out.detach().float().cpu().numpy())
            return out

        def encode_text(*a, **k):
            out = enc_t(*a, **k)
            m = out.detach().float().mean(dim=0)
            rec["text_class"].append((m / m.norm()).cpu().numpy())
            return out

        model.encode_image, model.encode_text = encode_image, encode_text
        return model

    module.get_clip_model = get_clip_model

    def flush():
        if rec["image"] or rec["text_class"]:
            np.savez(path, image=np.concatenate(rec["image"]) if rec["image"] else np.zeros((0, 0), np.float32),
                     text_class=np.stack(rec["text_class"]) if rec["text_class"] else np.zeros((0, 0), np.float32),
                     kind=np.array(rec["kind"]))

    atexit.register(flush)


class _Loader(importlib.abc.Loader):
    def __init__(self, wrapped):
        self._wrapped = wrapped

    def create_module(self, spec):
        return self._wrapped.create_module(spec)

    def exec_module(self, module):
        _restore_linear_with_bias()
        self._wrapped.exec_module(module)
        dump = os.environ.get("MSCLIP_TOOL_DUMP")
        if dump:
            _record(module, dump)


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
                    spec.loader = _Loader(spec.loader)
                    return spec
            return None
        finally:
            self._busy = False


if os.environ.get("MSCLIP_DROPIN") == "1":
    import msclip_b200.dropin

    msclip_b200.dropin.install()
sys.meta_path.insert(0, _Finder())
