"""Drop-in for ``dataset.languages.SimpleTokenizer`` (lib/dataset/languages/simple_tokenizer.py of the reference) backed by
the native tokenizer of libmsclip_b200 (csrc/tokenizer.cu): same constructor argument (the path of the reference's own
``bpe_simple_vocab_16e6.txt.gz`` - the file is not shipped here), same ``encode`` / ``tokenize`` / ``get_*`` methods, batches
tokenised on all host cores.  ``basic_clean`` (ftfy.fix_text + html.unescape, simple_tokenizer.py:53-56) is third-party text
repair and stays Python: ftfy is used when it is installed, exactly like the reference; without it only html.unescape runs.
"""
from __future__ import annotations

import ctypes as C
import gzip
import html
from typing import List, Union

import torch

from . import _lib

try:                                    # the reference imports ftfy unconditionally (simple_tokenizer.py:6)
    import ftfy as _ftfy
except ImportError:                     # pragma: no cover - not installed in this image
    _ftfy = None


def basic_clean(text: str) -> str:
    if _ftfy is not None:
        text = _ftfy.fix_text(text)
    text = html.unescape(html.unescape(text))
    return text.strip()


class SimpleTokenizer:
    def __init__(self, bpe_path: str):
        self._lib = _lib.lib()
        with gzip.open(bpe_path) as f:
            merges = f.read()
        self._h = C.c_void_p()
        _lib.check(self._lib.msclip_tokenizer_create(merges, len(merges), C.byref(self._h)), "msclip_tokenizer_create")
        v, s, e = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self._lib.msclip_tokenizer_info(self._h, C.byref(v), C.byref(s), C.byref(e)), "msclip_tokenizer_info")
        self._vocab, self._sot, self._eot = v.value, s.value, e.value

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.msclip_tokenizer_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- reference API (simple_tokenizer.py:125-166) ---------------------------------------------------------------
    def encode(self, text: str) -> List[int]:
        raw = basic_clean(text).encode("utf-8")
        cap = 4 * len(raw) + 8
        buf = (C.c_int32 * cap)()
        n = self._lib.msclip_tokenizer_encode(self._h, raw, len(raw), buf, cap)
        if n < 0:
            raise _lib.MsclipError("msclip_tokenizer_encode failed")
        return list(buf[:n])

    def tokenize(self, texts: Union[str, List[str]], context_length: int = 77) -> torch.Tensor:
        if isinstance(texts, str):
            texts = [texts]
        raws = [basic_clean(t).encode("utf-8") for t in texts]
        offs = [0]
        for r in raws:
            offs.append(offs[-1] + len(r))
        blob = b"".join(raws)
        out = torch.zeros(len(raws), context_length, dtype=torch.long)
        _lib.check(self._lib.msclip_tokenizer_tokenize(self._h, blob, (C.c_int64 * len(offs))(*offs), len(raws), int(context_length),
                                                       C.c_void_p(out.data_ptr()), 0), "msclip_tokenizer_tokenize")
        return out

    __call__ = tokenize

    def get_vocab_size(self) -> int:
        return 49408

    def get_eot_token(self) -> int:
        return self._eot

    def get_sot_token(self) -> int:
        return self._sot

    def check_added_tokens(self) -> int:
        return 0

    def get_tokenizer_obj(self):
        return None
