"""Host-side mirror of the reference model API, backed by the sm_100a library.

Mirrors lib/models/clip_openai_pe_res_v1.py ("M.py") of Hxyou/MSCLIP for the MS-CLIP-S path:

* ``get_clip_model(config, vocab_size=None, eot_token=None)``           M.py:3182-3227
* ``CLIP.encode_image(image, norm=True, action=None)``                 M.py:2979-2985
* ``CLIP.encode_text(text, norm=True, action=None)``                   M.py:3043-3079
* ``CLIP.forward(image, text) -> logits``                              M.py:3126-3155
* ``CLIP.state_dict() / load_state_dict()`` with the reference's key names, including the text-tower
  keys that alias the vision tower's shared attention / MLP parameters (M.py:2786-2830)
* ``CLIP.logit_scale``, ``CLIP.dtype``                                 M.py:2850, 2973-2975

plus the call the reference lacks: ``contrastive_loss(image, text)`` — the symmetric cross-entropy over
the global batch, computed by one fused kernel that never materialises the logits.

PyTorch is plumbing here (parameter storage, device memory, streams); all arithmetic happens in
libmsclip_b200.so through ctypes.  There is no eager / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Optional

import torch
from torch import nn

from . import _lib
from .config import MSCLIPConfig, from_reference_config
from .synth import alias_of, state_dict_spec

_IMAGE_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


class TokenRangeError(_lib.MsclipError, IndexError):
    """Out-of-range token id: an IndexError like nn.Embedding's (M.py:3047) and a library error alike."""


class _Node(nn.Module):
    """Anonymous container: only there so parameters get the reference's dotted names."""


def _descend(root: nn.Module, path):
    node = root
    for name in path:
        child = node._modules.get(name)
        if child is None:
            child = _Node()
            node.add_module(name, child)
        node = child
    return node


def _init_tensor(key: str, shape) -> torch.Tensor:
    """Reference-style initialisation (trunc-normal 0.02 weights, zero biases, identity norms,
    M.py:2344-2355, 2937-2948); real use loads a checkpoint over it."""
    if key == "logit_scale":
        return torch.ones(())                                       # M.py:2850
    if key.endswith("num_batches_tracked"):
        return torch.zeros((), dtype=torch.long)
    if key.endswith("running_mean"):
        return torch.zeros(shape)
    if key.endswith("running_var"):
        return torch.ones(shape)
    is_norm = any(s in key for s in (".bn", "downsample.1", "residual_bn", ".ln_", "ln_final", "ln_pre", "ln_post",
                                     "ln_adapt"))
    if is_norm:
        return torch.ones(shape) if key.endswith(".weight") else torch.zeros(shape)
    if key.endswith("bias"):
        return torch.zeros(shape)
    t = torch.empty(shape)
    nn.init.trunc_normal_(t, std=0.02)
    return t


class _DeviceArray:
    """A raw device pointer as a CUDA-array-interface object (zero-copy torch view of a library-owned buffer)."""
    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape) if len(shape) else (1,), "typestr": "<f4",
                                         "data": (int(ptr), False), "version": 2}


class CLIP(nn.Module):
    def __init__(self, cfg: MSCLIPConfig, precision: Optional[str] = None):
        """``precision``: MMA operand type of the library build to use - "bf16" (default, the north star's
        dtype) or "fp16" (same speed, 3 more mantissa bits, saturating conversions); env MSCLIP_PRECISION
        sets the default.  Accumulation, residual stream, LayerNorm and softmax are fp32 in both."""
        super().__init__()
        self.cfg = cfg
        self.precision = precision or _lib.default_precision()
        self._spec = state_dict_spec(cfg)
        made = {}
        for key, shape in self._spec.items():
            *path, leaf = key.split(".")
            node = _descend(self, path)
            src = alias_of(cfg, key)
            if leaf in ("running_mean", "running_var", "num_batches_tracked"):
                node.register_buffer(leaf, _init_tensor(key, shape))
            else:
                p = made[src] if src is not None else nn.Parameter(_init_tensor(key, shape), requires_grad=False)
                made[key] = p
                node.register_parameter(leaf, p)
        self._handle = C.c_void_p()
        self._synced = None         # fingerprint of the tensors last handed to the library
        self._tensors = None        # cached (key, tensor) list behind that fingerprint
        self._comm = None           # (rank, world, max_b_local) once the peer exchange is set up
        self._comm_mode = "p2p"     # "p2p": in-kernel NVLink gather; "gather": gather_tensors fallback (multi-node)
        self._checked_b = set()     # local batch sizes already verified to be equal on every rank
        self._last_loss_b = 0       # local batch of the last loss call (contrastive_loss_backward)

    # ---- reference attributes -----------------------------------------------------------------------
    @property
    def dtype(self):                                                # M.py:2973-2975
        return self.visual.proj.dtype

    @property
    def device(self):
        return self.visual.proj.device

    # ---- library handle ------------------------------------------------------------------------------
    def _library(self):
        return _lib.lib(self.precision)

    def _check(self, rc, what=""):
        _lib.check(rc, what, self.precision)

    def _ensure_handle(self):
        if not self._handle:
            c = self.cfg
            cc = _lib.Config(c.patch_size, c.layers, c.width, c.embed_dim, c.image_resolution, c.context_length,
                             c.vocab_size, (C.c_int32 * 4)(*c.early_strides), (C.c_int32 * 5)(*c.parallel_strides),
                             (C.c_int32 * 5)(*c.t2b_kernels))
            self._check(self._library().msclip_create(C.byref(cc), C.byref(self._handle)), "msclip_create")
        return self._handle

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._library().msclip_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    def _stream(self):
        # the library allocates and launches on the CURRENT device: every call runs under _on_device()
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _on_device(self):
        """Make the parameters' device current for the duration of a library call (model.to('cuda:1') without
        torch.cuda.set_device(1) must not put workspace and launches on cuda:0)."""
        dev = self.device
        if dev.type != "cuda":
            raise _lib.MsclipError("msclip_b200 needs an sm_100 GPU: there is no CPU fallback (move the model with .cuda())")
        return torch.cuda.device(dev)

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() replace parameter storage: drop the cached tensor list
        self._tensors = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._tensors = None
        return super().load_state_dict(*args, **kwargs)

    def _sync_weights(self):
        """Hand the current state_dict to the library if any tensor changed since the last call.  The (key, tensor)
        list is cached, so the steady-state cost is one (data_ptr, version) comparison per tensor (zero-shot makes
        ~1000 small encode_text calls, tools/zero_shot.py:125-131)."""
        if self._tensors is None:
            self._tensors = list(self.state_dict(keep_vars=True).items())
        finger = tuple((t.data_ptr(), t._version) for _, t in self._tensors)
        if finger == self._synced:
            return
        if not torch.cuda.is_available():
            raise _lib.MsclipError("msclip_b200 needs an sm_100 GPU: there is no CPU fallback")
        h = self._ensure_handle()
        L = self._library()
        for key, t in self._tensors:
            if t.dtype == torch.long:
                dt, tt = _lib.I64, t
            else:
                dt, tt = _lib.F32, t.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * max(tt.dim(), 1))(*tt.shape)
            self._check(L.msclip_set_weight(h, key.encode(), C.c_void_p(tt.data_ptr()), dt, tt.dim(), shape),
                       f"msclip_set_weight({key})")
        self._check(L.msclip_finalize_weights(h, self._stream()), "msclip_finalize_weights")
        self._synced = finger

    # ---- reference methods ---------------------------------------------------------------------------
    @torch.no_grad()
    def encode_image(self, image: torch.Tensor, norm: bool = True, action=None) -> torch.Tensor:
        if action is not None:
            raise AssertionError("action must be None (Gumbel search is outside the MS-CLIP-S envelope, M.py:942)")
        c = self.cfg
        if image.dim() != 4 or tuple(image.shape[1:]) != (3, c.image_resolution, c.image_resolution):
            raise ValueError(f"expected [B, 3, {c.image_resolution}, {c.image_resolution}], got {tuple(image.shape)}")
        if image.dtype not in _IMAGE_DTYPES:
            image = image.float()
        image = image.contiguous()
        with self._on_device():
            self._sync_weights()
            out = torch.empty((image.shape[0], c.embed_dim), dtype=torch.float32, device=image.device)
            self._check(self._library().msclip_encode_image(self._handle, C.c_void_p(image.data_ptr()), _IMAGE_DTYPES[image.dtype],
                                                      image.shape[0], C.c_void_p(out.data_ptr()), int(bool(norm)),
                                                      self._stream()), "msclip_encode_image")
        return out

    def prefetch_images(self, image: torch.Tensor) -> None:
        """Start the host->device copy of a (pinned) CPU image batch now; the next encode_image / contrastive_loss
        call on the same tensor consumes it (tools/zero_shot.py:262 does `.cuda(non_blocking=True)`)."""
        if image.is_cuda:
            return
        if image.dtype not in _IMAGE_DTYPES or not image.is_contiguous():
            raise ValueError("prefetch_images needs a contiguous float32 / bfloat16 / float16 CPU tensor")
        with self._on_device():
            self._sync_weights()
            self._check(self._library().msclip_stage_images(self._handle, C.c_void_p(image.data_ptr()), _IMAGE_DTYPES[image.dtype],
                                                           image.shape[0], self._stream()), "msclip_stage_images")

    @torch.no_grad()
    def encode_text(self, text: torch.Tensor, norm: bool = True, action=None) -> torch.Tensor:
        if action is not None:
            raise AssertionError("action must be None (Gumbel search is outside the MS-CLIP-S envelope, M.py:942)")
        c = self.cfg
        if text.dim() != 2 or text.shape[1] != c.context_length:
            raise ValueError(f"expected [B, {c.context_length}] token ids, got {tuple(text.shape)}")
        text = text.to(torch.long).contiguous()
        if text.numel():
            # nn.Embedding raises on out-of-range ids (M.py:3047); so do we, before anything is launched
            lo, hi = int(text.min()), int(text.max())
            if lo < 0 or hi >= c.vocab_size:
                raise TokenRangeError(f"token id out of range [0, {c.vocab_size}): min {lo}, max {hi}")
        with self._on_device():
            self._sync_weights()
            out = torch.empty((text.shape[0], c.embed_dim), dtype=torch.float32, device=text.device)
            self._check(self._library().msclip_encode_text(self._handle, C.c_void_p(text.data_ptr()), text.shape[0],
                                                     C.c_void_p(out.data_ptr()), int(bool(norm)), self._stream()),
                       "msclip_encode_text")
        return out

    def set_text_trim(self, enable: bool) -> None:
        """encode_text runs the causal tower only over the longest live prefix (up to the EOT token, M.py:3057-3060) of
        the batch - bit-identical outputs, less work for short prompts.  On by default."""
        self._ensure_handle()
        self._check(self._library().msclip_set_text_trim(self._handle, int(bool(enable))), "msclip_set_text_trim")

    @torch.no_grad()
    def similarity_logits(self, image_features: torch.Tensor, text_features: torch.Tensor, scale: float) -> torch.Tensor:
        """scale * I @ T^T (M.py:3141/3146; tools/zero_shot.py:266 with scale = 100)."""
        fi = image_features.float().contiguous()
        ft = text_features.float().contiguous()
        with self._on_device():
            self._ensure_handle()
            out = torch.empty((fi.shape[0], ft.shape[0]), dtype=torch.float32, device=fi.device)
            self._check(self._library().msclip_similarity_logits(self._handle, C.c_void_p(fi.data_ptr()), fi.shape[0],
                                                           C.c_void_p(ft.data_ptr()), ft.shape[0], float(scale),
                                                           C.c_void_p(out.data_ptr()), self._stream()),
                       "msclip_similarity_logits")
        return out

    @torch.no_grad()
    def zeroshot_classifier(self, tokens: torch.Tensor) -> torch.Tensor:
        """The tool's ``zeroshot_classifier`` (tools/zero_shot.py:121-132) in one call: ``tokens`` [n_classes,
        n_templates, context_length] -> ``zeroshot_weights`` [embed_dim, n_classes] (a transposed view of the
        [n_classes, embed_dim] result), per class the re-normalised mean of the normalised prompt embeddings."""
        c = self.cfg
        if tokens.dim() != 3 or tokens.shape[2] != c.context_length:
            raise ValueError(f"expected [n_classes, n_templates, {c.context_length}] token ids, got {tuple(tokens.shape)}")
        tokens = tokens.to(torch.long).contiguous()
        n_cls, n_tpl = int(tokens.shape[0]), int(tokens.shape[1])
        with self._on_device():
            self._sync_weights()
            out = torch.empty((n_cls, c.embed_dim), dtype=torch.float32, device=self.device)
            self._check(self._library().msclip_zeroshot_classifier(self._handle, C.c_void_p(tokens.data_ptr()), n_cls, n_tpl,
                                                                  C.c_void_p(out.data_ptr()), self._stream()),
                        "msclip_zeroshot_classifier")
        return out.t()

    @torch.no_grad()
    def zeroshot_predict(self, image_features: torch.Tensor, zeroshot_weights: torch.Tensor, topk: int = 1,
                         scale: float = 100.0, return_logits: bool = False):
        """``scale * features @ zeroshot_weights`` and the top-k classes per image (tools/zero_shot.py:266, 150-163),
        entirely on the device.  ``zeroshot_weights`` is [embed_dim, n_classes] as the tool stacks it."""
        fi = image_features.float().contiguous()
        w = zeroshot_weights.t().float().contiguous()           # [n_classes, embed_dim]
        with self._on_device():
            self._ensure_handle()
            idx = torch.empty((fi.shape[0], topk), dtype=torch.int32, device=fi.device)
            logits = torch.empty((fi.shape[0], w.shape[0]), dtype=torch.float32, device=fi.device) if return_logits else None
            self._check(self._library().msclip_zeroshot_predict(self._handle, C.c_void_p(fi.data_ptr()), fi.shape[0],
                                                               C.c_void_p(w.data_ptr()), w.shape[0], float(scale), int(topk),
                                                               C.c_void_p(idx.data_ptr()),
                                                               C.c_void_p(logits.data_ptr()) if return_logits else None,
                                                               self._stream()), "msclip_zeroshot_predict")
        return (idx, logits) if return_logits else idx

    @torch.no_grad()
    def forward(self, image: torch.Tensor, text: torch.Tensor) -> torch.Tensor:
        """CLIP.forward (M.py:3126-3155): logits over the (gathered) batch.  With gather_tensors and an
        initialised process group the features are all-gathered in rank order exactly like
        lib/utils/comm.py:140-154 (this materialising path is the reference-compatible one; the
        B200-native training-step path is ``contrastive_loss``)."""
        fi = self.encode_image(image)
        ft = self.encode_text(text)
        scale = float(self.logit_scale.detach().float().exp())
        if self.cfg.gather_tensors and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            from .comm import gather_tensors
            fi, ft = gather_tensors(fi), gather_tensors(ft)
        return self.similarity_logits(fi, ft, scale)

    # ---- the fused training-step path ------------------------------------------------------------------
    def setup_data_parallel(self, max_b_local: int, group=None):
        """Register the peer-visible embedding buffers of all ranks (one process per GPU).  ``max_b_local`` is the
        largest local batch the loss will see (with micro-batching: the whole local shard).  Ranks that cannot reach
        each other through CUDA IPC (more than one node) fall back to ``gather_tensors`` + logits + cross-entropy,
        the reference-shaped path (lib/utils/comm.py:140-154)."""
        from .comm import setup_peer_exchange
        with self._on_device():
            self._sync_weights()
            self._comm, self._comm_mode = setup_peer_exchange(self._handle, max_b_local, group, self.precision)
        self._group = group
        self._checked_b = set()
        return self._comm

    def _require_equal_shards(self, b: int):
        """gather_tensors / the P2P exchange both assume equal shards on every rank; verify once per batch size."""
        world = self._comm[1] if self._comm else 1
        if world == 1 or b in self._checked_b:
            return
        t = torch.tensor([b], dtype=torch.int64, device=self.device)
        parts = [torch.empty_like(t) for _ in range(world)]
        torch.distributed.all_gather(parts, t, group=getattr(self, "_group", None))
        sizes = [int(p) for p in parts]
        if any(x != b for x in sizes):
            raise ValueError(f"contrastive_loss needs the same local batch on every rank, got {sizes}")
        self._checked_b.add(b)

    @torch.no_grad()
    def encode_pairs(self, image: torch.Tensor, text: torch.Tensor, row_offset: int = 0) -> None:
        """Micro-batching: run both towers on ``image`` / ``text`` and keep the normalised embeddings as rows
        [row_offset, row_offset + B) of this rank's shard of the next ``loss_of_encoded`` (msclip_encode_pairs)."""
        if image.shape[0] != text.shape[0]:
            raise ValueError("image and text batch sizes differ")
        if image.dtype not in _IMAGE_DTYPES:
            image = image.float()
        image, text = image.contiguous(), text.to(torch.long).contiguous()
        with self._on_device():
            self._sync_weights()
            self._check(self._library().msclip_encode_pairs(self._handle, C.c_void_p(image.data_ptr()), _IMAGE_DTYPES[image.dtype],
                                                           C.c_void_p(text.data_ptr()), image.shape[0], int(row_offset),
                                                           self._stream()), "msclip_encode_pairs")

    @torch.no_grad()
    def loss_of_encoded(self, b_local: int, reduce: bool = True):
        """Symmetric cross-entropy over the rows retained by ``encode_pairs`` (b_local per rank)."""
        world = self._comm[1] if self._comm else 1
        self._require_equal_shards(b_local)
        self._last_loss_b = int(b_local)
        with self._on_device():
            parts = torch.empty(2, dtype=torch.float32, device=self.device)
            loss = torch.empty((), dtype=torch.float32, device=self.device)
            scale = C.c_float()
            self._check(self._library().msclip_logit_scale_exp(self._handle, C.byref(scale)), "msclip_logit_scale_exp")
            self._check(self._library().msclip_contrastive_loss(self._handle, int(b_local), scale.value, C.c_void_p(parts.data_ptr()),
                                                               C.c_void_p(loss.data_ptr()) if world == 1 else None, self._stream()),
                       "msclip_contrastive_loss")
        if world == 1:
            return loss
        if reduce:
            torch.distributed.all_reduce(parts, group=getattr(self, "_group", None))
        return parts.sum() / (2.0 * world * b_local)

    @torch.no_grad()
    def contrastive_loss(self, image: torch.Tensor, text: torch.Tensor, reduce: bool = True, micro_batch: Optional[int] = None):
        """Symmetric cross-entropy of exp(logit_scale) * I_all @ T_all^T over the global batch
        (SURVEY.md section 8a rows G, C, L).  Returns a 0-dim tensor; with world > 1 and ``reduce`` the
        per-rank partial sums are summed with one 2-float all-reduce.  ``micro_batch``: encode the local batch in
        pieces of that many pairs (their embeddings are retained on the device) and take ONE loss over all of them -
        how a global batch of 32 768 runs on fewer than 8 GPUs (needs ``setup_data_parallel(len(image))`` first)."""
        if image.shape[0] != text.shape[0]:
            raise ValueError("image and text batch sizes differ")
        b = image.shape[0]
        world = self._comm[1] if self._comm else 1
        if self._comm_mode == "gather" and world > 1:
            # ranks on different nodes: reference-shaped path (all-gather, logits, CE) - correct everywhere, not fused
            logits = self.forward(image, text)
            target = torch.arange(logits.shape[0], device=logits.device)
            ce = torch.nn.functional.cross_entropy
            return 0.5 * (ce(logits, target) + ce(logits.t(), target))
        if micro_batch is not None and micro_batch < b:
            for lo in range(0, b, micro_batch):
                self.encode_pairs(image[lo:lo + micro_batch], text[lo:lo + micro_batch], lo)
            return self.loss_of_encoded(b, reduce)
        if image.dtype not in _IMAGE_DTYPES:
            image = image.float()
        image, text = image.contiguous(), text.to(torch.long).contiguous()
        self._require_equal_shards(b)
        self._last_loss_b = int(b)
        with self._on_device():
            self._sync_weights()
            dev = self.device
            parts = torch.empty(2, dtype=torch.float32, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            self._check(self._library().msclip_forward_loss(self._handle, C.c_void_p(image.data_ptr()), _IMAGE_DTYPES[image.dtype],
                                                      C.c_void_p(text.data_ptr()), b, C.c_void_p(parts.data_ptr()),
                                                      C.c_void_p(loss.data_ptr()) if world == 1 else None, self._stream()),
                       "msclip_forward_loss")
        if world == 1:
            return loss
        if reduce:
            torch.distributed.all_reduce(parts, group=getattr(self, "_group", None))
        return parts.sum() / (2.0 * world * b)

    @torch.no_grad()
    def contrastive_loss_backward(self):
        """(d loss / d image_features, d loss / d text_features) [b_local, embed_dim] of the last ``contrastive_loss`` /
        ``loss_of_encoded`` call - the first piece of the training backward.  Every rank must
        call it; like ``gather_tensors`` (lib/utils/comm.py:151-152) the gradient reaches only the local shard."""
        with self._on_device():
            b = self._last_loss_b
            gi = torch.empty((b, self.cfg.embed_dim), dtype=torch.float32, device=self.device)
            gt = torch.empty_like(gi)
            self._check(self._library().msclip_contrastive_loss_backward(self._handle, C.c_void_p(gi.data_ptr()),
                                                                        C.c_void_p(gt.data_ptr()), self._stream()),
                        "msclip_contrastive_loss_backward")
        return gi, gt

    # ---- input pipeline on the GPU ------------------------------------------------
    @torch.no_grad()
    def preprocess(self, images, size: Optional[int] = None, mean=(0.48145466, 0.4578275, 0.40821073),
                   std=(0.26862954, 0.26130258, 0.27577711), dtype: torch.dtype = torch.float32, source_on_device: bool = False,
                   return_u8: bool = False):
        """The tool's ``transform_CLIP`` (tools/zero_shot.py:202-207: Resize(size, BICUBIC), CenterCrop, ToTensor, Normalize
        with lib/config/default.py:84-85) for a list of decoded RGB images - uint8 [H, W, 3] numpy arrays or tensors of any
        sizes - in two kernel launches, bit-exact with torchvision on PIL images.  Returns ([n, 3, size, size] tensor on the
        model's device ready for ``encode_image``, and with ``return_u8`` the resized + cropped bytes [n, size, size, 3])."""
        import numpy as np
        size = int(size or self.cfg.image_resolution)
        arrs = [im.cpu().numpy() if isinstance(im, torch.Tensor) else np.asarray(im) for im in images]
        for a in arrs:
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("preprocess expects uint8 [H, W, 3] RGB images")
        n = len(arrs)
        offs, total = [], 0
        for a in arrs:
            offs.append(total)
            total += a.size
        packed = torch.empty(total, dtype=torch.uint8, pin_memory=not source_on_device)
        flat = packed.numpy()
        for a, o in zip(arrs, offs):
            flat[o:o + a.size] = a.reshape(-1)
        with self._on_device():
            self._ensure_handle()
            if source_on_device:
                packed = packed.to(self.device)
            out = torch.empty((n, 3, size, size), dtype=dtype, device=self.device)
            u8 = torch.empty((n, size, size, 3), dtype=torch.uint8, device=self.device) if return_u8 else None
            self._check(self._library().msclip_preprocess_images(
                self._handle, C.c_void_p(packed.data_ptr()), (C.c_int64 * n)(*offs), (C.c_int * n)(*[a.shape[0] for a in arrs]),
                (C.c_int * n)(*[a.shape[1] for a in arrs]), n, size, (C.c_float * 3)(*mean), (C.c_float * 3)(*std),
                C.c_void_p(out.data_ptr()), _IMAGE_DTYPES[dtype], C.c_void_p(u8.data_ptr()) if return_u8 else None, self._stream()),
                "msclip_preprocess_images")
            torch.cuda.current_stream(self.device).synchronize()      # the pinned staging buffer is released on return
        return out, u8

    # ---- training: backward + optimiser hooks -----------------------------------
    def enable_training(self, enable: bool = True) -> None:
        """Make the library keep what the backward pass needs (transposed weight copies, per-block inputs of the
        towers, one fp32 gradient buffer per trainable key) - see include/msclip_b200.h.  The gradient buffers are
        attached zero-copy as ``.grad`` of the corresponding parameters (aliased text / vision parameters are one
        Parameter and one buffer, M.py:2786-2830), so ``torch.optim`` optimisers work unchanged and
        ``msclip_b200.optim.AdamW`` updates them with one fused kernel.  The convolutional front stays frozen."""
        with self._on_device():
            self._ensure_handle()
            self._check(self._library().msclip_train_enable(self._handle, int(bool(enable))), "msclip_train_enable")
            self._synced = None                       # re-send the weights: transposed copies are packed at finalize
            self._train_keys = {}
            if enable:
                self._sync_weights()
                self._attach_grads()

    def _attach_grads(self) -> None:
        L = self._library()
        params = dict(self.named_parameters(remove_duplicate=False))
        seen = {}
        for i in range(L.msclip_num_grads(self._handle)):
            key, ptr, numel = C.c_char_p(), C.c_void_p(), C.c_int64()
            self._check(L.msclip_grad_info(self._handle, i, C.byref(key), C.byref(ptr), C.byref(numel)), "msclip_grad_info")
            name = key.value.decode()
            p = params[name]
            if p.numel() != numel.value or p.dtype != torch.float32:
                raise _lib.MsclipError(f"gradient buffer of {name} does not match the fp32 parameter")
            if ptr.value not in seen:
                seen[ptr.value] = torch.as_tensor(_DeviceArray(ptr.value, (p.numel(),)), device=self.device).view(p.shape)
            p.grad = seen[ptr.value]
            self._train_keys[name] = p
        ls = self.logit_scale
        if ls.grad is None:
            ls.grad = torch.zeros_like(ls)
        self._train_keys["logit_scale"] = ls

    def trainable_parameters(self):
        """{state-dict key: Parameter} of everything that receives a gradient (one entry per key; aliased keys map to
        the same Parameter)."""
        if not getattr(self, "_train_keys", None):
            raise _lib.MsclipError("call enable_training() first")
        return dict(self._train_keys)

    def zero_grad(self, set_to_none: bool = False):             # the buffers belong to the library: always zero in place
        if getattr(self, "_train_keys", None):
            with self._on_device():
                self._check(self._library().msclip_zero_grad(self._handle, self._stream()), "msclip_zero_grad")
            self.logit_scale.grad.zero_()
            return
        return super().zero_grad(set_to_none)

    @torch.no_grad()
    def backward_features(self, d_image_features: Optional[torch.Tensor], d_text_features: Optional[torch.Tensor]) -> None:
        """Accumulate d loss / d parameter for the last ``encode_image`` / ``encode_text`` (or ``contrastive_loss``) call,
        given d loss / d features [B, embed_dim] of the features those calls returned (msclip_backward)."""
        gi = d_image_features.float().contiguous() if d_image_features is not None else None
        gt = d_text_features.float().contiguous() if d_text_features is not None else None
        with self._on_device():
            self._check(self._library().msclip_backward(self._handle, C.c_void_p(gi.data_ptr()) if gi is not None else None,
                                                       C.c_void_p(gt.data_ptr()) if gt is not None else None, self._stream()),
                        "msclip_backward")

    @torch.no_grad()
    def loss_and_backward(self, image: torch.Tensor, text: torch.Tensor, micro_batch: Optional[int] = None) -> torch.Tensor:
        """One training forward + backward: the fused global-batch loss (``contrastive_loss``), its backward to the local
        embeddings (second in-kernel peer pass) and the backward of both towers.  Gradients accumulate in ``.grad``;
        with more than one rank they are LOCAL (all-reduce them like DDP would).  Returns the loss.

        ``micro_batch``: local batches beyond what one tape holds (4096 pairs) - e.g. the metric's global batch of 32 768 on
        one GPU - are processed GradCache-style: every micro-batch is encoded without a tape and its embeddings retained
        (``encode_pairs``), ONE loss and ONE loss backward run over the whole global batch, then every micro-batch is encoded
        again with the tape and back-propagated with its rows of the embedding gradient.  Same gradients as the one-shot
        call (needs ``setup_data_parallel(len(image))`` first, like ``contrastive_loss(micro_batch=...)``)."""
        b = image.shape[0]
        if micro_batch is None or micro_batch >= b:
            loss = self.contrastive_loss(image, text)
            gi, gt = self.contrastive_loss_backward()
            self.backward_features(gi, gt)
            # d loss / d logit_scale = s * d loss / d s = sum_i I_i . dI_i  (dI_i = s / 2G * sum_j w_ij T_j)
            fi = self.last_image_features()
            self.logit_scale.grad += (gi * fi).sum().to(self.logit_scale.dtype)
            return loss
        loss = self.contrastive_loss(image, text, micro_batch=micro_batch)
        gi, gt = self.contrastive_loss_backward()
        acc = torch.zeros((), dtype=torch.float32, device=self.device)
        for lo in range(0, b, micro_batch):
            hi = min(b, lo + micro_batch)
            fi = self.encode_image(image[lo:hi])          # taped (training is enabled): the tape holds this micro-batch
            self.encode_text(text[lo:hi])
            self.backward_features(gi[lo:hi], gt[lo:hi])
            acc += (gi[lo:hi] * fi).sum()
        self.logit_scale.grad += acc.to(self.logit_scale.dtype)
        return loss

    @torch.no_grad()
    def allreduce_grads(self, average: bool = False) -> int:
        """Sum the LOCAL parameter gradients of ``loss_and_backward`` over the ranks (what DistributedDataParallel does), in a
        handful of bucketed collectives (``comm.allreduce_gradients``).  The global-batch loss is already normalised by the
        global batch, so the SUM is the gradient of the global loss (``average`` stays False)."""
        from .comm import allreduce_gradients
        return allreduce_gradients([p.grad for p in self.trainable_parameters().values()], getattr(self, "_group", None),
                                   average=average)

    @torch.no_grad()
    def last_image_features(self) -> torch.Tensor:
        """Normalised image features of the last taped image-tower call (read back from the library's tape)."""
        if not getattr(self, "_train_keys", None):
            raise _lib.MsclipError("call enable_training() first")
        with self._on_device():
            b = self._last_loss_b
            out = torch.empty((b, self.cfg.embed_dim), dtype=torch.float32, device=self.device)
            self._check(self._library().msclip_taped_features(self._handle, 0, C.c_void_p(out.data_ptr()), b, self._stream()),
                        "msclip_taped_features")
        return out

    @torch.no_grad()
    def refresh_weights(self, keys=None) -> None:
        """Re-pack trainable parameters from their (updated) fp32 masters - device-side, no allocation
        (msclip_update_weight).  ``keys``: state-dict keys to refresh; default all trainable ones."""
        todo = self.trainable_parameters()
        if keys is not None:
            todo = {k: todo[k] for k in keys}
        done = set()
        with self._on_device():
            for key, p in todo.items():
                if p.data_ptr() in done:
                    continue
                done.add(p.data_ptr())
                self._check(self._library().msclip_update_weight(self._handle, key.encode(), C.c_void_p(p.data_ptr()),
                                                                self._stream()), f"msclip_update_weight({key})")
        if self._tensors is not None:       # the packed copies are current: do not trigger a full re-upload
            self._synced = tuple((t.data_ptr(), t._version) for _, t in self._tensors)

    def launch_count(self) -> int:
        return int(self._library().msclip_launch_count(self._handle)) if self._handle else 0


def get_clip_model(config: Any, vocab_size: Optional[int] = None, eot_token: Optional[int] = None,
                   precision: Optional[str] = None, **kwargs) -> CLIP:
    """Drop-in for ``clip_openai_pe_res_v1.get_clip_model`` (M.py:3182-3227): accepts the reference's
    yacs config (or any duck-typed equivalent) and refuses flag combinations outside MS-CLIP-S."""
    cfg = config if isinstance(config, MSCLIPConfig) else from_reference_config(config)
    if vocab_size is not None and int(vocab_size) != cfg.vocab_size:
        cfg = MSCLIPConfig(**{**cfg.to_dict(), "vocab_size": int(vocab_size)})
    return CLIP(cfg, precision)
