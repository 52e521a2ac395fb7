"""Data-parallel plumbing for the contrastive step (one process per GPU).

Mirrors lib/utils/comm.py of the reference for the one collective on the path:

* ``gather_tensors(t)``  (comm.py:140-154) — rank-ordered all-gather; kept as the reference-compatible
  (and comparator) path used by ``CLIP.forward`` when logits must be materialised.
* ``setup_peer_exchange`` — the B200-native replacement: no collective at all on the data path.  Every
  rank's library handle owns an exchange buffer; its CUDA IPC handle is swapped once through
  ``torch.distributed`` (plumbing), after which the fused loss kernel of each rank reads the peers'
  embeddings straight over NVLink.

``shard_range`` is the row-ownership rule both paths share: global row = rank * b_local + i.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch
import torch.distributed as dist

from . import _lib


def rank_world(group=None) -> Tuple[int, int]:
    """(rank, world) — degrades to (0, 1) without an initialised process group (comm.py:16-30)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(rank: int, world: int, global_batch: int) -> Tuple[int, int]:
    """Rows [lo, hi) of the global batch owned by ``rank`` (equal shards, rank order)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    b = global_batch // world
    return rank * b, (rank + 1) * b


def gather_tensors(tensor: torch.Tensor, group=None) -> torch.Tensor:
    """Rank-ordered concat of every rank's tensor along dim 0; the local shard is re-inserted so it
    keeps its identity (comm.py:145-153)."""
    rank, world = rank_world(group)
    if world == 1:
        return tensor
    parts = [torch.empty_like(tensor) for _ in range(world)]
    dist.all_gather(parts, tensor.contiguous(), group=group)
    parts[rank] = tensor
    return torch.cat(parts, dim=0)


def exchange_bytes(payload: bytes, group=None) -> bytes:
    """All-gather a fixed-size byte string from every rank, rank-ordered (works on gloo and nccl)."""
    rank, world = rank_world(group)
    if world == 1:
        return payload
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return b"".join(bytes(p.cpu().tolist()) for p in parts)


def allreduce_gradients(tensors, group=None, bucket_bytes: int = 256 << 20, average: bool = False) -> int:
    """Sum (or average) gradient tensors over the ranks of ``group`` in place, like DistributedDataParallel would: the
    tensors are packed into flat buckets of at most ``bucket_bytes`` so that a model's ~150 gradient buffers travel in a
    handful of collectives (NVSwitch makes the cost latency-, not link-bound: few large messages).  Tensors that share
    storage (parameters shared by the two towers, M.py:2786-2830) are reduced once.  Returns the number of collectives."""
    rank, world = rank_world(group)
    if world == 1:
        return 0
    uniq, seen = [], set()
    for t in tensors:
        if t is None or t.data_ptr() in seen:
            continue
        seen.add(t.data_ptr())
        uniq.append(t)
    calls, i = 0, 0
    while i < len(uniq):
        bucket, size = [], 0
        while i < len(uniq) and (not bucket or size + uniq[i].numel() * uniq[i].element_size() <= bucket_bytes) \
                and (not bucket or (uniq[i].dtype == bucket[0].dtype and uniq[i].device == bucket[0].device)):
            bucket.append(uniq[i])
            size += uniq[i].numel() * uniq[i].element_size()
            i += 1
        flat = torch.cat([t.reshape(-1) for t in bucket]) if len(bucket) > 1 else bucket[0].reshape(-1)
        dist.all_reduce(flat, group=group)
        if average:
            flat /= world
        if len(bucket) > 1:
            off = 0
            for t in bucket:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n
        elif average or flat.data_ptr() != bucket[0].data_ptr():
            bucket[0].copy_(flat.view_as(bucket[0]))
        calls += 1
    return calls


MAX_P2P_WORLD = 64     # publish-flag slots of the exchange buffer (msclip_comm_init refuses more)


def same_node(group=None) -> bool:
    """True when every rank of the group runs on this host (CUDA IPC reaches only GPUs of one node)."""
    import socket
    rank, world = rank_world(group)
    if world == 1:
        return True
    names = [None] * world
    dist.all_gather_object(names, socket.gethostname(), group=group)
    return len(set(names)) == 1


def setup_peer_exchange(handle, max_b_local: int, group=None, precision=None):
    """comm_init -> export IPC handle -> all-gather the 64-byte handles -> import.  Returns
    ((rank, world, max_b_local), mode): mode "p2p" = in-kernel NVLink exchange; "gather" = the ranks span several
    nodes (or exceed the flag slots), so the caller must use ``gather_tensors`` + logits + CE instead.  Every rank
    must ask for the same ``max_b_local``."""
    rank, world = rank_world(group)
    L = _lib.lib(precision)
    if world > 1:
        sizes = [None] * world
        dist.all_gather_object(sizes, int(max_b_local), group=group)
        if len(set(sizes)) != 1:
            raise ValueError(f"setup_peer_exchange: max_b_local differs between ranks: {sizes}")
        if world > MAX_P2P_WORLD or not same_node(group):
            return (rank, world, int(max_b_local)), "gather"
    _lib.check(L.msclip_comm_init(handle, rank, world, int(max_b_local)), "msclip_comm_init", precision)
    if world > 1:
        buf = (C.c_uint8 * 64)()
        _lib.check(L.msclip_comm_export(handle, buf), "msclip_comm_export", precision)
        everyone = exchange_bytes(bytes(buf), group)
        assert len(everyone) == 64 * world
        arr = (C.c_uint8 * len(everyone)).from_buffer_copy(everyone)
        _lib.check(L.msclip_comm_import(handle, arr), "msclip_comm_import", precision)
        dist.barrier(group)
    return (rank, world, int(max_b_local)), "p2p"
