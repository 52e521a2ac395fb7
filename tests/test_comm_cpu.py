"""Host logic of the data-parallel path on CPU: world_size-2 gloo process group (the GPU path uses the
same code with nccl for the one-off handle exchange and no collective at all on the data path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msclip_b200 import comm


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert comm.rank_world() == (rank, world)
        lo, hi = comm.shard_range(rank, world, 8)
        full = torch.arange(8 * 4, dtype=torch.float32).view(8, 4)
        mine = full[lo:hi].clone()
        g = comm.gather_tensors(mine)
        ok_gather = torch.equal(g, full) and g[lo:hi].data_ptr() != 0
        payload = bytes([rank] * 64)
        both = comm.exchange_bytes(payload)
        ok_bytes = both == bytes([0] * 64) + bytes([1] * 64)
        # per-rank partial losses combine like the fused kernel's partial sums
        parts = torch.tensor([1.0 + rank, 2.0 + rank])
        dist.all_reduce(parts)
        # bucketed gradient all-reduce (what CLIP.allreduce_grads runs over NCCL): shared storage reduced once, several buckets
        g = torch.Generator().manual_seed(7)
        base = [torch.randn(n, generator=g) for n in (5, 300, 1, 64, 1000)]
        grads = [(b * (rank + 1)).clone() for b in base]
        grads.append(grads[1])                                   # an aliased parameter: same tensor twice
        calls = comm.allreduce_gradients(grads, bucket_bytes=1300)
        ok_grads = all(torch.allclose(x, b * 3.0) for x, b in zip(grads[:5], base)) and calls >= 2
        avg = [(b * (rank + 1)).clone() for b in base]
        comm.allreduce_gradients(avg, average=True)
        ok_grads = ok_grads and all(torch.allclose(x, b * 1.5) for x, b in zip(avg, base))
        out.put((rank, ok_gather, ok_bytes and ok_grads, parts.tolist()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_gather_and_handle_exchange():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_gather, ok_bytes, parts in res:
        assert ok_gather and ok_bytes and parts == [3.0, 5.0]


def test_single_process_degrades_to_world_1():
    assert comm.rank_world() == (0, 1)
    t = torch.randn(3, 4)
    assert comm.gather_tensors(t) is t
    assert comm.exchange_bytes(b"abc") == b"abc"
    assert comm.shard_range(3, 8, 32768) == (12288, 16384)
    assert comm.allreduce_gradients([torch.ones(3)]) == 0
    with pytest.raises(ValueError):
        comm.shard_range(0, 3, 8)
