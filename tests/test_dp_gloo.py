"""CPU, world_size 2 over gloo: the data-parallel gradient-bucket logic of engine.GradBuckets (bucket formation over
one flat buffer, firing as gradients complete in arbitrary order, unfired buckets reduced at finish, mean semantics).
The same class runs over NCCL/NVLink on the GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from medical_image_generation_b200.engine import GradBuckets
        sizes = [64, 128, 64, 256, 64, 192, 64]              # padded member sizes, buffer order
        flat = torch.zeros(sum(sizes))
        gb = GradBuckets(flat, sizes, cap_elems=256)
        # buckets are contiguous, cover the buffer exactly, and respect the cap
        assert gb.buckets[0]["lo"] == 0 and gb.buckets[-1]["hi"] == flat.numel()
        for a, b in zip(gb.buckets, gb.buckets[1:]):
            assert a["hi"] == b["lo"]
        assert len(gb.buckets) == 4 and [b["n"] for b in gb.buckets] == [3, 1, 2, 1]
        offs = [sum(sizes[:i]) for i in range(len(sizes))]
        for step in range(3):
            gb.reset()
            flat.zero_()
            g = torch.Generator().manual_seed(100 * step + rank)
            local = torch.randn(flat.numel(), generator=g)
            order = [3, 0, 6, 2, 1, 5] if step % 2 == 0 else [5, 4, 3, 2, 1, 0, 6]   # member 4 unused on even steps
            for i in order:
                flat[offs[i]:offs[i] + sizes[i]] = local[offs[i]:offs[i] + sizes[i]]
                gb.ready(i)
                gb.ready(i)                                    # duplicate notifications are ignored
            fired = [b["work"] is not None for b in gb.buckets]
            if step % 2 == 0:
                assert fired == [True, True, False, True]      # bucket holding member 4 waits for finish()
            gb.finish()
            want = torch.zeros_like(flat)
            for r in range(world):
                gr = torch.Generator().manual_seed(100 * step + r)
                lr = torch.randn(flat.numel(), generator=gr)
                for i in order:
                    want[offs[i]:offs[i] + sizes[i]] += lr[offs[i]:offs[i] + sizes[i]] / world
            assert torch.allclose(flat, want, atol=1e-6), (rank, step)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def _worker_sharded(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from medical_image_generation_b200.engine import ShardedBuckets
        bucket = 256                                              # multiple of world * 64
        # sharded members (buffer order); member 1 spans buckets 0-1, member 3 spans buckets 1-2
        sizes = [128, 192, 64, 320, 64]
        shard_numel = 768                                         # 3 buckets
        spans, off = [], 0
        for sz in sizes:
            spans.append((off, off + sz))
            off += sz
        assert off == shard_numel
        repl = [64, 64]                                           # replicated region: two fp32-consumed members
        flat = torch.zeros(shard_numel + sum(repl) + 64)          # + never-used tail
        sb = ShardedBuckets(flat, spans, shard_numel, bucket, shard_numel, shard_numel + sum(repl), repl)
        assert sb.nb == 3 and sb.piece == 128 and [b["n"] for b in sb.buckets] == [2, 3, 2, 2]
        assert sb.member_buckets == [[0], [0, 1], [1], [1, 2], [2], [3], [3]]
        offs = [s[0] for s in spans] + [shard_numel, shard_numel + 64]
        allsz = sizes + repl
        for step in range(2):
            sb.reset()
            flat.zero_()
            g = torch.Generator().manual_seed(10 * step + rank)
            local = torch.randn(flat.numel(), generator=g)
            order = [6, 4, 3, 2, 1, 0, 5] if step == 0 else [0, 1, 2, 3, 5, 6]    # member 4 unused in step 1
            for i in order:
                flat[offs[i]:offs[i] + allsz[i]] = local[offs[i]:offs[i] + allsz[i]]
                fired_before = [b["work"] is not None for b in sb.buckets]
                sb.ready(i)
                if step == 0 and i == 3:      # bucket 2 needs members 3 and 4 (done), bucket 1 still waits for member 1
                    assert [b["work"] is not None for b in sb.buckets][:3] == [False, False, True], fired_before
            sb.finish()
            want = torch.zeros_like(flat)
            for r in range(world):
                lr = torch.randn(flat.numel(), generator=torch.Generator().manual_seed(10 * step + r))
                for i in order:
                    want[offs[i]:offs[i] + allsz[i]] += lr[offs[i]:offs[i] + allsz[i]] / world
            for b in range(sb.nb):            # the OWNED slice of every bucket holds the mean gradient
                lo = b * bucket + rank * sb.piece
                assert torch.allclose(flat[lo:lo + sb.piece], want[lo:lo + sb.piece], atol=1e-6), (rank, step, b)
            lo, hi = shard_numel, shard_numel + sum(repl)
            assert torch.allclose(flat[lo:hi], want[lo:hi], atol=1e-6)           # replicated region: whole mean
            # "optimiser": every rank rewrites its owned slices, gather() makes the buffer identical everywhere
            upd = torch.full((flat.numel(),), -1.0)
            for b in range(sb.nb):
                sb.owned(upd, b).copy_(sb.owned(want, b) * 2)
            sb.gather(upd)
            assert torch.allclose(upd[:shard_numel], want[:shard_numel] * 2, atol=1e-6)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_buckets_world2_gloo():
    """ShardedBuckets (reduce-scatter of uniform buckets + replicated region + all-gather of the owned slices) with
    parameters spanning several buckets; gloo falls back to all_reduce / all_gather lists, the host logic is the same."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_sharded, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_grad_buckets_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
