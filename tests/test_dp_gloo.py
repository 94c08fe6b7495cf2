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


def test_grad_buckets_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
