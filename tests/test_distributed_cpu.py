"""Host logic of the batch-sharded path on 2 CPU processes (gloo): shard bounds and the final gather."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastfourierdiffusion_b200.distributed import gather_samples, sample_sharded, shard_bounds, warm_communicator


def test_shard_bounds_cover_exactly_once():
    for n in (0, 1, 7, 8, 16384, 16385):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
                assert e0 == b1 and e0 >= b0
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        warm_communicator(torch.device("cpu"))

        def fake_sampler(n_local, r):
            b, _ = shard_bounds(n, world, r)
            # sample i of the global batch is filled with the value i: the gather must restore the order
            return (torch.arange(b, b + n_local, dtype=torch.float32)[:, None, None]).expand(n_local, 5, 2).contiguous()

        full = sample_sharded(fake_sampler, n)
        assert full.shape == (n, 5, 2)
        assert torch.equal(full[:, 0, 0], torch.arange(n, dtype=torch.float32))
        # a second gather through the lower-level entry, ragged shard sizes included
        b, e = shard_bounds(n, world, rank)
        again = gather_samples(full[b:e].clone(), n)
        assert torch.equal(again, full)
        # ... and into a preallocated result
        pre = torch.full((n, 5, 2), -1.0)
        assert gather_samples(full[b:e].clone(), n, out=pre) is pre and torch.equal(pre, full)
        torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])
def test_two_rank_gather_gloo(tmp_path, n):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(a, b) and a.shape[0] == n


def test_single_process_passthrough():
    x = torch.randn(4, 3, 1)
    assert gather_samples(x, 4) is x
    assert torch.equal(sample_sharded(lambda n, r: x[:n], 4), x)
