"""CPU, world_size 2 over gloo: the bucketed gradient averaging used by the clip-sharded data-parallel path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import off_b200  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ranges, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from off_b200.dist import GradAllReducer
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(n, generator=g)
    red = GradAllReducer(flat, ranges)
    for i in range(len(ranges)):   # stage/head bucket first, then one bucket per OFF unit (the order the backward finishes them)
        red.launch(i)
    red.finish()
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_bucketed_allreduce_averages_across_ranks(tmp_path):
    from off_b200 import engine as E
    eng = E.OFFEngine(2, 3, "rgb", "cpu", "tf32")             # the plan (and its bucket layout) is pure host logic
    n = eng.n_flat
    ranges = [eng.stage_range] + list(eng.unit_ranges.values())
    # the buckets tile the flat gradient buffer exactly once, on 16-byte boundaries
    cover = sorted(ranges)
    assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert all(lo % 4 == 0 for lo, _ in ranges) and list(eng.unit_ranges) == ["3a", "3b", "3c", "4a", "4b", "4c", "4d", "5a", "5b"]
    assert eng.unit_ranges["3a"][1] - eng.unit_ranges["3a"][0] >= 160 * 256 + 160 + 32 * 9 + 32
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), n, ranges, out), nprocs=2, join=True)
    want = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(2)) / 2
    got = torch.load(out)
    assert torch.allclose(got, want, atol=1e-7)
