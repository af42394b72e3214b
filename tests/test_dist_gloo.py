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
    red.launch(0)          # stage/head bucket goes first (overlaps the unit backward on the GPU path)
    red.launch(1)
    red.finish()
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_bucketed_allreduce_averages_across_ranks(tmp_path):
    from off_b200 import spec as S
    _, n = S.flat_layout("rgb")
    unit_end = 895200
    ranges = [(unit_end, n), (0, unit_end)]
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), n, ranges, out), nprocs=2, join=True)
    want = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(2)) / 2
    got = torch.load(out)
    assert torch.allclose(got, want, atol=1e-7)
