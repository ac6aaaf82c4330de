"""Multi-GPU host logic on CPU: world_size-2 gloo process groups exercise the row-band split and the tile
gather that bench.py / render_frame_distributed use with NCCL on the GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adanerf_b200.tiling import gather_bands, row_bands


def test_row_bands_cover_the_frame():
    for H in (1, 7, 800, 801, 1600):
        for world in (1, 2, 3, 4, 8):
            b = row_bands(H, world)
            assert len(b) == world and b[0][0] == 0
            assert sum(r for _, r in b) == H
            for (r0, rows), (n0, _) in zip(b, b[1:]):
                assert r0 + rows == n0
            assert max(r for _, r in b) - min(r for _, r in b) <= 1


def _fake_band(row0, rows, W):
    y = torch.arange(row0, row0 + rows).repeat_interleave(W)
    x = torch.arange(W).repeat(rows)
    return torch.stack([y.float(), x.float(), (y * W + x).float()], 1)


def _worker(rank, world, port, H, W, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        row0, rows = row_bands(H, world)[rank]
        band = _fake_band(row0, rows, W)
        full = gather_bands(band, W, H)
        ok = torch.equal(full, _fake_band(0, H, W))
        on_root = gather_bands(band, W, H, dst=0)
        ok = ok and ((on_root is None) if rank != 0 else torch.equal(on_root, _fake_band(0, H, W)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H,W", [(8, 5), (7, 3)])   # equal bands (all_gather_into_tensor) and ragged bands
def test_gather_bands_world2_gloo(H, W):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
