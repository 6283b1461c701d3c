"""World-size-2 (and 3) test of the multi-GPU host logic on CPU with the gloo backend: band ownership covers the frame once,
and the all-gather assembly of per-rank rows reproduces the full frame.  (The NCCL path runs the same code on B200s.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opencl_render_b200 import api, dist as odist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, band, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = (np.arange(3 * h * w, dtype=np.int64).reshape(3, h, w) % 32749).astype(np.int16)   # the "rendered" frame
        part = odist.BandPartition(h, w, rank, world, band)
        planes = torch.zeros((3, h, w), dtype=torch.int16)
        rows = torch.as_tensor(part.rows, dtype=torch.long)
        planes[:, rows, :] = torch.from_numpy(full)[:, rows, :]                 # this rank only "rendered" its own rows
        g = odist.PlaneGather(None, part, torch.device("cpu"), planes=planes)
        out = g.run().numpy()
        q.put((rank, bool(np.array_equal(out, full)), part.owned_rows))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h,w,band", [(2, 100, 37, 16), (3, 130, 8, 16), (2, 64, 5, 128)])
def test_band_gather_gloo(world, h, w, band):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, band, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == h


def test_owned_rows_match_c_partition():
    # python mirror (dist.owned_rows) == C-ABI oclr_band_partition == kernel map_row()
    for h, world, band in [(1080, 8, 16), (1080, 8, 128), (2160, 4, 128), (50, 3, 16)]:
        for r in range(world):
            rows = odist.owned_rows(h, r, world, band)
            c = [y for b, e in api.band_partition(h, r, world, band) for y in range(b, e)]
            assert rows.tolist() == c
