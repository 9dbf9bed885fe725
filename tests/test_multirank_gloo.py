"""world_size-2 gloo test of the multi-GPU host logic: rank 0 broadcasts the model image,
every rank parses it and owns a contiguous block of streams; no data-path collective exists."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import nnnoiseless_b200 as nb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = torch.zeros(1, dtype=torch.int64)
        if rank == 0:
            img = nb.RnnModel().to_bytes()
            n[0] = len(img)
        dist.broadcast(n, 0)
        buf = torch.zeros(int(n[0]), dtype=torch.uint8)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(img), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        model = nb.RnnModel.from_bytes(buf.numpy().tobytes())
        assert model is not None
        start, count = nb.shard_streams(37, world, rank)
        # aggregate: total streams and a checksum of the model every rank parsed
        t = torch.tensor([count, int(np.frombuffer(model.to_bytes(), np.uint8).astype(np.int64).sum())], dtype=torch.int64)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        q.put((rank, start, count, [x.tolist() for x in g]))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_shard_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, c0, g0), (r1, s1, c1, g1) = res
    assert (s0, c0, s1, c1) == (0, 19, 19, 18)
    assert g0 == g1 and g0[0][1] == g0[1][1] and g0[0][0] + g0[1][0] == 37
