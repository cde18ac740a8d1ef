"""world_size-2 gloo run of the multi-rank path: frames sharded by rank, only the IoU counts are exchanged."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from muvo_b200 import synth
from muvo_b200.distributed import shard_frames
from muvo_b200.metrics import all_reduce_counts

N_FRAMES, C = 6, 9
SIZE = (24, 24, 8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        yp, yt = synth.occupancy_pair(N_FRAMES, C, 4004, size=SIZE)
        mine = shard_frames(N_FRAMES, rank, world)
        # rank-local counts (CPU stand-in for the kernel: this test covers the exchange, not the kernel)
        local = torch.from_numpy(O.ssc_add_batch_counts(yp[mine], yt[mine], C))
        total = all_reduce_counts(local.clone())
        q.put((rank, mine, total.numpy()))
    finally:
        dist.destroy_process_group()


def test_counts_allreduce_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    yp, yt = synth.occupancy_pair(N_FRAMES, C, 4004, size=SIZE)
    want = O.ssc_add_batch_counts(yp, yt, C)
    frames = sorted(sum((r[1] for r in res), []))
    assert frames == list(range(N_FRAMES))              # disjoint cover, no data-path collective
    for _, _, tot in res:
        assert np.array_equal(tot, want)                # every rank holds the global counts


def test_all_reduce_is_noop_without_process_group():
    t = torch.arange(9)
    assert torch.equal(all_reduce_counts(t.clone()), t)
