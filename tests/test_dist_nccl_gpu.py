"""BASELINE.json configs[3] on hardware: the ONE collective of the path.  Two ranks (NCCL, one GPU each) run
``SSCMetrics(sync_dist=...)`` over their frame shards; every rank must end up with the single-process oracle's counts
over ALL frames, for C = 2 and C = 9 (muvo/trainer.py:483-490, muvo/metrics.py:77-100).  Skipped on boxes with one GPU
(there ``bench.py`` under torchrun asserts the same equality, ``stages.ssc_counts*.allreduced_counts_equal_oracle``)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from muvo_b200 import synth
from muvo_b200.distributed import shard_frames

pytestmark = pytest.mark.gpu
N_FRAMES = 8
SIZE = (96, 96, 32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import muvo_b200
        out = {}
        for C in (2, 9):
            yp, yt = synth.occupancy_pair(N_FRAMES, C, 4400 + C, size=SIZE)
            mine = shard_frames(N_FRAMES, rank, world)
            tp, tt = torch.from_numpy(yp[mine]).cuda(), torch.from_numpy(yt[mine]).cuda()
            for mode in (True, "epoch"):
                m = muvo_b200.SSCMetrics(C, sync_dist=mode)
                for k in range(0, len(mine), 2):                      # two frames per add_batch, several batches
                    m.add_batch(tp[k:k + 2], tt[k:k + 2])
                m.get_stats()
                out[(C, str(mode))] = m.counts_exact.numpy().copy()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL all-reduce of the IoU counts)")
def test_sharded_counts_allreduced_over_nccl_equal_single_process_oracle():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for C in (2, 9):
        yp, yt = synth.occupancy_pair(N_FRAMES, C, 4400 + C, size=SIZE)
        want = O.ssc_add_batch_counts(yp, yt, C)
        for _, out in res:
            for mode in ("True", "epoch"):
                assert np.array_equal(out[(C, mode)], want), (C, mode)
