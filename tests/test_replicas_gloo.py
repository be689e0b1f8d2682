"""The N > 1 path of bench.py (replicas only: independent seeds, barrier + MAX-over-ranks timing) on CPU with gloo,
world_size 2.  The data path itself has no collective, so this is all the distributed logic there is."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from voxelmapplus_fastlio2_b200 import replicas, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, _ = replicas.rank_env()
    seq = synth.Sequence(sensor=synth.SensorConfig(pts_per_scan=256), seed=replicas.rank_seed(r))
    pk = next(iter(seq.packages(1)))
    local_ms = 10.0 * (rank + 1)                      # rank 1 is the slow one
    dist.barrier()
    value, ms_per_step = replicas.whole_job_throughput(5, local_ms)
    out.put((rank, w, float(pk.cloud[:, :3].sum()), value, ms_per_step))
    dist.barrier()
    dist.destroy_process_group()


def test_replica_protocol_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, s0, v0, m0), (r1, w1, s1, v1, m1) = res
    assert (r0, r1) == (0, 1) and w0 == w1 == 2
    assert s0 != s1, "replicas must run different sequences (seed = base + rank)"
    # whole-job value = units of all ranks / MAX over ranks of the time: 2 * 5 scans / 20 ms
    assert v0 == v1 == 2 * 5 / 0.020
    assert m0 == m1 == 4.0


def test_single_process_passthrough():
    v, ms = replicas.whole_job_throughput(10, 5.0)
    assert v == 10 / 0.005 and ms == 0.5
    assert replicas.rank_seed(3) == 0xC0FFEE + 3
    assert np.isfinite(replicas.max_over_ranks(1.5))
