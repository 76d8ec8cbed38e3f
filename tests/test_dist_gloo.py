"""N > 1 host logic on CPU: world_size-2 gloo groups (one process per rank, 127.0.0.1 rendezvous)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nrslam_b200 import dist as nd, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # replicas: every rank draws its own stream and times its own steps
    seed = nd.stream_seed(1235, rank)
    p = synth.tracking_problem("c1", seed=seed, n=60)
    ms_local = 10.0 * (rank + 1)           # rank 1 is the slow one
    value, ms_max = nd.aggregate_throughput(5, ms_local, dist)
    # landmark shards are a pure function of the geometry: identical on every rank without communication
    q = synth.tracking_problem("c1", seed=99, n=60)   # the BA geometry is common to all ranks
    owner, halo = nd.shard_landmarks(q["last_world_position"], q["graph"].rowptr, q["graph"].col, world)
    t = torch.tensor(owner.astype(np.int64))
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    out[rank] = dict(seed=seed, value=value, ms_max=ms_max, uv0=float(p["uv"][0, 0]),
                     same_partition=bool(all(torch.equal(g[0], x) for x in g)))
    dist.destroy_process_group()


def test_replica_streams_and_throughput_aggregation():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert a["seed"] != b["seed"] and a["uv0"] != b["uv0"]            # different streams
    assert a["ms_max"] == b["ms_max"] == 20.0                         # max over ranks
    assert abs(a["value"] - 10 / 0.020) < 1e-9 and a["value"] == b["value"]   # units of all ranks / max time
    assert a["same_partition"] and b["same_partition"]


def test_shard_landmarks_partition_and_halo():
    p = synth.tracking_problem("c2", n=800)
    P, g = p["last_world_position"], p["graph"]
    for n_shards in (1, 2, 4, 8):
        owner, halo = nd.shard_landmarks(P, g.rowptr, g.col, n_shards)
        counts = np.bincount(owner, minlength=n_shards)
        assert counts.sum() == len(P) and counts.max() - counts.min() <= 1
        rows = np.repeat(np.arange(len(P)), np.diff(g.rowptr))
        for s in range(n_shards):
            need = np.unique(g.col[(owner[rows] == s) & (owner[g.col] != s)])
            assert np.array_equal(need, halo[s])
            assert not np.any(owner[halo[s]] == s)
        if n_shards > 1:   # spatial coherence: most edges stay inside a shard
            assert (owner[rows] == owner[g.col]).mean() > 0.8


def test_single_process_aggregation_without_group():
    v, ms = nd.aggregate_throughput(7, 14.0, None)
    assert abs(v - 500.0) < 1e-9 and ms == 14.0
