"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

The tracking path does not shard: a frame is ~0.5 MB of latency-bound work and one NVLink all-reduce
costs more than ten CG iterations, so N GPUs run N independent camera streams ("replicas only") with no
data-path collective. The only collective is the timing reduction of the benchmark: whole-job throughput = units of all
ranks / max-over-ranks time.

For bundle adjustment the path does shard over landmarks: `shard_landmarks` is the spatially coherent
partition (Morton order, equal counts) with the halo bookkeeping a landmark-sharded BA needs — the landmarks of another
shard that a shard's regulariser edges touch. `attach_shards` + `gather_sharded_ba` wrap the landmark-sharded BA of the C ABI
(nrslam_b200_local_ba_sharded): the ranks exchange halo rows and partial sums inside the persistent kernel through
peer-mapped buffers; torch.distributed only carries the IPC handles once and gathers the results.
"""
import numpy as np


def stream_seed(base_seed, rank):
    """Every rank tracks its own stream: a different seeded frame of the same shape (weak scaling)."""
    return int(base_seed) + int(rank)


def aggregate_throughput(units_local, ms_local, dist=None, device="cpu"):
    """(sum of units over ranks) / (max of time over ranks), as units per second. `dist` = torch.distributed or None."""
    import torch
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u[0]) / (float(t[0]) * 1e-3), float(t[0])


def morton_order(P):
    """Permutation that sorts points along a 30-bit Morton curve (same construction as the engine's row order,
    nrs_api.cu: sort_rows)."""
    P = np.asarray(P, np.float64)
    lo = P.min(0)
    ext = max(float((P.max(0) - lo).max()), 1e-300)
    q = np.clip(((P - lo) / ext * 1023.0).astype(np.int64), 0, 1023)
    code = np.zeros(len(P), np.int64)
    for b in range(9, -1, -1):
        for a in range(3):
            code = (code << 1) | ((q[:, a] >> b) & 1)
    return np.argsort(code, kind="stable")


def shard_landmarks(positions, rowptr, col, n_shards):
    """Partition landmarks into `n_shards` spatially coherent, equally sized shards.
    Returns owner[n] (shard of every landmark) and halo[s] = sorted array of landmarks owned by another shard that
    shard s reads through its regulariser edges (CSR graph rowptr / col over landmarks)."""
    n = len(positions)
    order = morton_order(positions)
    owner = np.empty(n, np.int32)
    bounds = np.linspace(0, n, n_shards + 1).astype(np.int64)
    for s in range(n_shards):
        owner[order[bounds[s]:bounds[s + 1]]] = s
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    cross = owner[rows] != owner[col]
    halo = []
    for s in range(n_shards):
        m = cross & (owner[rows] == s)
        halo.append(np.unique(col[m]).astype(np.int32))
    return owner, halo


def attach_shards(core, dist, max_rows, max_poses):
    """One-time set-up of the sharded BA on an initialised process group: every rank allocates its exchange buffer,
    the 64-byte CUDA IPC handles are all-gathered, every rank maps its peers' buffers."""
    rank, world = dist.get_rank(), dist.get_world_size()
    handle = core.shard_init(rank, world, max_rows, max_poses)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    core.shard_attach(handles)
    dist.barrier()


def gather_sharded_ba(result, dist):
    """Completes X on every rank from the owners' parts (the poses are already identical everywhere)."""
    import torch
    world = dist.get_world_size()
    parts = [None] * world
    dist.all_gather_object(parts, result["X"])
    X = result["X"].copy()
    for r in range(world):
        m = result["owner"] == r
        X[m] = parts[r][m]
    return X
