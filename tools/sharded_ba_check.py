"""Landmark-sharded BA vs the single-GPU BA on the same window (SURVEY §8e correctness gate). Launch with
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/sharded_ba_check.py [--config c3] [--landmarks 1500 --keyframes 10 --visible 6] [--reps 3]
One process per GPU; rank 0 prints one JSON line and every rank exits non-zero on a mismatch."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402

from nrslam_b200 import api, dist as nd, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--landmarks", type=int, default=None)
    ap.add_argument("--keyframes", type=int, default=None)
    ap.add_argument("--visible", type=int, default=None)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--pose-tol", type=float, default=1e-5)
    ap.add_argument("--pt-tol", type=float, default=2e-4)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if rank != 0:  # phase counters (NRSLAM_B200_PROF) from one rank only: the ranks' stderr lines interleave
        os.environ.pop("NRSLAM_B200_PROF", None)
    kw = {k: v for k, v in dict(n=a.landmarks, n_kf=a.keyframes, run=a.visible).items() if v is not None}
    p = synth.ba_problem(a.config, **kw)
    args = (p["cam"], p["kf_pose"], p["obs_kf"], p["obs_vertex"], p["uv"], p["X"], p["graph"], p["scale"])
    core = api.Core()
    single = core.local_ba(*args)
    t_single = min(core.resolve(2)["gpu_ms"] for _ in range(a.reps))
    part = api.shard_partition(world, *args[1:])
    max_rows = int((part["n_own"] + part["n_halo"]).max())
    nd.attach_shards(core, dist, max_rows, len(p["kf_pose"]))
    ms = []
    for _ in range(a.reps):
        dist.barrier()
        t0 = time.perf_counter()
        sh = core.local_ba_sharded(*args)
        ms.append((time.perf_counter() - t0) * 1e3)
    X = nd.gather_sharded_ba(sh, dist)
    poses = [None] * world
    dist.all_gather_object(poses, sh["kf_pose"])
    same_pose = all(np.array_equal(poses[0], q) for q in poses)
    d_pose = float(np.abs(sh["kf_pose"] - single["kf_pose"]).max())
    d_pt = float(np.abs(X - single["X"]).max())
    ta, tb = np.array(single["stats"]["chi2_trace"]), np.array(sh["stats"]["chi2_trace"])
    trace_ok = len(ta) == len(tb) and bool(np.allclose(ta, tb, rtol=1e-6))
    ok = same_pose and d_pose < a.pose_tol and d_pt < a.pt_tol and trace_ok
    gms = [None] * world
    dist.all_gather_object(gms, sh["stats"]["gpu_ms"])
    if rank == 0:
        print(json.dumps(dict(
            ok=bool(ok), world=world, config=a.config, n_obs=int(len(p["obs_kf"])), n_kf=int(len(p["kf_pose"])),
            rows_per_rank=part["n_own"].tolist(), halo_per_rank=part["n_halo"].tolist(),
            poses_identical_on_all_ranks=bool(same_pose), max_pose_diff=d_pose, max_point_diff=d_pt,
            chi2_trace_single=ta.tolist(), chi2_trace_sharded=tb.tolist(),
            lm_iterations=sh["stats"]["lm_iterations"], pcg_iterations_single=single["stats"]["pcg_iterations"],
            pcg_iterations_sharded=sh["stats"]["pcg_iterations"], single_gpu_ms=t_single,
            sharded_gpu_ms_max=float(max(gms)), sharded_call_ms=float(min(ms)))), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    core.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
