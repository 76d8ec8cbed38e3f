#!/usr/bin/env python
"""bench.py — headline benchmark of the NR-SLAM hot path on B200 (contract: see README / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): tracking frames/s on configs[1] — synthetic 640x480 pinhole, 2000 landmarks, REF mode — where
one frame (= one step) is what Tracking::TrackCameraAndDeformation runs per image (tracking.cc:291-330):
LucasKanadeTracker::Track of the 2000 points on a synthetic 640x480 image pair, CameraPoseOptimization (3 x 10 LM
iterations) and CameraPoseAndDeformationOptimization (2 x 10 LM iterations, + 10 for lost points in the end-to-end
path). The same frame without the KLT stage is reported under "without_klt". The line also carries the second quantity
of the metric, deformable-BA LM iterations/s on configs[2] (5000 landmarks / 30 keyframes / 50k observations), under
"ba".

  value  whole-job frames/s with the staged problem resident in HBM (device time of the LM kernels, CUDA events on
         the library's launch stream, L2 flushed between steps)
  e2e    the same frames through the C ABI with HOST buffers: host edge selection, H2D, kernels, D2H, host gating
  roofline     algorithmic bytes (SURVEY.md §8(d) formulas) / kernel time against the measured HBM peak
  cpu_baseline the oracle (CPU restatement of the reference algorithm) on one host core, bounded sample

N > 1 (torchrun): the path does not need a collective for tracking — every rank tracks its own camera stream
(replicas, weak scaling); value = frames of all ranks / max-over-ranks time.
--impl reference: the CPU restatement on all host cores (one independent frame per core), same metric and config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tracking_frames_per_sec"
UNIT = "frames/s"
WORKLOAD = ("configs[1]: synthetic 640x480 pinhole, 2000 landmarks, REF mode (one deformation vertex per landmark, "
            "symmetric 10-NN regularisation graph), frame = KLT track (2000 points, 640x480 pair) + pose_only(3x10 LM) "
            "+ pose_deform(2x10 LM)")


SEED = 1235


def base_config():
    """The workload description both arms print (the driver compares the two lines' `config`)."""
    return {"workload": WORKLOAD, "seed": SEED, "landmarks": 2000, "frames_per_step": 1,
            "l2_flush_between_steps": True}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, windows=None):
        """windows: [(t0, t1), ...] perf_counter intervals of the timed regions. Samples that arrived inside them are
        used; a run too short to catch one (nvidia-smi needs ~0.3 s to start) falls back to every sample taken while
        the sampler ran (warm-up and timed steps, the GPU is busy throughout) and says so in samples_in_timed_region."""
        if self.proc:
            self.proc.terminate()
        lines = list(self.lines)
        inside = [ln for (t, ln) in lines if windows and any(a <= t <= b for a, b in windows)]
        n_inside = len(inside)
        use = inside if inside else [ln for (_, ln) in lines]
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in use:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_in_timed_region": n_inside}


def algorithmic_bytes(st, n_pose_edges=0):
    """bytes one launch must move at fp32 storage / int32 indices, independent of implementation.
    B_sweep = 16 O + 60|48 V + 16 S + 24 D + 16 P + 136 F ; B_mv = 16 O + 36 V + 16 S + 24 D + 16 P + 76 F ;
    chi2 pass = B_sweep - 36 V. One launch = n_sweeps sweeps + pcg_iterations matvecs + n_chi2 passes."""
    O = st["n_reproj_edges"]
    V = st["n_points"]
    F = max(st["n_poses"], 1)
    S = st["n_pair_edges"] + st["n_spring_edges"]
    P = st["n_pair_edges"]
    D = st["n_damper_edges"]
    per_v = 60 if st["n_pair_edges"] else 48
    b_sweep = 16 * O + per_v * V + 16 * S + 24 * D + 16 * P + 136 * F
    b_mv = 16 * O + 36 * V + 16 * S + 24 * D + 16 * P + 76 * F
    b_chi = b_sweep - 36 * V
    return st["n_sweeps"] * b_sweep + st["pcg_iterations"] * b_mv + st["n_chi2_passes"] * b_chi, b_sweep, b_mv


def run_reference(args):
    """CPU arm: the oracle (restatement of the reference algorithm; the reference binary cannot be built here — no
    Eigen / OpenCV C++ in the image). The reference tracks ONE stream on ONE compute thread (frames are sequential,
    g2o's OpenMP is off: third_party/g2o/CMakeLists.txt:155), so `value` is single-stream, single-thread; the
    whole-box figure with one independent stream per host core is reported beside it under "multi_stream"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oracle_lib.build()
    for _ in range(min(args.warmup, 1)):
        _ref_frame(0)
    t0 = time.time()
    for s in range(args.steps):
        _ref_frame(s)
    dt = time.time() - t0
    val = args.steps / dt
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_ref_frame, range(cores))
        t1 = time.time()
        pool.map(_ref_frame, range(cores))
        dt_all = time.time() - t1
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (fp32 camera model)",
            "data": "synthetic",
            "config": base_config(),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "%d frames of configs[1], one stream on one thread like the reference; "
                                       "restatement of the reference algorithm (g2o LM + exact sparse Cholesky), "
                                       "not the reference binary" % args.steps},
            "multi_stream": {"value": cores / dt_all, "unit": UNIT, "cores": cores,
                             "sample": "one independent stream per host core, one frame each"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


_REF_KLT = {}


def _ref_frame(i):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from nrslam_b200 import synth
    p = synth.tracking_problem("c2", seed=SEED)  # the frame the B200 arm tracks on rank 0
    o = oracle_lib.Oracle()
    if "k" not in _REF_KLT:  # the reference image is set once per keyframe, not per frame (tracking.cc:361-367)
        im = synth.klt_pair(seed=77, n_points=2000)
        k = oracle_lib.OracleKLT()
        k.set_reference(im["ref"], im["pts"])
        _REF_KLT.update(k=k, im=im)
    im = _REF_KLT["im"]
    _REF_KLT["k"].track(im["cur"], im["pts"], im["status"])
    r = o.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
    o.pose_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"], p["graph"].copy(),
                  p["scale"], r["pose"], p["last_world_position"])
    return 0


SWEEP = [  # observations -> (camera config, landmarks, keyframes, visibility run)
    (1000, "c3", 200, 5, 5), (2000, "c3", 400, 5, 5), (5000, "c3", 1000, 10, 5), (10000, "c3", 1000, 20, 10),
    (20000, "c3", 2000, 30, 10), (50000, "c3", 5000, 30, 10), (100000, "c3", 10000, 60, 10),
    (200000, "c4", 20000, 100, 10)]


def run_sweep(args):
    """BASELINE configs[4] in REF mode (the parity-graded mode; EDG 'nodes' have no reference counterpart):
    deformable BA over 1k ... 200k observations on ONE GPU. Per size: device time of one optimize(5) launch on the
    HBM-resident staged window with the L2 flushed before every launch (256 MB write), LM iterations/s, CG iterations,
    the algorithmic bytes of SURVEY 8(d) and the fraction of the measured HBM peak; the oracle (one host core) is timed
    beside it where it finishes in seconds. Inside a launch the window is re-read hundreds of times and stays
    L2-resident up to ~100k observations — the fraction is a latency-bound figure, not a bandwidth claim."""
    import numpy as np
    import torch
    import nrslam_b200  # noqa: F401
    from nrslam_b200 import api, synth
    torch.cuda.set_device(0)
    core = api.Core()
    pk, pk_kind = peaks()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    rows = []
    for n_obs, cfg, n, n_kf, run in SWEEP:
        q = synth.ba_problem(cfg, n=n, n_kf=n_kf, run=run)
        bargs = (q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
        t0 = time.perf_counter()
        b = core.local_ba(*bargs)
        e2e_first = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        b = core.local_ba(*bargs)
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        ms = []
        for _ in range(max(3, min(args.steps, 5))):
            flush.zero_()
            torch.cuda.synchronize()
            sb = core.resolve(2)
            ms.append(sb["gpu_ms"])
        stb = dict(sb)
        stb.update({k: b["stats"][k] for k in ("n_reproj_edges", "n_points", "n_poses", "n_pair_edges",
                                                "n_spring_edges", "n_damper_edges")})
        alg, bsw, bmv = algorithmic_bytes(stb)
        launch = float(np.median(ms))
        row = {"observations": int(len(q["obs_kf"])), "landmarks": n, "keyframes": n_kf, "camera": cfg,
               "springs": b["stats"]["n_spring_edges"], "dampers": b["stats"]["n_damper_edges"],
               "launch_ms": launch, "e2e_ms": e2e_ms, "lm_iterations": sb["lm_iterations"],
               "lm_iterations_per_sec": sb["lm_iterations"] / (launch * 1e-3),
               "e2e_lm_iterations_per_sec": b["stats"]["lm_iterations"] / (e2e_ms * 1e-3),
               "pcg_iterations": sb["pcg_iterations"], "grid_ctas": sb["grid_ctas"],
               "bytes_per_sweep": bsw, "bytes_per_matvec": bmv, "algorithmic_bytes_per_launch": alg,
               "achieved_gbs": alg / (launch * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (launch * 1e-3) / 1e9 / pk["hbm_gbs"],
               "us_per_cg_iteration": 1e3 * launch / max(sb["pcg_iterations"], 1),
               "matvec_us_at_hbm_peak": bmv / (pk["hbm_gbs"] * 1e9) * 1e6}
        if n_obs <= 2000 and not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib
            t0 = time.perf_counter()
            ab = oracle_lib.Oracle().local_ba(*bargs)
            dt = time.perf_counter() - t0
            row["cpu_baseline"] = {"value": ab["stats"]["lm_iterations"] / dt, "unit": "iters/s", "cores": 1,
                                   "kind": "port", "sample": "one call, %.1f s" % dt}
        rows.append(row)
        print("sweep %7d obs: %8.2f ms  %7.1f LM it/s  %5d CG  %6.2f us/CG  frac %.4f" % (
            row["observations"], launch, row["lm_iterations_per_sec"], row["pcg_iterations"],
            row["us_per_cg_iteration"], row["frac_of_hbm_peak"]), file=sys.stderr, flush=True)
    print(json.dumps({"metric": "deformable_ba_lm_iterations_per_sec", "unit": "iters/s", "n_gpus": 1,
                      "config": {"workload": "configs[4]: BA sweep over observations, REF mode, optimize(5), one GPU",
                                 "l2_flush_between_launches": True},
                      "peak_gbs": pk["hbm_gbs"], "peak_kind": pk_kind, "sweep": rows}), flush=True)
    core.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-ba", action="store_true", help="skip the secondary BA measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sweep", action="store_true", help="configs[4]: BA sweep over 1k..200k observations (REF mode), one JSON line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)  # one step = one frame = ~0.6 s of CPU work: K + W frames stay within minutes
    args.warmup = max(args.warmup, 3)
    if args.sweep:
        return run_sweep(args)

    import numpy as np
    import torch
    import nrslam_b200  # noqa: F401
    from nrslam_b200 import api, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # one process per GPU shares the host cores: cap the library's staging threads (BA graph build) per rank
    os.environ.setdefault("NRSLAM_B200_HOST_THREADS", str(max(1, min(16, len(os.sched_getaffinity(0)) // world))))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to stdout at the first collective; stdout must carry ONE JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    opt = api.default_options()
    opt.device = local_rank
    core = api.Core(opt)
    # every rank tracks its own stream: a different seeded frame of the same shape (replicas, weak scaling)
    from nrslam_b200 import dist as nrs_dist
    p = synth.tracking_problem("c2", seed=nrs_dist.stream_seed(SEED, rank))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    im = synth.klt_pair(seed=77 + rank, n_points=2000)
    klt = api.KLT(core)
    klt.set_reference(im["ref"], im["pts"])

    # the call refreshes the caller's regularisation graph in place (as the reference does): every step gets its own
    # copy of the frame's graph, made before the timed region (a caller does not copy its map per frame)
    graphs = []

    def frame_e2e():
        # Tracking::TrackCameraAndDeformation (tracking.cc:291-301): data association, then pose-only and
        # pose+deformation as ONE C-ABI call (same results as the two calls; tests/test_gpu_parity.py)
        klt.track(im["cur"], im["pts"], im["status"])
        return core.track_pose_and_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"],
                                          graphs.pop() if graphs else p["graph"].copy(), p["scale"], p["seed_pose"],
                                          p["last_world_position"])

    # ---- end-to-end through the C ABI (host buffers in, host results out)
    sampler = ClockSampler(local_rank)   # started before the warm-up: nvidia-smi needs a few 100 ms to deliver its first line
    sampler.start()
    for _ in range(args.warmup):
        frame_e2e()
    graphs.extend(p["graph"].copy() for _ in range(args.steps))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r0, r1 = frame_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_window = (t0, t0 + e2e_s)
    # the same frames with the symbolic analysis of the exact solve redone on every frame (the bench tracks the same
    # frame again and again, so the plan cache always hits; a sequence whose tracked point set changes rebuilds it)
    os.environ["NRSLAM_B200_PLAN_CACHE"] = "0"
    frame_e2e()
    graphs.extend(p["graph"].copy() for _ in range(args.steps))
    t0c = time.perf_counter()
    for _ in range(args.steps):
        frame_e2e()
    torch.cuda.synchronize()
    e2e_cold_s = time.perf_counter() - t0c
    os.environ.pop("NRSLAM_B200_PLAN_CACHE", None)
    h2d = r0["stats"]["h2d_bytes"] + r1["stats"]["h2d_bytes"]
    d2h = r0["stats"]["d2h_bytes"] + r1["stats"]["d2h_bytes"]
    e2e_launches = r0["stats"]["kernel_launches"] + r1["stats"]["kernel_launches"]

    # ---- device-resident: re-run the staged programs (pose_only, pose_deform main rounds) on HBM-resident inputs
    has_lost = len(r1["lost"]) > 0
    for _ in range(args.warmup):
        klt.retrack()
        core.resolve(0)
        core.resolve(1)
        if has_lost:
            core.resolve(3)
    barrier()
    dev_ms = 0.0
    klt_ms = 0.0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        kms = klt.retrack()   # pyramid of the resident image + track kernel
        s0 = core.resolve(0)
        s1 = core.resolve(1)
        s3 = core.resolve(3) if has_lost else {"gpu_ms": 0.0, "lm_iterations": 0, "pcg_iterations": 0}
        klt_ms += kms
        dev_ms += kms + s0["gpu_ms"] + s1["gpu_ms"] + s3["gpu_ms"]
    barrier()
    wall_s = time.perf_counter() - wall0
    clocks = sampler.stop([e2e_window, (wall0, wall0 + wall_s)])
    # the same frames without the KLT stage (optimisation only), end to end
    def frame_opt_only():
        core.track_pose_and_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"],
                                   p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame_opt_only()
    opt_e2e_s = time.perf_counter() - t0
    t = torch.tensor([dev_ms, e2e_s * 1e3, dev_ms - klt_ms, opt_e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, dev_opt_ms_max, e2e_opt_ms_max = (float(v) for v in t)

    value = world * args.steps / (dev_ms_max * 1e-3)
    e2e_value = world * args.steps / (e2e_ms_max * 1e-3)

    # ---- roofline of the dominant kernel (the pose+deformation LM launch)
    pk, pk_kind = peaks()
    st = dict(s1)
    st.update(n_reproj_edges=r1["stats"]["n_reproj_edges"], n_points=r1["stats"]["n_points"], n_poses=1,
              n_pair_edges=r1["stats"]["n_pair_edges"], n_spring_edges=0, n_damper_edges=0)
    alg, b_sweep, b_mv = algorithmic_bytes(st)
    direct = s1.get("direct_solves", 0) > 0
    # exact-solve engine: a damped solve must read H once (one matvec's worth), write the factor once and read it in
    # the forward and the backward substitution: B_solve = B_mv + 3 * 8 * nnz(L')
    b_solve = b_mv + 3 * 8 * int(s1.get("factor_doubles", 0))
    if direct:
        alg += s1["direct_solves"] * b_solve
    ach = alg / (s1["gpu_ms"] * 1e-3) / 1e9
    traffic = None
    prof = "r02_track_direct_ncu_summary.json" if direct else "r01_lm_track_ncu_summary.json"
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full` capture (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", prof)))["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": ("nrs_track_direct_kernel" if direct else "nrs_lm_kernel") +
                " (pose+deformation launch)", "achieved": ach,
                "peak": pk["hbm_gbs"], "peak_kind": pk_kind, "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": traffic, "traffic_source": "profiles/" + prof, "algorithmic_bytes_per_launch": alg,
                "launch_ms": s1["gpu_ms"],
                "bytes_per_sweep": b_sweep, "bytes_per_matvec": b_mv, "bytes_per_exact_solve": b_solve if direct else None,
                "sweeps": s1["n_sweeps"], "matvecs": s1["pcg_iterations"], "exact_solves": s1.get("direct_solves", 0),
                "factor_doubles": int(s1.get("factor_doubles", 0)), "chi2_passes": s1["n_chi2_passes"],
                "note": "working set (factor 5.7 MB + update matrices, all L2-resident) never leaves the chip inside the "
                        "launch; the kernel is bound by the sequential pivot chain of the factorisation and by "
                        "synchronisation latency, not HBM (SURVEY.md 0.9, DESIGN.md 3b)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (fp32 camera model)", "data": "synthetic",
            "config": base_config(),
            "engine": {"pair_edges": int(r1["stats"]["n_pair_edges"]), "parallelism": "replicas x%d" % world,
                       "solver": "exact block L D L^T (multifrontal, nested dissection)" if direct
                       else "preconditioned CG, rel tol %g" % opt.pcg_rel_tol,
                       "grid_ctas": s1["grid_ctas"], "block_threads": s1["block_threads"],
                       "lost_points": int(len(r1["lost"]))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms_max / args.steps, "launches_per_step": e2e_launches,
                    "plan_reused": bool(r1["stats"].get("plan_reused", 0)),
                    "value_plan_rebuilt_every_frame": world * args.steps / e2e_cold_s,
                    "ms_per_step_plan_rebuilt_every_frame": 1e3 * e2e_cold_s / args.steps},
            # per step: KLT pyramid (1 level-0 + 4 pyrDown + 5 Scharr) + 1 track kernel, 3 LM kernels (pose-only, main
            # rounds, lost-point stage)
            "gpu_launches": (11 + 2 + (1 if has_lost else 0)) * args.steps, "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "klt": {"device_ms_per_step": klt_ms / args.steps, "points": 2000, "image": "640x480",
                    "kernels_per_step": 11},
            "without_klt": {"value": world * args.steps / (dev_opt_ms_max * 1e-3),
                            "e2e_value": world * args.steps / (e2e_opt_ms_max * 1e-3), "unit": UNIT},
            "lm_iterations_per_step": s0["lm_iterations"] + s1["lm_iterations"] + s3["lm_iterations"],
            "pcg_iterations_per_step": s0["pcg_iterations"] + s1["pcg_iterations"] + s3["pcg_iterations"],
            "device_ms": {"klt": klt_ms / args.steps, "pose_only": s0["gpu_ms"], "pose_deform_main": s1["gpu_ms"],
                          "lost_points": s3["gpu_ms"]},
            "roofline": roofline, "clocks": clocks}

    # ---- second quantity of the metric: deformable-BA LM iterations/s (one GPU)
    #   ba_window5: the window the reference actually runs (5 keyframes, g2o_optimization.cc:894; configs[0] shape:
    #               500 landmarks, 2500 observations), with the oracle timed beside it;
    #   ba:         configs[2] (5000 landmarks / 30 keyframes / 50k observations) — beyond what the CPU path finishes
    #               in minutes, GPU only.
    def ba_measure(q, tag, workload):
        b = core.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
        ks = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ks):
            b = core.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"],
                              q["scale"])
        e2e_ms = 1e3 * (time.perf_counter() - t0) / ks
        ms = 0.0
        for _ in range(ks):
            flush.zero_()
            torch.cuda.synchronize()
            sb = core.resolve(2)
            ms += sb["gpu_ms"]
        stb = dict(sb)
        stb.update({k: b["stats"][k] for k in ("n_reproj_edges", "n_points", "n_poses", "n_pair_edges",
                                                "n_spring_edges", "n_damper_edges")})
        alg_b, bsw, bmv = algorithmic_bytes(stb)
        traffic_b = None
        if tag == "c3":  # dram__bytes_read.sum + dram__bytes_write.sum of this launch, one `ncu --set full` capture
            try:
                traffic_b = json.load(open(os.path.join(ROOT, "profiles", "r02_lm_ba_wide_ncu_summary.json")))[
                    "dram_bytes_per_launch"]
            except Exception:
                pass
        return {"metric": "deformable_ba_lm_iterations_per_sec", "unit": "iters/s",
                "value": world * ks * sb["lm_iterations"] / (ms * 1e-3),
                "e2e_value": world * b["stats"]["lm_iterations"] / (e2e_ms * 1e-3),
                "workload": workload % (b["stats"]["n_reproj_edges"], b["stats"]["n_spring_edges"],
                                        b["stats"]["n_damper_edges"]),
                "launch_ms": ms / ks, "e2e_ms": e2e_ms, "lm_iterations": sb["lm_iterations"],
                "pcg_iterations": sb["pcg_iterations"], "grid_ctas": sb["grid_ctas"],
                "roofline": {"bound": "hbm", "achieved": alg_b / (ms / ks * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                             "unit": "GB/s", "frac": alg_b / (ms / ks * 1e-3) / 1e9 / pk["hbm_gbs"],
                             "traffic": traffic_b, "algorithmic_bytes_per_launch": alg_b,
                             "bytes_per_sweep": bsw, "bytes_per_matvec": bmv}}

    if not args.no_ba:
        q5 = synth.ba_problem("c1")
        line["ba_window5"] = ba_measure(q5, "w5", "the reference's own window: 960x720 pinhole, 500 landmarks / 5 "
                                        "keyframes / %d observations, %d springs, %d dampers, optimize(5)")
        q3 = synth.ba_problem("c3")
        line["ba"] = ba_measure(q3, "c3", "configs[2]: 640x480 pinhole, 5000 landmarks / 30 "
                                "keyframes / %d observations, %d springs, %d dampers, optimize(5)")
        single3 = core.local_ba(q3["cam"], q3["kf_pose"], q3["obs_kf"], q3["obs_vertex"], q3["uv"], q3["X"],
                                q3["graph"], q3["scale"]) if world > 1 else None
        # the north-star target window: configs[3], Endomapper-shaped 1440x1080 fisheye, 20k landmarks / 100 keyframes
        q4 = synth.ba_problem("c4")
        line["ba_c4"] = ba_measure(q4, "c4", "configs[3]: 1440x1080 KannalaBrandt8, 20000 landmarks / 100 "
                                   "keyframes / %d observations, %d springs, %d dampers, optimize(5)")
        if world > 1:
            # ---- the path's one real exchange: ONE window landmark-sharded over all ranks — halo rows and
            # partial sums cross NVLink inside the persistent kernels (strong scaling of one problem, reported beside
            # the replica numbers). The G-GPU == 1-GPU gate is evaluated HERE, in the driver-run bench: every rank also
            # solves the window alone and the sharded result is compared with it.
            single4 = core.local_ba(q4["cam"], q4["kf_pose"], q4["obs_kf"], q4["obs_vertex"], q4["uv"], q4["X"],
                                    q4["graph"], q4["scale"])
            parts = {}
            for tag, q in (("c3", q3), ("c4", q4)):
                parts[tag] = api.shard_partition(world, q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"],
                                                 q["graph"], q["scale"])
            max_rows = max(int((pt["n_own"] + pt["n_halo"]).max()) for pt in parts.values())
            nrs_dist.attach_shards(core, dist, max_rows, max(len(q3["kf_pose"]), len(q4["kf_pose"])))

            def sharded_measure(q, part, single, single_ms, label):
                bargs = (q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
                core.local_ba_sharded(*bargs)
                ks = max(2, min(args.steps, 5))
                sh_ms, sh_wall = 0.0, 0.0
                for _ in range(ks):
                    barrier()
                    t0 = time.perf_counter()
                    rs = core.local_ba_sharded(*bargs)
                    sh_wall += 1e3 * (time.perf_counter() - t0)
                    sh_ms += rs["stats"]["gpu_ms"]
                tt = torch.tensor([sh_ms, sh_wall], dtype=torch.float64, device="cuda")
                per_rank = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
                dist.all_gather(per_rank, tt)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                # equality gate: poses identical on every rank, sharded == single-GPU within fp32 output precision
                X = nrs_dist.gather_sharded_ba(rs, dist)
                poses = [None] * world
                dist.all_gather_object(poses, rs["kf_pose"])
                its = rs["stats"]["lm_iterations"]
                ta, tb_ = np.array(single["stats"]["chi2_trace"]), np.array(rs["stats"]["chi2_trace"])
                return {
                    "metric": "deformable_ba_lm_iterations_per_sec", "unit": "iters/s", "scaling": "strong",
                    "value": ks * its / (float(tt[0]) * 1e-3), "e2e_value": ks * its / (float(tt[1]) * 1e-3),
                    "workload": "ONE %s window (%d observations, %d keyframes) landmark-sharded over %d GPUs: %s rows + "
                                "%s halo rows per rank" % (label, len(q["obs_kf"]), len(q["kf_pose"]), world,
                                                           part["n_own"].tolist(), part["n_halo"].tolist()),
                    "launch_ms": float(tt[0]) / ks, "e2e_ms": float(tt[1]) / ks, "lm_iterations": its,
                    "pcg_iterations": rs["stats"]["pcg_iterations"], "single_gpu_launch_ms": single_ms,
                    "per_rank_launch_ms": [float(v[0]) / ks for v in per_rank],
                    "per_rank_call_ms": [float(v[1]) / ks for v in per_rank],
                    "poses_identical": bool(all(np.array_equal(poses[0], q_) for q_ in poses)),
                    "max_pose_diff": float(np.abs(rs["kf_pose"] - single["kf_pose"]).max()),
                    "max_point_diff": float(np.abs(X - single["X"]).max()),
                    "chi2_trace_equal_1e-6": bool(len(ta) == len(tb_) and np.allclose(ta, tb_, rtol=1e-6)),
                    "pcg_iterations_single": single["stats"]["pcg_iterations"]}

            line["ba_sharded"] = sharded_measure(q4, parts["c4"], single4, line["ba_c4"]["launch_ms"], "configs[3]")
            line["ba_sharded_c3"] = sharded_measure(q3, parts["c3"], single3, line["ba"]["launch_ms"], "configs[2]")

    # ---- one frame's batch of DeformableTriangulation candidates (Mapping::FrameMapping,
    # mapping.cc:60-113) and the RegularizationGraph update loop of the tracking frame (g2o_optimization.cc:458-474)
    tb = synth.triangulation_batch(seed=21, n_cand=600, fail_frac=0.1)
    tri = api.Triangulator(core)
    tri.run_batch(tb)
    ks = max(3, min(args.steps, 10))
    tri_dev = sum(tri.rerun() for _ in range(ks)) / ks
    t0 = time.perf_counter()
    for _ in range(ks):
        rt = tri.run_batch(tb)
    tri_e2e = 1e3 * (time.perf_counter() - t0) / ks
    line["triangulation"] = {
        "metric": "deformable_triangulations_per_sec", "unit": "candidates/s",
        "workload": "600 candidates of one 640x480 frame, tracks of 5..20 frames, <= 11 neighbours, optimize(10) each "
                    "(%d succeed)" % int((rt["status"] == 0).sum()),
        "value": world * tb["n_cand"] / (tri_dev * 1e-3), "e2e_value": world * tb["n_cand"] / (tri_e2e * 1e-3),
        "launch_ms": tri_dev, "e2e_ms": tri_e2e, "lm_iterations": int(rt["lm_iterations"].sum()),
        "grid_ctas": tb["n_cand"], "block_threads": 128}
    gq = p["graph"].copy()
    gverts = np.unique(p["point_vertex"]).astype(np.int32)
    gpos = p["last_world_position"] + synth.smooth_field(np.random.default_rng(3), p["last_world_position"], 0.05)
    core.graph_update_vertices(gq.copy(), gverts, gpos)
    t0 = time.perf_counter()
    for _ in range(ks):
        core.graph_update_vertices(gq.copy(), gverts, gpos)
    gu_e2e = 1e3 * (time.perf_counter() - t0) / ks
    t0 = time.perf_counter()
    for _ in range(ks):
        gh = gq.copy()
        for v in gverts:
            core.graph_update_vertex(gh, int(v), gpos)
    gu_host = 1e3 * (time.perf_counter() - t0) / ks
    line["graph_update"] = {"vertices": int(len(gverts)), "edges": int(gq.n_edges), "e2e_ms": gu_e2e,
                            "per_vertex_host_entry_point_ms": gu_host,
                            "note": "UpdateVertex loop of one tracking frame through the C ABI with host buffers "
                                    "(H2D of the graph arrays + one kernel + D2H); the second figure is the sequential "
                                    "host entry point called per vertex from Python (ctypes overhead included)"}

    # ---- CPU baseline: the oracle on one host core, bounded sample (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        t0 = time.perf_counter()
        oracle_lib.deformable_triangulation(tb)
        dtt = time.perf_counter() - t0
        line["triangulation"]["cpu_baseline"] = {"value": tb["n_cand"] / dtt, "unit": "candidates/s", "cores": 1,
                                                 "kind": "port", "sample": "the same 600 candidates once (%.2f s)" % dtt}
        orc = oracle_lib.Oracle()
        oklt = oracle_lib.OracleKLT()
        oklt.set_reference(im["ref"], im["pts"])
        n_frames = 0
        t0 = time.perf_counter()
        while n_frames < 3 or (time.perf_counter() - t0 < 10.0 and n_frames < 20):
            oklt.track(im["cur"], im["pts"], im["status"])
            a0 = orc.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
            orc.pose_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"],
                            p["graph"].copy(), p["scale"], a0["pose"], p["last_world_position"])
            n_frames += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_frames / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "%d frames of the same configs[1] problem, single thread (the reference "
                                          "is single-threaded: g2o OpenMP off); restatement of the reference "
                                          "algorithm, not the reference binary" % n_frames,
                                "host_cores_available": len(os.sched_getaffinity(0))}
        if not args.no_ba:   # one call of the 5-keyframe window (bounded: ~10 s)
            t0 = time.perf_counter()
            ab = orc.local_ba(q5["cam"], q5["kf_pose"], q5["obs_kf"], q5["obs_vertex"], q5["uv"], q5["X"], q5["graph"],
                              q5["scale"])
            dtb = time.perf_counter() - t0
            line["ba_window5"]["cpu_baseline"] = {"value": ab["stats"]["lm_iterations"] / dtb, "unit": "iters/s",
                                                  "cores": 1, "kind": "port",
                                                  "sample": "one LocalDeformableBundleAdjustment call on the same "
                                                            "window (%d LM iterations, %.1f s)" % (
                                                                ab["stats"]["lm_iterations"], dtb)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    klt.close()
    tri.close()
    core.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
