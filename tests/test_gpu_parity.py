"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerances (fp, stated per BASELINE.json north_star): the reference solves each damped system exactly in fp64 with
an fp32 camera model; the CUDA engine solves it by PCG to relative residual 1e-8 in fp64 with the same fp32 camera
model. Outputs are fp32 at the boundary.
  pose (unit quaternion, translation)  abs 2e-6          deformation / BA positions  abs 2e-5 (|values| ~ 0.1 .. 3)
  per-iteration robust chi2 trace      rel 1e-5          final per-point chi2        rel 1e-4 + abs 1e-3
Index bookkeeping (statuses, inlier flags, lost-id set, edge counts, graph statuses) is bit-exact.
"""
import numpy as np
import pytest

from nrslam_b200 import abi, api, synth

pytestmark = pytest.mark.gpu

POSE_TOL = 2e-6
PT_TOL = 2e-5
# pose+deformation tracking runs on the exact-solve engine (block L D L^T like the reference's Cholesky): measured against
# the oracle on these problems (tools/paritycheck.py): poses <= 4e-9, deformations <= 3e-7, accepted chi2 trace <= 9e-8
# relative — the bars are 10x tighter than the CG engine's (round 1) and still leave > 5x headroom
TRACK_POSE_TOL = 2e-7
TRACK_PT_TOL = 2e-6
TRACK_TRACE_RTOL = 1e-6


def check_trace(a, b, rtol=1e-5):
    ta, tb = np.array(a["stats"]["chi2_trace"]), np.array(b["stats"]["chi2_trace"])
    assert len(ta) == len(tb)
    assert np.allclose(ta, tb, rtol=rtol, atol=1e-9)
    assert a["stats"]["lm_iterations"] == b["stats"]["lm_iterations"]
    assert a["stats"]["lm_trials"] == b["stats"]["lm_trials"]


@pytest.mark.parametrize("cfg,n", [("c1", None), ("c2", None), ("c4", 1500), ("c1", 40)])
def test_pose_only(core, oracle, cfg, n):
    p = synth.tracking_problem(cfg, n=n)
    a = oracle.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
    b = core.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
    if cfg != "c4":
        assert np.abs(a["pose"] - b["pose"]).max() < POSE_TOL
        assert np.array_equal(a["inliers"], b["inliers"])
        check_trace(a, b, rtol=1e-6)
    else:
        # KannalaBrandt8: the device's atan2f / sinf / cosf differ from glibc's by ulps, so chi2 differs at ~1e-6
        # relative and the LM stagnation exit (rho == 0, levenberg.cpp:147-150) may trigger at a different
        # iteration; the converged estimate is what is compared.
        assert np.abs(a["pose"] - b["pose"]).max() < 2e-5
        assert (a["inliers"] != b["inliers"]).mean() < 0.01
        ta, tb = np.array(a["stats"]["chi2_trace"]), np.array(b["stats"]["chi2_trace"])
        assert abs(ta[-1] - tb[-1]) < 1e-4 * ta[-1]
    assert b["stats"]["kernel_launches"] == 1


def run_pd(engine, p, seed_pose=None):
    g = p["graph"].copy()
    r = engine.pose_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"], g, p["scale"],
                           p["seed_pose"] if seed_pose is None else seed_pose, p["last_world_position"])
    return r, g


@pytest.mark.parametrize("cfg,n,kw", [("c1", None, {}), ("c2", None, {}), ("c2", 700, dict(outlier_frac=0.3)),
                                      ("c1", 120, dict(extra_frac=0.0)), ("c4", 800, {})])
def test_pose_deform(core, oracle, cfg, n, kw):
    p = synth.tracking_problem(cfg, n=n, **kw)
    a, ga = run_pd(oracle, p)
    b, gb = run_pd(core, p)
    kb8 = cfg == "c4"
    assert np.abs(a["pose"] - b["pose"]).max() < (TRACK_POSE_TOL if not kb8 else 2e-5)
    assert np.abs(a["deformation"] - b["deformation"]).max() < (TRACK_PT_TOL if not kb8 else 2e-4)
    assert np.abs(a["X"] - b["X"]).max() < (TRACK_PT_TOL if not kb8 else 2e-4)
    assert np.abs(a["last_pos"] - b["last_pos"]).max() < (TRACK_PT_TOL if not kb8 else 2e-4)
    assert np.allclose(a["chi2"], b["chi2"], rtol=1e-4 if not kb8 else 1e-2, atol=1e-3 if not kb8 else 5e-2)
    assert abs(a["median"] - b["median"]) < (TRACK_PT_TOL if not kb8 else 2e-4)
    # bookkeeping: bit-exact
    if not kb8:
        assert np.array_equal(a["status"], b["status"])
        assert np.array_equal(ga.status, gb.status)
        check_trace(a, b, rtol=TRACK_TRACE_RTOL)
    else:  # fp32 transcendental ulps can flip a threshold tie; report-level check
        assert (a["status"] != b["status"]).mean() < 0.01
    assert np.array_equal(a["lost"], b["lost"])
    assert a["stats"]["n_pair_edges"] == b["stats"]["n_pair_edges"]
    if not kb8:
        assert a["stats"]["n_fixed_edges"] == b["stats"]["n_fixed_edges"]
    # graph attributes are fp32 functions of positions that agree to PT_TOL each: distances to 2 PT_TOL absolute,
    # weights exp(-d^2 / 2 sigma^2) to |dw| <= d / sigma^2 * 2 PT_TOL
    dtol = 2 * (TRACK_PT_TOL if not kb8 else 2e-4)
    assert np.allclose(ga.weight, gb.weight, rtol=1e-4, atol=dtol)
    assert np.allclose(ga.max_distance, gb.max_distance, rtol=0, atol=dtol)
    assert np.allclose(ga.min_distance, gb.min_distance, rtol=0, atol=dtol)


def test_pose_deform_with_bad_graph_edges(core, oracle):
    """BAD edges stop the neighbour walk (g2o_optimization.cc:258-261)."""
    p = synth.tracking_problem("c1", n=400)
    p["graph"].status[::9] = abi.EDGE_BAD
    a, ga = run_pd(oracle, p)
    b, gb = run_pd(core, p)
    assert a["stats"]["n_pair_edges"] == b["stats"]["n_pair_edges"]
    assert np.abs(a["pose"] - b["pose"]).max() < TRACK_POSE_TOL
    assert np.abs(a["deformation"] - b["deformation"]).max() < TRACK_PT_TOL
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["lost"], b["lost"])
    assert np.array_equal(ga.status, gb.status)


def test_tracking_frame_chain(core, oracle):
    """pose_only -> pose_deform chained like Tracking::TrackCameraAndDeformation (tracking.cc:291-330)."""
    p = synth.tracking_problem("c1")
    a0 = oracle.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
    b0 = core.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
    a, _ = run_pd(oracle, p, a0["pose"])
    b, _ = run_pd(core, p, b0["pose"])
    assert np.abs(a["pose"] - b["pose"]).max() < POSE_TOL
    assert np.abs(a["deformation"] - b["deformation"]).max() < PT_TOL
    assert np.array_equal(a["status"], b["status"])


@pytest.mark.parametrize("cfg,kw", [("c1", {}), ("c1", dict(n=150, n_kf=3, run=3)), ("c3", dict(n=800, n_kf=8, run=4)),
                                    ("c4", dict(n=600, n_kf=6, run=3))])
def test_local_ba(core, oracle, cfg, kw):
    p = synth.ba_problem(cfg, **kw)
    args = (p["cam"], p["kf_pose"], p["obs_kf"], p["obs_vertex"], p["uv"], p["X"], p["graph"], p["scale"])
    a = oracle.local_ba(*args)
    b = core.local_ba(*args)
    kb8 = cfg == "c4"
    assert a["stats"]["n_spring_edges"] == b["stats"]["n_spring_edges"]
    assert a["stats"]["n_damper_edges"] == b["stats"]["n_damper_edges"]
    assert np.abs(a["kf_pose"] - b["kf_pose"]).max() < (POSE_TOL if not kb8 else 2e-5)
    assert np.abs(a["X"] - b["X"]).max() < (PT_TOL if not kb8 else 2e-4)
    if not kb8:
        check_trace(a, b, rtol=1e-5)
    else:
        ta, tb = np.array(a["stats"]["chi2_trace"]), np.array(b["stats"]["chi2_trace"])
        assert len(ta) == len(tb) and np.allclose(ta, tb, rtol=1e-3)


def test_local_ba_too_few_keyframes(core):
    p = synth.ba_problem("c1", n=100)
    two = p["obs_kf"] < 2
    r = core.local_ba(p["cam"], p["kf_pose"][:2], p["obs_kf"][two], p["obs_vertex"][two], p["uv"][two], p["X"][two],
                      p["graph"], p["scale"])
    assert r["rc"] == 1
    assert np.array_equal(r["X"], p["X"][two]) and np.array_equal(r["kf_pose"], p["kf_pose"][:2])


def test_resolve_is_idempotent(core):
    """Re-running the staged device program on the HBM-resident inputs reproduces the same LM trajectory."""
    p = synth.tracking_problem("c1", n=300)
    b, _ = run_pd(core, p)
    s1 = core.resolve(1)
    s2 = core.resolve(1)
    assert s1["chi2_trace"] == s2["chi2_trace"]
    main = b["stats"]["chi2_trace"][: len(s1["chi2_trace"])]
    assert np.allclose(s1["chi2_trace"], main, rtol=1e-12)


def test_fused_frame_call_equals_the_two_calls(core):
    """nrslam_b200_track_pose_and_deform (tracking.cc:291-330 as one call: the host staging of the second problem
    overlaps the pose-only kernel, the seed pose stays in HBM) returns bit for bit what CameraPoseOptimization followed by
    CameraPoseAndDeformationOptimization return."""
    p = synth.tracking_problem("c2", n=600)
    a0 = core.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
    ga = p["graph"].copy()
    a1 = core.pose_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"], ga, p["scale"],
                          a0["pose"], p["last_world_position"])
    gb = p["graph"].copy()
    b0, b1 = core.track_pose_and_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"],
                                        gb, p["scale"], p["seed_pose"], p["last_world_position"])
    assert np.array_equal(a0["pose"], b0["pose"]) and np.array_equal(a0["inliers"], b0["inliers"])
    for k in ("pose", "deformation", "X", "chi2", "status", "lost", "last_pos"):
        assert np.array_equal(a1[k], b1[k]), k
    assert a1["median"] == b1["median"]
    assert np.array_equal(ga.weight, gb.weight) and np.array_equal(ga.status, gb.status)
    assert a0["stats"]["chi2_trace"] == b0["stats"]["chi2_trace"] and a1["stats"]["chi2_trace"] == b1["stats"]["chi2_trace"]


def _compare_with_full_size_fixture(core, cfg, pose_tol, pt_tol):
    """CUDA BA on a full-size BASELINE window against the ORACLE's result of the same seeded window
    (tests/golden/ba_full_<cfg>.npz, generated by tests/golden/make_ba_full.py with the oracle's converged-CG solver)."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ba_full_%s.npz" % cfg))
    p = synth.ba_problem(cfg)
    assert len(p["obs_kf"]) == int(G["n_obs"])
    args = (p["cam"], p["kf_pose"], p["obs_kf"], p["obs_vertex"], p["uv"], p["X"], p["graph"], p["scale"])
    b = core.local_ba(*args)
    tr = np.array(b["stats"]["chi2_trace"])
    # index / control-flow bookkeeping: exact
    assert b["stats"]["lm_iterations"] == int(G["lm_iterations"]) and b["stats"]["lm_trials"] == int(G["lm_trials"])
    # accepted chi2 trace 1e-5 relative, poses / points absolute (values 0.1 ... 3)
    assert len(tr) == len(G["chi2_trace"]) and np.abs(tr / G["chi2_trace"] - 1).max() < 1e-5
    d_pose = np.abs(b["kf_pose"] - G["kf_pose"]).max()
    d_pt = np.abs(b["X"][::int(G["stride"])] - G["X_sub"]).max()
    assert d_pose < pose_tol and d_pt < pt_tol, (d_pose, d_pt)
    # size-independent properties on top: monotone chi2, unit quaternions, determinism
    assert np.all(np.diff(tr) <= 0)
    assert np.allclose(np.linalg.norm(b["kf_pose"][:, :4], axis=1), 1.0, atol=1e-6)
    b2 = core.local_ba(*args)
    assert np.array_equal(b["X"], b2["X"]) and np.array_equal(b["kf_pose"], b2["kf_pose"])
    return d_pose, d_pt


def test_full_size_c3_matches_the_oracle(core):
    """BASELINE configs[2] at FULL size (5k landmarks / 30 KFs / 50k observations) against the oracle: LM iteration and
    trial counts exact, chi2 trace 1e-5, poses 5e-6, points 5e-5."""
    _compare_with_full_size_fixture(core, "c3", 5e-6, 5e-5)


def test_pose_deform_on_a_graph_store_view(core):
    """The owning graph store (nrslam_b200_graph_store_*: AddEdge batches, one resident graph across frames) drives
    CameraPoseAndDeformationOptimization like a caller-built CSR: same results bit for bit, and the call's graph refresh
    (UpdateVertex loop) lands in the store's own attribute arrays."""
    p = synth.tracking_problem("c1", n=400)
    g0 = p["graph"]
    v1 = np.repeat(np.arange(g0.n_vertices, dtype=np.int32), np.diff(g0.rowptr))
    keep = v1 < g0.col
    v1, v2 = v1[keep], g0.col[keep]
    store = api.GraphStore(g0.weight_sigma, g0.stretching_th)
    assert store.add_edges(v1, v2, p["last_world_position"][v2] - p["last_world_position"][v1]) == 0
    arrays = store.arrays()
    args = (p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"])
    a = core.pose_deform(*args, arrays, p["scale"], p["seed_pose"], p["last_world_position"])
    b = core.pose_deform(*args, store, p["scale"], p["seed_pose"], p["last_world_position"])
    for k in ("pose", "deformation", "X", "chi2", "status", "lost", "last_pos"):
        assert np.array_equal(a[k], b[k]), k
    after = store.arrays()
    for name in ("weight", "min_distance", "max_distance", "status"):
        assert np.array_equal(getattr(after, name), getattr(arrays, name)), name
    assert (after.max_distance > after.first_distance).any()   # the refresh really wrote into the store
    store.close()
