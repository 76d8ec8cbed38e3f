"""Oracle driver behaviour on seeded synthetic problems (the reference ships no tests or fixtures of its own —
SURVEY.md §4 — so these are property checks: convergence, g2o LM invariants, exact-vs-PCG solve agreement,
bookkeeping invariants)."""
import numpy as np
import pytest

from nrslam_b200 import abi, synth


def quat_angle(q1, q2):
    d = abs(float(np.dot(q1, q2)))
    return 2 * np.arccos(min(1.0, d))


def test_pose_only_recovers_pose_and_flags_outliers(oracle):
    p = synth.tracking_problem("c1", deform_amp=0.0)
    r = oracle.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
    assert r["rc"] == 0
    assert quat_angle(r["pose"][:4], p["true_pose"][:4]) < 2e-3
    assert np.abs(r["pose"][4:] - p["true_pose"][4:]).max() < 5e-3
    # gross outliers (+-20 px) are mostly rejected, inliers mostly kept
    assert (r["inliers"][p["outliers"]] == 0).mean() > 0.8
    assert (r["inliers"][~p["outliers"]] == 1).mean() > 0.95
    st = r["stats"]
    assert 3 <= st["lm_iterations"] <= 30 and st["lm_trials"] >= st["lm_iterations"]


def test_lm_chi2_is_monotone_within_a_round(oracle):
    p = synth.tracking_problem("c1")
    r = oracle.pose_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"],
                           p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
    tr = np.array(r["stats"]["chi2_trace"])
    assert len(tr) == r["stats"]["lm_iterations"]
    it = synth_iters = 10
    for a in range(0, len(tr), it):
        seg = tr[a:a + it]
        assert np.all(np.diff(seg) <= 1e-9 * seg[:-1])  # accepted chi2 never increases (levenberg.cpp:128-136)


def test_pose_deform_bookkeeping(oracle):
    p = synth.tracking_problem("c1")
    g = p["graph"].copy()
    r = oracle.pose_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"], g, p["scale"],
                           p["seed_pose"], p["last_world_position"])
    n = p["n"]
    st = r["status"]
    assert set(np.unique(st)).issubset({abi.TRACKED_WITH_3D, abi.TRACKED, abi.BAD})
    # reprojection outliers are demoted (g2o_optimization.cc:426-429)
    assert np.all(st[r["chi2"] > 5.99] != abi.TRACKED_WITH_3D)
    # accepted points moved by their deformation, gated points keep the rest position (:434-449)
    moved = np.any(r["X"] != p["X_rest"], axis=1)
    assert np.allclose(r["X"][moved], (p["X_rest"] + r["deformation"])[moved], atol=1e-6)
    # lost set: ascending graph vertices that are in the frame, not TRACKED_WITH_3D / JUST_TRIANGULATED (:264-273)
    lost = r["lost"]
    assert np.all(np.diff(lost) > 0)
    vfs = p["vertex_frame_status"]
    assert np.all(vfs[lost] == abi.TRACKED)
    # graph refresh only ever raises max / lowers min distance and marks BAD (regularization_graph.cc:89-128)
    assert np.all(g.max_distance >= p["graph"].max_distance) and np.all(g.min_distance <= p["graph"].min_distance)
    assert np.all(g.weight <= p["graph"].weight + 1e-7)
    assert r["stats"]["n_pair_edges"] > 2 * n


def test_exact_and_pcg_solves_agree(oracle):
    """The CUDA engine replaces the exact sparse LL^T by block-Jacobi PCG; on the CPU both must give the same LM
    trajectory to the stated tolerance (this is what makes the GPU parity tolerance meaningful)."""
    p = synth.tracking_problem("c1", n=300)
    args = (p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"])
    a = oracle.pose_deform(*args, p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"], pcg=0)
    b = oracle.pose_deform(*args, p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"], pcg=1)
    assert np.abs(a["pose"] - b["pose"]).max() < 1e-6
    assert np.abs(a["deformation"] - b["deformation"]).max() < 1e-5
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["lost"], b["lost"])
    ta, tb = np.array(a["stats"]["chi2_trace"]), np.array(b["stats"]["chi2_trace"])
    assert len(ta) == len(tb) and np.allclose(ta, tb, rtol=1e-5)


def test_local_ba_reduces_chi2_and_respects_window_rule(oracle):
    p = synth.ba_problem("c1", n=200)
    r = oracle.local_ba(p["cam"], p["kf_pose"], p["obs_kf"], p["obs_vertex"], p["uv"], p["X"], p["graph"], p["scale"])
    assert r["rc"] == 0
    tr = np.array(r["stats"]["chi2_trace"])
    assert len(tr) == 5 and tr[-1] < tr[0]  # optimize(5), g2o_optimization.cc:1143
    assert r["stats"]["n_spring_edges"] > 0 and r["stats"]["n_damper_edges"] > 0
    # fewer than 3 keyframes: untouched (g2o_optimization.cc:922-924)
    two = p["obs_kf"] < 2
    r2 = oracle.local_ba(p["cam"], p["kf_pose"][:2], p["obs_kf"][two], p["obs_vertex"][two], p["uv"][two],
                         p["X"][two], p["graph"], p["scale"])
    assert r2["rc"] == 1 and np.array_equal(r2["X"], p["X"][two]) and np.array_equal(r2["kf_pose"], p["kf_pose"][:2])


def test_get_edges_order_and_cut(oracle):
    p = synth.tracking_problem("c1", n=200)
    g = p["graph"].copy()
    g.status[::7] = abi.EDGE_BAD
    mw = np.float32(np.exp(-np.float32(1.5 * g.weight_sigma) ** 2 / (2 * np.float32(g.weight_sigma) ** 2)))
    for v in (0, 5, 17, g.n_vertices - 1):
        ent = oracle.graph_get_edges(g, v)
        e = g.eid[ent]
        key = list(zip(g.status[e].tolist(), (-g.weight[e]).tolist(), g.col[ent].tolist()))
        assert key == sorted(key)  # status asc, weight desc, neighbour asc
        assert np.all(g.weight[e] >= mw * (1 - 1e-6))
        assert np.all((ent >= g.rowptr[v]) & (ent < g.rowptr[v + 1]))


def test_converged_cg_solver_path_equals_the_factorisation():
    """The oracle's ORC_SOLVER=cg path (fp64 CG to a relative residual of 1e-14; used only to GENERATE the full-size BA
    fixtures, tests/golden/make_ba_full.py) reproduces the sparse-Cholesky path on the reference's own 5-keyframe
    window: same LM iteration / trial counts, identical fp32 outputs, chi2 trace to 1e-12."""
    import json
    import os
    import subprocess
    import sys
    code = (
        "import sys, json, numpy as np\\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\\n"
        "import oracle_lib\\n"
        "from nrslam_b200 import synth\\n"
        "q = synth.ba_problem('c1', n=200)\\n"
        "r = oracle_lib.Oracle().local_ba(q['cam'], q['kf_pose'], q['obs_kf'], q['obs_vertex'], q['uv'], q['X'], q['graph'], q['scale'])\\n"
        "print(json.dumps(dict(pose=r['kf_pose'].tolist(), X=r['X'].tolist(), trace=r['stats']['chi2_trace'],"
        " it=r['stats']['lm_iterations'], tr=r['stats']['lm_trials'])))\\n") % (
            os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for mode in ("chol", "cg"):
        env = dict(os.environ, ORC_SOLVER=mode)
        r = subprocess.run([sys.executable, "-c", code.replace("\\n", "\n")], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    a, b = outs
    assert a["it"] == b["it"] and a["tr"] == b["tr"]
    assert np.abs(np.array(a["pose"]) - np.array(b["pose"])).max() < 1e-7
    assert np.abs(np.array(a["X"]) - np.array(b["X"])).max() < 1e-6
    assert np.abs(np.array(a["trace"]) / np.array(b["trace"]) - 1).max() < 1e-12
