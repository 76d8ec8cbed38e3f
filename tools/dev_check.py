"""Developer check: CUDA path vs oracle on seeded problems (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nrslam_b200
from nrslam_b200 import synth, api
import oracle_lib

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

core = api.Core()
print("device", core.device_info())
orc = oracle_lib.Oracle()
which = sys.argv[1:] or ["po", "pd", "ba"]
for cfg, n in (("c1", None), ("c2", None)):
    p = synth.tracking_problem(cfg, n=n)
    if "po" in which:
        t = time.time(); ro = orc.pose_only(p['cam'], p['uv'], p['X_rest'], p['seed_pose']); tc = time.time() - t
        t = time.time(); rg = core.pose_only(p['cam'], p['uv'], p['X_rest'], p['seed_pose']); tg = time.time() - t
        print(cfg, "pose_only cpu %.4fs gpu %.4fs (gpu_ms %.3f) pose diff %.3e inlier mismatches %d  iters %d/%d trials %d/%d" % (
            tc, tg, rg['stats']['gpu_ms'], np.abs(ro['pose'] - rg['pose']).max(), int((ro['inliers'] != rg['inliers']).sum()),
            ro['stats']['lm_iterations'], rg['stats']['lm_iterations'], ro['stats']['lm_trials'], rg['stats']['lm_trials']))
        tr_o, tr_g = np.array(ro['stats']['chi2_trace']), np.array(rg['stats']['chi2_trace'])
        m = min(len(tr_o), len(tr_g))
        print("   chi2 trace rel diff", rel(tr_g[:m], tr_o[:m]), "pcg iters", rg['stats']['pcg_iterations'])
        seed2 = ro['pose']
    else:
        seed2 = p['seed_pose']
    if "pd" in which:
        go, gg = p['graph'].copy(), p['graph'].copy()
        t = time.time(); ro = orc.pose_deform(p['cam'], p['uv'], p['X_rest'], p['point_vertex'], p['vertex_frame_status'], go, p['scale'], seed2, p['last_world_position']); tc = time.time() - t
        t = time.time(); rg = core.pose_deform(p['cam'], p['uv'], p['X_rest'], p['point_vertex'], p['vertex_frame_status'], gg, p['scale'], seed2, p['last_world_position']); tg = time.time() - t
        so, sg = ro['stats'], rg['stats']
        print(cfg, "pose_deform cpu %.3fs gpu %.3fs (gpu_ms %.2f stage_ms %.2f) pose diff %.3e def diff %.3e (max |d| %.3e) status mism %d lost equal %s lastpos diff %.3e" % (
            tc, tg, sg['gpu_ms'], sg['stage_ms'], np.abs(ro['pose'] - rg['pose']).max(), np.abs(ro['deformation'] - rg['deformation']).max(),
            np.abs(ro['deformation']).max(), int((ro['status'] != rg['status']).sum()), np.array_equal(ro['lost'], rg['lost']),
            np.abs(ro['last_pos'] - rg['last_pos']).max()))
        print("   iters %d/%d trials %d/%d pcg %d pairs %d/%d fixed %d/%d chi2 diff %.3e graph w diff %.3e status eq %s" % (
            so['lm_iterations'], sg['lm_iterations'], so['lm_trials'], sg['lm_trials'], sg['pcg_iterations'], so['n_pair_edges'], sg['n_pair_edges'],
            so['n_fixed_edges'], sg['n_fixed_edges'], np.abs(ro['chi2'] - rg['chi2']).max(), np.abs(go.weight - gg.weight).max(), np.array_equal(go.status, gg.status)))
        tr_o, tr_g = np.array(so['chi2_trace']), np.array(sg['chi2_trace'])
        m = min(len(tr_o), len(tr_g))
        print("   chi2 trace rel diff", rel(tr_g[:m], tr_o[:m]), len(tr_o), len(tr_g))
        ts = []
        for _ in range(3):
            s = core.resolve(1); ts.append(s['gpu_ms'])
        print("   resolve gpu_ms", ts, "pcg", s['pcg_iterations'])
if "ba" in which:
    for cfg, kw in (("c1", {}), ("c3", dict(n=1000, n_kf=8, run=5))):
        p = synth.ba_problem(cfg, **kw)
        t = time.time(); ro = orc.local_ba(p['cam'], p['kf_pose'], p['obs_kf'], p['obs_vertex'], p['uv'], p['X'], p['graph'], p['scale']); tc = time.time() - t
        t = time.time(); rg = core.local_ba(p['cam'], p['kf_pose'], p['obs_kf'], p['obs_vertex'], p['uv'], p['X'], p['graph'], p['scale']); tg = time.time() - t
        so, sg = ro['stats'], rg['stats']
        print(cfg, "local_ba O=%d cpu %.3fs gpu %.3fs (gpu_ms %.2f stage %.2f) pose diff %.3e X diff %.3e  springs %d/%d dampers %d/%d iters %d/%d trials %d/%d pcg %d" % (
            len(p['obs_kf']), tc, tg, sg['gpu_ms'], sg['stage_ms'], np.abs(ro['kf_pose'] - rg['kf_pose']).max(), np.abs(ro['X'] - rg['X']).max(),
            so['n_spring_edges'], sg['n_spring_edges'], so['n_damper_edges'], sg['n_damper_edges'], so['lm_iterations'], sg['lm_iterations'],
            so['lm_trials'], sg['lm_trials'], sg['pcg_iterations']))
        tr_o, tr_g = np.array(so['chi2_trace']), np.array(sg['chi2_trace'])
        m = min(len(tr_o), len(tr_g))
        print("   chi2 trace", tr_o[:m], tr_g[:m])
        ts = []
        for _ in range(3):
            s = core.resolve(2); ts.append(s['gpu_ms'])
        print("   resolve gpu_ms", ts)
