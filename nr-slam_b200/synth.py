"""Seeded synthetic problems of SURVEY.md §8(d) (the reference bundles no images / trajectories / goldens).

Scene: landmarks on a wavy sheet z = z0 + a sin(fx x) cos(fy y) at depth z0 = 3 (the reference normalises the median
depth to 3, tracking.cc:153-157); smooth low-frequency deformation; sigma = 0.5 px noise (g2o_optimization.cc:203);
10 % outlier observations; symmetric k-NN regularisation graph with sigma_w = 3 std(depth) scale
(tracking.cc:159-160,200), first_distance = rest distance, all edges NEUTRAL.
"""
import numpy as np
from scipy.spatial import cKDTree

from .abi import (EDGE_NEUTRAL, JUST_TRIANGULATED, TRACKED, TRACKED_WITH_3D, Camera, GraphArrays)

CONFIGS = {
    # name: camera, image size, N landmarks, F keyframes, L visibility run
    "c1": dict(cam=("pinhole", 472.64955, 472.64955, 479.5, 359.5), size=(960, 720), n=500, kf=5, run=5),
    "c2": dict(cam=("pinhole", 520.0, 520.0, 320.0, 240.0), size=(640, 480), n=2000, kf=0, run=0),
    "c3": dict(cam=("pinhole", 747.2929, 747.2929, 320.0, 240.0), size=(640, 480), n=5000, kf=30, run=10),
    "c4": dict(cam=("kb8", 717.2104, 717.4816, 735.3566, 552.7982, -0.1389272, -0.001239606, 0.0009125824,
                    -4.071615e-05), size=(1440, 1080), n=20000, kf=100, run=10),
}


def make_camera(spec):
    return Camera.pinhole(*spec[1:]) if spec[0] == "pinhole" else Camera.kb8(*spec[1:])


def project(cam, P):
    """fp32 projection of camera-frame points (same formulas as calibration/*.cc) — generator use only."""
    P = np.asarray(P, np.float32)
    p = np.array(cam.params[:], np.float32)
    if cam.model == 0:
        return np.stack([p[0] * P[:, 0] / P[:, 2] + p[2], p[1] * P[:, 1] / P[:, 2] + p[3]], 1)
    r2 = P[:, 0] ** 2 + P[:, 1] ** 2
    th = np.arctan2(np.sqrt(r2), P[:, 2])
    psi = np.arctan2(P[:, 1], P[:, 0])
    r = th + p[4] * th ** 3 + p[5] * th ** 5 + p[6] * th ** 7 + p[7] * th ** 9
    return np.stack([p[0] * r * np.cos(psi) + p[2], p[1] * r * np.sin(psi) + p[3]], 1).astype(np.float32)


def unproject_plane(cam, uv, z):
    """Rays through pixels scaled to depth z (pinhole exact; KB8 by Newton on theta) — generator use only."""
    p = np.array(cam.params[:], np.float64)
    x = (uv[:, 0] - p[2]) / p[0]
    y = (uv[:, 1] - p[3]) / p[1]
    if cam.model == 0:
        return np.stack([x * z, y * z, np.full_like(x, z)], 1)
    rd = np.hypot(x, y)
    th = rd.copy()
    for _ in range(20):
        f = th + p[4] * th ** 3 + p[5] * th ** 5 + p[6] * th ** 7 + p[8 - 1] * th ** 9 - rd
        fd = 1 + 3 * p[4] * th ** 2 + 5 * p[5] * th ** 4 + 7 * p[6] * th ** 6 + 9 * p[7] * th ** 8
        th = th - f / fd
    s = np.where(rd > 1e-9, np.tan(th) / np.maximum(rd, 1e-9), 1.0)
    return np.stack([x * s * z, y * s * z, np.full_like(x, z)], 1)


def rotvec_to_quat(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.array([0, 0, 0, 1.0])
    a = w / th
    return np.concatenate([a * np.sin(th / 2), [np.cos(th / 2)]])


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_pose(rng, max_deg, max_t):
    w = rng.normal(size=3)
    w *= np.deg2rad(max_deg) * rng.uniform(0.3, 1.0) / np.linalg.norm(w)
    t = rng.normal(size=3)
    t *= max_t * rng.uniform(0.3, 1.0) / np.linalg.norm(t)
    q = rotvec_to_quat(w)
    return np.concatenate([q, t]).astype(np.float32)


def sheet_points(rng, cam, size, n, z0=3.0, amp=0.3, margin=0.06):
    """n landmarks whose projections fill the image, lying on the wavy sheet (world frame == seed camera frame)."""
    w, h = size
    uv = np.stack([rng.uniform(margin * w, (1 - margin) * w, n), rng.uniform(margin * h, (1 - margin) * h, n)], 1)
    P = unproject_plane(cam, uv, z0)
    ext = max(np.ptp(P[:, 0]), np.ptp(P[:, 1]))
    fx, fy = 2 * np.pi * 1.5 / ext, 2 * np.pi * 1.0 / ext
    zz = z0 + amp * np.sin(fx * P[:, 0]) * np.cos(fy * P[:, 1])
    P = P * (zz / z0)[:, None]
    return P.astype(np.float32)


def smooth_field(rng, P, amp):
    """Low-frequency displacement field sampled at points P."""
    ext = max(np.ptp(P[:, 0]), np.ptp(P[:, 1]), 1e-6)
    out = np.zeros_like(P, dtype=np.float64)
    for _ in range(3):
        k = rng.normal(size=3) * 2 * np.pi * 0.6 / ext
        ph = rng.uniform(0, 2 * np.pi)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        out += np.outer(np.sin(P @ k + ph), d)
    return (out * amp / np.sqrt(3)).astype(np.float32)


def knn_graph(P, k, weight_sigma, stretching_th=1.1):
    """Symmetric k-NN graph over points P in CSR form (rows ascending = ascending map-point id)."""
    n = len(P)
    tree = cKDTree(P.astype(np.float64))
    kk = min(k + 1, n)
    _, nb = tree.query(P.astype(np.float64), k=kk)
    a = np.repeat(np.arange(n), kk - 1)
    b = nb[:, 1:].reshape(-1)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = np.unique(lo.astype(np.int64) * n + hi)
    ei, ej = (key // n).astype(np.int32), (key % n).astype(np.int32)
    E = len(ei)
    rel = P[ej].astype(np.float32) - P[ei].astype(np.float32)
    dist = np.sqrt((rel[:, 0] * rel[:, 0] + rel[:, 1] * rel[:, 1]) + rel[:, 2] * rel[:, 2]).astype(np.float32)
    s = np.float32(weight_sigma)
    weight = np.exp(-(dist * dist) / (np.float32(2) * s * s)).astype(np.float32)
    rows = np.concatenate([ei, ej])
    cols = np.concatenate([ej, ei])
    eids = np.concatenate([np.arange(E), np.arange(E)]).astype(np.int32)
    order = np.lexsort((cols, rows))
    rows, cols, eids = rows[order], cols[order], eids[order]
    rowptr = np.zeros(n + 1, np.int32)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return GraphArrays(rowptr, cols, eids, weight, dist, dist.copy(), dist.copy(),
                       np.full(E, EDGE_NEUTRAL, np.uint8), weight_sigma, stretching_th)


def tracking_problem(config="c2", seed=None, n=None, extra_frac=0.08, outlier_frac=0.10, knn=10,
                     deform_amp=0.02, noise_px=0.5, max_deg=2.0, max_t=0.02):
    """One frame of pose(+deformation) tracking. Returns a dict of numpy arrays (+ 'cam', 'graph')."""
    cfg = CONFIGS[config]
    if seed is None:
        seed = 1234 + list(CONFIGS).index(config)
    rng = np.random.default_rng(seed)
    cam = make_camera(cfg["cam"])
    n = n or cfg["n"]
    m = n + int(round(n * extra_frac))  # extra map points: lost / just-triangulated / not in frame
    z0 = 3.0
    P = sheet_points(rng, cam, cfg["size"], m, z0)
    scale = np.float32(3.0 / np.median(P[:, 2]))
    weight_sigma = np.float32(3.0 * np.std(P[:, 2]) * scale)
    graph = knn_graph(P, knn, weight_sigma)
    # frame statuses for every map point (graph vertex order == ascending map-point id)
    vfs = np.full(m, -1, np.int8)
    perm = rng.permutation(m)
    tracked3d = np.sort(perm[:n])
    rest = perm[n:]
    third = len(rest) // 3
    vfs[tracked3d] = TRACKED_WITH_3D
    vfs[rest[:third]] = TRACKED               # in frame without 3-D -> "lost" neighbours
    vfs[rest[third:2 * third]] = JUST_TRIANGULATED
    # rest[2*third:] stay -1 (not in the frame)
    # the frame lists its TRACKED_WITH_3D points in frame-index order; use a shuffled order to exercise indexing
    order = rng.permutation(n)
    point_vertex = tracked3d[order].astype(np.int32)
    X_rest = P[point_vertex]
    # ground truth: camera motion + smooth deformation
    seed_pose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    true_pose = make_pose(rng, max_deg, max_t * z0)
    D = smooth_field(rng, P, deform_amp * z0)
    Xd = (X_rest + D[point_vertex]).astype(np.float64)
    Rm = quat_to_R(true_pose[:4].astype(np.float64))
    Pc = Xd @ Rm.T + true_pose[4:].astype(np.float64)
    uv = project(cam, Pc).astype(np.float64)
    uv += rng.normal(scale=noise_px, size=uv.shape)
    out = rng.random(n) < outlier_frac
    uv[out] += rng.uniform(-20, 20, size=(int(out.sum()), 2))
    return dict(cam=cam, n=n, m=m, uv=uv.astype(np.float32), X_rest=X_rest.astype(np.float32),
                point_vertex=point_vertex, vertex_frame_status=vfs, graph=graph, scale=float(scale),
                seed_pose=seed_pose, true_pose=true_pose, last_world_position=P.astype(np.float32).copy(),
                outliers=out, config=config, seed=seed, size=cfg["size"])


def ba_problem(config="c1", seed=None, n=None, n_kf=None, run=None, knn=10, deform_amp=0.01, noise_px=0.5,
               outlier_frac=0.02, max_deg=1.5, max_t=0.015):
    """A keyframe window for LocalDeformableBundleAdjustment: each landmark visible in a contiguous run of KFs."""
    cfg = CONFIGS[config]
    if seed is None:
        seed = 2234 + list(CONFIGS).index(config)
    rng = np.random.default_rng(seed)
    cam = make_camera(cfg["cam"])
    n = n or cfg["n"]
    F = n_kf or cfg["kf"]
    L = min(run or cfg["run"], F)
    z0 = 3.0
    P = sheet_points(rng, cam, cfg["size"], n, z0, margin=0.12)
    scale = np.float32(3.0 / np.median(P[:, 2]))
    weight_sigma = np.float32(3.0 * np.std(P[:, 2]) * scale)
    graph = knn_graph(P, knn, weight_sigma)
    # smooth trajectory and cumulative deformation
    poses_true = [np.array([0, 0, 0, 1, 0, 0, 0], np.float64)]
    Rm, t = np.eye(3), np.zeros(3)
    for _ in range(1, F):
        dp = make_pose(rng, max_deg, max_t * z0).astype(np.float64)
        dR = quat_to_R(dp[:4])
        Rm, t = dR @ Rm, dR @ t + dp[4:]
        # quaternion of Rm
        from scipy.spatial.transform import Rotation
        q = Rotation.from_matrix(Rm).as_quat()
        if q[3] < 0:
            q = -q
        poses_true.append(np.concatenate([q, t]))
    start = rng.integers(0, F - L + 1, size=n) if F > L else np.zeros(n, np.int64)
    obs_kf, obs_vertex, uv_l, X_l = [], [], [], []
    Dcum = np.zeros_like(P, dtype=np.float64)
    kf_pose = []
    for k in range(F):
        Dcum = Dcum + smooth_field(rng, P, deform_amp * z0)
        vis = np.nonzero((start <= k) & (k < start + L))[0]
        vis = vis[rng.permutation(len(vis))]  # keyframe index order is arbitrary w.r.t. map-point id
        Xk = P[vis].astype(np.float64) + Dcum[vis]
        Rk = quat_to_R(poses_true[k][:4])
        Pc = Xk @ Rk.T + poses_true[k][4:]
        uvk = project(cam, Pc).astype(np.float64) + rng.normal(scale=noise_px, size=(len(vis), 2))
        o = rng.random(len(vis)) < outlier_frac
        uvk[o] += rng.uniform(-20, 20, size=(int(o.sum()), 2))
        obs_kf.append(np.full(len(vis), k, np.int32))
        obs_vertex.append(vis.astype(np.int32))
        uv_l.append(uvk.astype(np.float32))
        # stored per-KF positions: truth + estimation noise ; stored pose: truth perturbed
        X_l.append((Xk + rng.normal(scale=0.004 * z0, size=Xk.shape)).astype(np.float32))
        pert = make_pose(rng, 0.3, 0.003 * z0).astype(np.float64)
        dR = quat_to_R(pert[:4])
        from scipy.spatial.transform import Rotation
        q = Rotation.from_matrix(dR @ Rk).as_quat()
        if q[3] < 0:
            q = -q
        kf_pose.append(np.concatenate([q, dR @ poses_true[k][4:] + pert[4:]]).astype(np.float32))
    return dict(cam=cam, n_kf=F, kf_pose=np.stack(kf_pose).astype(np.float32), obs_kf=np.concatenate(obs_kf),
                obs_vertex=np.concatenate(obs_vertex), uv=np.concatenate(uv_l), X=np.concatenate(X_l),
                graph=graph, scale=float(scale), config=config, seed=seed, n=n,
                poses_true=np.stack(poses_true).astype(np.float32))


def _texture(rng, h, w):
    """Band-limited random texture in [0, 1] (separable box smoothing of white noise at two scales; numpy only)."""
    def smooth(a, r):
        k = np.ones(2 * r + 1) / (2 * r + 1)
        for _ in range(3):  # three box passes ~ Gaussian
            a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, mode="valid"), 0, a)
            a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, mode="valid"), 1, a)
        return a
    n = rng.normal(size=(h, w))
    t = smooth(n, 1) * 3 + smooth(n, 4) * 8
    return (t - t.min()) / (t.max() - t.min())


def klt_pair(seed=1, size=(640, 480), n_points=500, shift=(2.6, -1.7), gain=1.1, bias=-8.0, noise=1.0, margin=40):
    """A reference / current image pair for the KLT tracker: band-limited texture, the current
    image is the reference translated by a sub-pixel shift (bilinear resampling), with affine gain/bias and noise.
    Returns ref, cur (uint8 HxW), pts (reference keypoints), pts_true, status (all TRACKED)."""
    from .abi import TRACKED
    rng = np.random.default_rng(seed)
    w, h = size
    pad = 16
    big = _texture(rng, h + 2 * pad, w + 2 * pad) * 200 + 25
    ref = big[pad:pad + h, pad:pad + w]
    dx, dy = shift
    # cur(x, y) = ref(x - dx, y - dy): a point at p in ref appears at p + shift in cur
    xs = np.arange(w) - dx + pad
    ys = np.arange(h) - dy + pad
    x0, y0 = np.floor(xs).astype(int), np.floor(ys).astype(int)
    fx, fy = (xs - x0)[None, :], (ys - y0)[:, None]
    g = big
    cur = ((1 - fy) * ((1 - fx) * g[np.ix_(y0, x0)] + fx * g[np.ix_(y0, x0 + 1)]) +
           fy * ((1 - fx) * g[np.ix_(y0 + 1, x0)] + fx * g[np.ix_(y0 + 1, x0 + 1)]))
    cur = gain * cur + bias + rng.normal(scale=noise, size=cur.shape) if noise > 0 else gain * cur + bias
    pts = np.stack([rng.uniform(margin, w - margin, n_points), rng.uniform(margin, h - margin, n_points)], 1)
    return dict(ref=np.clip(np.rint(ref), 0, 255).astype(np.uint8), cur=np.clip(np.rint(cur), 0, 255).astype(np.uint8),
                pts=pts.astype(np.float32), pts_true=(pts + np.array([dx, dy])).astype(np.float32),
                status=np.full(n_points, TRACKED, np.uint8), shift=shift)


def _field_modes(rng, ext, n_modes=3):
    modes = []
    for _ in range(n_modes):
        k = rng.normal(size=3) * 2 * np.pi * 0.6 / ext
        ph = rng.uniform(0, 2 * np.pi)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        modes.append((k, ph, d))
    return modes


def _eval_field(modes, P, amp):
    out = np.zeros((len(P), 3))
    for k, ph, d in modes:
        out += np.outer(np.sin(P @ k + ph), d)
    return out * amp / np.sqrt(len(modes))


def triangulation_batch(seed=7, n_cand=200, cam_spec=CONFIGS["c2"]["cam"], size=CONFIGS["c2"]["size"], t_min=5,
                        t_max=20, n_map=130, noise_px=0.5, dropout=0.08, deform_amp=0.03, baseline=0.35,
                        fail_frac=0.12, nb_slots=12):
    """One frame's worth of DeformableTriangulation candidates (g2o_optimization.cc:559-814) as the flattened
    TemporalBuffer view of include/nrslam_b200.h: a deforming sheet observed by a camera translating sideways over
    t_max frames; map points (the neighbours) carry their per-frame world positions; a candidate is an untriangulated
    feature tracked over the last T in [t_min, t_max] frames. A fraction of the candidates is built to fail each of the
    reference's checks (no neighbour / close feature, gross keypoint error in the first or last frame, static camera,
    neighbours missing in one frame, neighbours behind the camera, incoherent neighbour flows, a common drift of the
    neighbours that drags the estimates off their rays)."""
    rng = np.random.default_rng(seed)
    cam = make_camera(cam_spec)
    w, h = size
    F = t_max
    # trajectory: camera_transform_world per frame
    poses = np.zeros((F, 7), np.float32)
    axis = rng.normal(size=3)
    axis[2] *= 0.2
    axis /= np.linalg.norm(axis)
    wrot = rng.normal(size=3)
    wrot *= np.deg2rad(3.0) / np.linalg.norm(wrot)
    for f in range(F):
        s = f / max(F - 1, 1)
        q = rotvec_to_quat(wrot * s)
        poses[f] = np.concatenate([q, axis * baseline * s + 0.002 * rng.normal(size=3)]).astype(np.float32)
    P_map = sheet_points(rng, cam, size, n_map).astype(np.float64)
    ext = max(np.ptp(P_map[:, 0]), np.ptp(P_map[:, 1]))
    m1, m2 = _field_modes(rng, ext), _field_modes(rng, ext)

    def positions(P, f):
        s = f / max(F - 1, 1)
        return P + _eval_field(m1, P, deform_amp) * s + _eval_field(m2, P, deform_amp) * np.sin(2.0 * s)

    def to_cam(pose, X):
        return X @ quat_to_R(pose[:4].astype(np.float64)).T + pose[4:].astype(np.float64)

    map_world = np.stack([positions(P_map, f) for f in range(F)]).astype(np.float32)  # F x n_map x 3
    map_uv_last = project(cam, to_cam(poses[F - 1], map_world[F - 1].astype(np.float64)))
    tree = cKDTree(map_uv_last.astype(np.float64))

    track_ptr = [0]
    uv_l, pose_l, pos_l, val_l, nnb_l, kind_l, truth_l = [], [], [], [], [], [], []
    kinds = ["too_close", "bad_first", "bad_last", "static", "no_nb", "behind", "noisy_nb", "drift"]
    pool = sheet_points(rng, cam, size, 40 * n_cand + 400).astype(np.float64)
    for ip in range(len(pool)):
        if len(nnb_l) >= n_cand:
            break
        Pc = pool[ip:ip + 1]
        T = int(rng.integers(t_min, t_max + 1))
        frames = np.arange(F - T, F)
        kind = "ok"
        if rng.uniform() < fail_frac:
            kind = kinds[int(rng.integers(len(kinds)))]
        Xc = np.stack([positions(Pc, f)[0] for f in frames])
        cposes = poses[frames].copy()
        if kind == "static":
            cposes[:] = cposes[-1]
            Xc[:] = Xc[-1]
        uv = np.stack([project(cam, to_cam(cposes[i], Xc[i:i + 1]))[0] for i in range(T)]).astype(np.float64)
        uv += rng.normal(size=uv.shape) * noise_px
        if not (np.all(uv[:, 0] > 5) and np.all(uv[:, 0] < w - 5) and np.all(uv[:, 1] > 5) and np.all(uv[:, 1] < h - 5)):
            continue
        if kind == "bad_first":
            uv[0] += rng.choice([-1, 1], 2) * rng.uniform(15, 40, 2)
        if kind == "bad_last":
            uv[-1] += rng.choice([-1, 1], 2) * rng.uniform(15, 40, 2)
        # GetClosestMapPointsToFeature(candidate, 10, 20, 500) on the last snapshot (temporal_buffer.cc:97-143)
        d_all, idx_all = tree.query(uv[-1], k=min(n_map, 64))
        d32 = d_all.astype(np.float32)
        too_close = bool(np.any(d32 < 20))
        if kind == "too_close" and not too_close:
            continue
        if kind != "too_close" and too_close:
            continue
        nb = [] if too_close else [int(i) for i, d in zip(idx_all, d32) if d <= 500][:11]
        if not too_close and len(nb) == 0:
            continue
        n_nb = len(nb)
        pos = np.zeros((T, nb_slots, 3), np.float32)
        val = np.zeros((T, nb_slots), np.uint8)
        if n_nb:
            pos[:, :n_nb] = map_world[frames][:, nb]
            val[:, :n_nb] = rng.uniform(size=(T, n_nb)) > dropout
            val[0, :n_nb] |= rng.uniform(size=n_nb) > 0.3   # keep most neighbours alive in the first frame
            if kind == "no_nb":
                val[int(rng.integers(T))] = 0
            elif kind == "behind":
                k = int(rng.integers(T))
                Rk = quat_to_R(cposes[k, :4].astype(np.float64))
                ck = -Rk.T @ cposes[k, 4:].astype(np.float64)         # camera centre in the world
                pos[k, :n_nb] = (2 * ck - pos[k, :n_nb].astype(np.float64)).astype(np.float32)
            elif kind == "noisy_nb":
                pos[:, :n_nb] += rng.normal(size=(T, n_nb, 3)).astype(np.float32) * np.float32(0.4)
            elif kind == "drift":
                pos[:, :n_nb, 0] += (np.arange(T, dtype=np.float32) * np.float32(0.12))[:, None]
            if kind not in ("no_nb",):
                for k in range(T):                                       # never leave a frame empty by accident
                    if not val[k, :n_nb].any():
                        val[k, 0] = 1
        track_ptr.append(track_ptr[-1] + T)
        uv_l.append(uv.astype(np.float32))
        pose_l.append(cposes)
        pos_l.append(pos)
        val_l.append(val)
        nnb_l.append(n_nb)
        kind_l.append(kind)
        truth_l.append(Xc[-1])
    if len(nnb_l) < n_cand:
        raise RuntimeError("triangulation_batch: candidate pool exhausted")
    return dict(cam=cam, n_cand=n_cand, track_ptr=np.array(track_ptr, np.int32),
                track_uv=np.concatenate(uv_l).astype(np.float32), track_pose=np.concatenate(pose_l).astype(np.float32),
                n_neighbours=np.array(nnb_l, np.int32), nb_pos=np.concatenate(pos_l).astype(np.float32),
                nb_valid=np.concatenate(val_l).astype(np.uint8), kinds=kind_l,
                truth=np.array(truth_l, np.float32), scale=1.0)
