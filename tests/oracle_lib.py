"""Loader for the CPU oracle (oracle/liborc.so). TEST INFRASTRUCTURE: importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

import nrslam_b200  # noqa: F401
from nrslam_b200.abi import Camera, Graph, Options, Stats, ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liborc.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_graph_get_edges.restype = C.c_int32
        _LIB.orc_graph_update_vertex.restype = C.c_int32
    return _LIB


def default_options():
    o = Options()
    lib().orc_default_options(C.byref(o))
    return o


class Oracle:
    """Same method names / argument meaning as nrslam_b200.api.Core so parity tests call both alike."""

    def __init__(self, opt=None):
        self.opt = opt or default_options()
        self.L = lib()

    def pose_only(self, cam, uv, X, pose):
        n = len(uv)
        pose = np.array(pose, np.float32)
        inl = np.zeros(n, np.uint8)
        st = Stats()
        rc = self.L.orc_pose_only(C.byref(self.opt), C.byref(cam), n, ptr(np.ascontiguousarray(uv, np.float32), C.c_float),
                                  ptr(np.ascontiguousarray(X, np.float32), C.c_float), ptr(pose, C.c_float),
                                  ptr(inl, C.c_uint8), C.byref(st))
        return dict(rc=rc, pose=pose, inliers=inl, stats=st.as_dict())

    def pose_deform(self, cam, uv, X_rest, point_vertex, vfs, graph, scale, pose, last_pos, pcg=0):
        n = len(uv)
        M = graph.n_vertices
        pose = np.array(pose, np.float32)
        last_pos = np.array(last_pos, np.float32)
        d = np.zeros((n, 3), np.float32)
        Xo = np.zeros((n, 3), np.float32)
        chi2 = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        med = C.c_float(0)
        lost = np.zeros(M, np.int32)
        nl = C.c_int32(0)
        st = Stats()
        g = graph.struct()
        timing = np.zeros(8, np.float64)
        uv = np.ascontiguousarray(uv, np.float32)
        X_rest = np.ascontiguousarray(X_rest, np.float32)
        pv = np.ascontiguousarray(point_vertex, np.int32)
        vfs = np.ascontiguousarray(vfs, np.int8)
        rc = self.L.orc_pose_deform_ex(C.byref(self.opt), C.byref(cam), n, ptr(uv, C.c_float), ptr(X_rest, C.c_float),
                                       ptr(pv, C.c_int32), ptr(vfs, C.c_int8), C.byref(g), C.c_float(scale),
                                       ptr(pose, C.c_float), ptr(last_pos, C.c_float), ptr(d, C.c_float),
                                       ptr(Xo, C.c_float), ptr(chi2, C.c_float), ptr(status, C.c_uint8), C.byref(med),
                                       ptr(lost, C.c_int32), C.byref(nl), C.byref(st), int(pcg),
                                       ptr(timing, C.c_double))
        return dict(rc=rc, pose=pose, deformation=d, X=Xo, chi2=chi2, status=status, median=med.value,
                    lost=lost[: nl.value].copy(), last_pos=last_pos, stats=st.as_dict(), timing=timing)

    def local_ba(self, cam, kf_pose, obs_kf, obs_vertex, uv, X, graph, scale, iterations=0, pcg=0):
        F = len(kf_pose)
        O = len(obs_kf)
        kf_pose = np.array(kf_pose, np.float32)
        X = np.array(X, np.float32)
        st = Stats()
        g = graph.struct()
        timing = np.zeros(8, np.float64)
        ok = np.ascontiguousarray(obs_kf, np.int32)
        ov = np.ascontiguousarray(obs_vertex, np.int32)
        uv = np.ascontiguousarray(uv, np.float32)
        rc = self.L.orc_local_ba_ex(C.byref(self.opt), C.byref(cam), F, ptr(kf_pose, C.c_float), O, ptr(ok, C.c_int32),
                                    ptr(ov, C.c_int32), ptr(uv, C.c_float), ptr(X, C.c_float), C.byref(g),
                                    C.c_float(scale), int(iterations), C.byref(st), int(pcg), ptr(timing, C.c_double))
        return dict(rc=rc, kf_pose=kf_pose, X=X, stats=st.as_dict(), timing=timing)

    def graph_get_edges(self, graph, vertex):
        g = graph.struct()
        out = np.zeros(graph.rowptr[vertex + 1] - graph.rowptr[vertex] + 1, np.int32)
        n = self.L.orc_graph_get_edges(C.byref(g), int(vertex), ptr(out, C.c_int32), len(out))
        return out[:n].copy()

    def graph_update_vertex(self, graph, vertex, positions):
        g = graph.struct()
        positions = np.ascontiguousarray(positions, np.float32)
        return self.L.orc_graph_update_vertex(C.byref(g), int(vertex), ptr(positions, C.c_float))


class OracleKLT:
    """LucasKanadeTracker restatement (oracle/orc_klt.cc). Same method names as nrslam_b200.api.KLT."""

    def __init__(self, win=21, max_level=4, max_iters=10, eps=1e-4, min_eig=1e-4):
        self.L = lib()
        self.L.orc_klt_create.restype = C.c_void_p
        self.win, self.max_level = win, max_level
        self.h = C.c_void_p(self.L.orc_klt_create(win, max_level, max_iters, C.c_float(eps), C.c_float(min_eig)))

    def close(self):
        if self.h:
            self.L.orc_klt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_points(self):
        return self.L.orc_klt_num_points(self.h)

    def set_reference(self, image, pts, mask=None):
        image = np.ascontiguousarray(image, np.uint8)
        pts = np.ascontiguousarray(pts, np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        return self.L.orc_klt_set_reference(self.h, ptr(image, C.c_uint8), image.shape[1], image.shape[0],
                                            image.strides[0], len(pts), ptr(pts, C.c_float), ptr(m, C.c_uint8),
                                            0 if m is None else m.strides[0])

    def track(self, image, pts, status, use_initial_flow=False, min_ssim=0.7, mask=None):
        image = np.ascontiguousarray(image, np.uint8)
        pts = np.array(pts, np.float32)
        status = np.array(status, np.uint8)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        nt = C.c_int32(0)
        rc = self.L.orc_klt_track(self.h, ptr(image, C.c_uint8), image.shape[1], image.shape[0], image.strides[0],
                                  len(pts), ptr(pts, C.c_float), ptr(status, C.c_uint8), int(use_initial_flow),
                                  C.c_float(min_ssim), ptr(m, C.c_uint8), 0 if m is None else m.strides[0],
                                  C.byref(nt))
        return dict(rc=rc, pts=pts, status=status, n_tracked=nt.value)

    def get_patch(self, idx):
        nl, a = self.max_level + 1, self.win * self.win
        gray = np.zeros((nl, a), np.int16)
        grad = np.zeros((nl, a, 2), np.int16)
        mean = np.zeros(nl, np.float32)
        mean2 = np.zeros(nl, np.float32)
        valid = np.zeros(nl, np.uint8)
        rc = self.L.orc_klt_get_patch(self.h, int(idx), ptr(gray, C.c_int16), ptr(grad, C.c_int16),
                                      ptr(mean, C.c_float), ptr(mean2, C.c_float), ptr(valid, C.c_uint8))
        return dict(rc=rc, gray=gray, grad=grad, mean=mean, mean2=mean2, valid=valid)

    def insert_patch(self, x, y, patch):
        return self.L.orc_klt_insert_patch(self.h, C.c_float(x), C.c_float(y), ptr(patch["gray"], C.c_int16),
                                           ptr(patch["grad"], C.c_int16), ptr(patch["mean"], C.c_float),
                                           ptr(patch["mean2"], C.c_float), ptr(patch["valid"], C.c_uint8))

    def clear(self):
        return self.L.orc_klt_clear(self.h)


def klt_pyramid(image, win, max_level, level):
    """Bordered level image / derivative of the oracle's pyramid restatement (pinning hook)."""
    image = np.ascontiguousarray(image, np.uint8)
    h, w = image.shape
    lw, lh = w, h
    for _ in range(level):
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
    img = np.zeros((lh + 2 * win, lw + 2 * win), np.uint8)
    der = np.zeros((lh + 2 * win, lw + 2 * win, 2), np.int16)
    ow, oh = C.c_int(0), C.c_int(0)
    n = lib().orc_klt_pyramid(ptr(image, C.c_uint8), w, h, image.strides[0], win, max_level, level, ptr(img, C.c_uint8),
                              ptr(der, C.c_int16), C.byref(ow), C.byref(oh))
    assert n > level and (ow.value, oh.value) == (lw, lh)
    return img, der


class OracleShiTomasi:
    """ShiTomasi restatement (oracle/orc_shi.cc). literal=True simulates the reference's border behaviour, False is the
    clean definition the CUDA kernel implements."""

    def __init__(self, nms_window=7):
        self.L = lib()
        self.L.orc_shi_create.restype = C.c_void_p
        self.h = C.c_void_p(self.L.orc_shi_create(int(nms_window)))

    def close(self):
        if self.h:
            self.L.orc_shi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def extract(self, image, existing=None, literal=False, capacity=20000, want_scores=False):
        image = np.ascontiguousarray(image, np.uint8)
        ex = np.zeros((0, 2), np.float32) if existing is None else np.ascontiguousarray(existing, np.float32)
        xy = np.zeros((capacity, 2), np.float32)
        ids = np.zeros(capacity, np.int32)
        sc = np.zeros(image.shape, np.float32) if want_scores else None
        n = self.L.orc_shi_extract(self.h, ptr(image, C.c_uint8), image.shape[1], image.shape[0], image.strides[0],
                                   ptr(ex, C.c_float), len(ex), int(literal), ptr(xy, C.c_float), ptr(ids, C.c_int32),
                                   capacity, ptr(sc, C.c_float))
        assert 0 <= n <= capacity
        return dict(n=n, xy=xy[:n].copy(), ids=ids[:n].copy(), scores=sc)


def deformable_triangulation(batch, order=0):
    """orc_deformable_triangulation over a synth.triangulation_batch dict -> (positions, statuses, lm_iterations)."""
    L = lib()
    n = int(batch["n_cand"])
    pos = np.zeros((n, 3), np.float32)
    status = np.zeros(n, np.int32)
    iters = np.zeros(n, np.int32)
    rc = L.orc_deformable_triangulation(
        C.byref(batch["cam"]), n, ptr(np.ascontiguousarray(batch["track_ptr"], np.int32), C.c_int32),
        ptr(np.ascontiguousarray(batch["track_uv"], np.float32), C.c_float),
        ptr(np.ascontiguousarray(batch["track_pose"], np.float32), C.c_float),
        ptr(np.ascontiguousarray(batch["n_neighbours"], np.int32), C.c_int32),
        ptr(np.ascontiguousarray(batch["nb_pos"], np.float32), C.c_float),
        ptr(np.ascontiguousarray(batch["nb_valid"], np.uint8), C.c_uint8), int(order), ptr(pos, C.c_float),
        ptr(status, C.c_int32), ptr(iters, C.c_int32))
    assert rc == 0
    return pos, status, iters


def graph_update_vertices(graph, vertices, positions):
    """Sequential RegularizationGraph::UpdateVertex loop (g2o_optimization.cc:458-474); graph arrays updated in place."""
    L = lib()
    v = np.ascontiguousarray(vertices, np.int32)
    good = np.zeros(len(v), np.int32)
    g = graph.struct()
    rc = L.orc_graph_update_vertices(C.byref(g), len(v), ptr(v, C.c_int32),
                                     ptr(np.ascontiguousarray(positions, np.float32), C.c_float), ptr(good, C.c_int32))
    assert rc == 0
    return good


def landmark_triangulation_frame(batch, rigid_ok, rad_per_pixel, min_track=5):
    """orc_landmark_triangulation_frame: deformable + rigid branch + vote of Mapping::LandmarkTriangulation."""
    L = lib()
    n = int(batch["n_cand"])
    dp, rp, sp = (np.zeros((n, 3), np.float32) for _ in range(3))
    ds, rs = np.zeros(n, np.int32), np.zeros(n, np.int32)
    sel = np.zeros(n, np.uint8)
    rc = L.orc_landmark_triangulation_frame(
        C.byref(batch["cam"]), n, ptr(np.ascontiguousarray(batch["track_ptr"], np.int32), C.c_int32),
        ptr(np.ascontiguousarray(batch["track_uv"], np.float32), C.c_float),
        ptr(np.ascontiguousarray(batch["track_pose"], np.float32), C.c_float),
        ptr(np.ascontiguousarray(batch["n_neighbours"], np.int32), C.c_int32),
        ptr(np.ascontiguousarray(batch["nb_pos"], np.float32), C.c_float),
        ptr(np.ascontiguousarray(batch["nb_valid"], np.uint8), C.c_uint8),
        ptr(np.ascontiguousarray(rigid_ok, np.uint8), C.c_uint8), C.c_float(rad_per_pixel), int(min_track),
        ptr(dp, C.c_float), ptr(ds, C.c_int32), ptr(rp, C.c_float), ptr(rs, C.c_int32), ptr(sp, C.c_float),
        ptr(sel, C.c_uint8))
    assert rc == 0
    return dict(deform_position=dp, deform_status=ds, rigid_position=rp, rigid_status=rs, selected_position=sp,
                selected=sel)


def point_reuse(cam, pose, image, X_world, patches, in_frame, forced=None, mask=None, klt_max_iters=10, klt_eps=1e-4,
                klt_min_eig=1e-4):
    """Tracking::PointReuse restated in C++ (oracle/orc_reuse.cc, tracking.cc:394-506); same outputs as api.point_reuse."""
    from nrslam_b200 import api
    rc, out = api._point_reuse_call(lib().orc_point_reuse, (), cam, pose, image, X_world, patches, in_frame, forced, mask,
                                    klt_max_iters, klt_eps, klt_min_eig)
    assert rc == 0
    return out
