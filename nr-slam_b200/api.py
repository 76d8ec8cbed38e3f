"""ctypes binding of libnrslam_b200.so (include/nrslam_b200.h) — the host-side mirror of the reference seam:

    Core.pose_only    <->  CameraPoseOptimization                 modules/optimization/g2o_optimization.h:27
    Core.pose_deform  <->  CameraPoseAndDeformationOptimization   modules/optimization/g2o_optimization.h:29-32
    Core.local_ba     <->  LocalDeformableBundleAdjustment        modules/optimization/g2o_optimization.h:39-40
    Core.graph_*      <->  RegularizationGraph::GetEdges/UpdateVertex   modules/map/regularization_graph.h:73,78

There is no CPU fallback: `load()` raises if the library has not been built, and `Core()` raises when no
sm_100 device is present.
"""
import ctypes as C
import os
import weakref

import numpy as np

from .abi import (Camera, FILTER_BORDER, FILTER_BRIGHT, FILTER_PREDEFINED, Graph, GraphArrays, MaskFilter, Options, Stats,  # noqa: F401
                  ptr)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnrslam_b200.so")
_LIB = None

ERRORS = {-1: "no sm_100 device", -2: "CUDA error", -3: "bad argument", -4: "allocation failed", -5: "NCCL error",
          -6: "not implemented", 1: "too few points / keyframes", 2: "non-finite result"}


IPC_HANDLE_BYTES = 64


class NrslamError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nrslam_b200 error %d (%s): %s" % (code, ERRORS.get(code, "?"), msg))
        self.code = code


def load():
    """Load the C-ABI library. Raises (no fallback) when it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libnrslam_b200.so not built — run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        lib.nrslam_b200_last_error.restype = C.c_char_p
        lib.nrslam_b200_graph_get_edges.restype = C.c_int32
        lib.nrslam_b200_graph_update_vertex.restype = C.c_int32
        lib.nrslam_b200_graph_store_destroy.restype = None
        lib.nrslam_b200_graph_store_destroy.argtypes = [C.c_void_p]
        lib.nrslam_b200_graph_store_add_edges.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32),
                                                          C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        lib.nrslam_b200_graph_store_set_sigma.argtypes = [C.c_void_p, C.c_float]
        lib.nrslam_b200_graph_store_view.argtypes = [C.c_void_p, C.POINTER(Graph)]
        lib.nrslam_b200_klt_num_points.restype = C.c_int32
        lib.nrslam_b200_pre_last_ms.restype = C.c_float
        lib.nrslam_b200_pre_last_launches.restype = C.c_int32
        lib.nrslam_b200_tri_last_ms.restype = C.c_float
        _LIB = lib
    return _LIB


class GraphStore:
    """RegularizationGraph as the library's owning store (include/nrslam_b200.h: nrslam_b200_graph_store_*): AddEdge /
    SetSigma (map/regularization_graph.cc:33-55) and a CSR view whose attribute arrays the optimisation entry points
    refresh in place. Pure host code: works without a GPU."""

    def __init__(self, weight_sigma, stretching_th=1.1):
        self.L = load()
        self._h = C.c_void_p()
        rc = self.L.nrslam_b200_graph_store_create(C.c_float(weight_sigma), C.c_float(stretching_th), C.byref(self._h))
        if rc:
            raise ValueError("graph_store_create: %d" % rc)

    def add_edges(self, v1, v2, relative_position):
        v1 = np.ascontiguousarray(v1, np.int32)
        v2 = np.ascontiguousarray(v2, np.int32)
        rel = np.ascontiguousarray(relative_position, np.float32).reshape(-1, 3)
        assert len(v1) == len(v2) == len(rel)
        return self.L.nrslam_b200_graph_store_add_edges(self._h, len(v1), ptr(v1, C.c_int32), ptr(v2, C.c_int32),
                                                        ptr(rel, C.c_float))

    def set_sigma(self, sigma):
        return self.L.nrslam_b200_graph_store_set_sigma(self._h, C.c_float(sigma))

    @property
    def n_vertices(self):
        return self.struct().n_vertices

    @property
    def n_edges(self):
        return self.struct().n_edges

    def struct(self):
        """The store's CSR view as an abi.Graph (valid until the next add_edges / close)."""
        g = Graph()
        rc = self.L.nrslam_b200_graph_store_view(self._h, C.byref(g))
        if rc:
            raise ValueError("graph_store_view: %d" % rc)
        return g

    def arrays(self):
        """Copies of the view as an abi.GraphArrays (for comparisons)."""
        g = self.struct()
        M, E = g.n_vertices, g.n_edges

        def arr(p, n, dt):
            return np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy() if n else np.zeros(0, dt)
        return GraphArrays(arr(g.rowptr, M + 1, np.int32) if M or E else np.zeros(1, np.int32), arr(g.col, 2 * E, np.int32),
                           arr(g.eid, 2 * E, np.int32), arr(g.weight, E, np.float32), arr(g.first_distance, E, np.float32),
                           arr(g.min_distance, E, np.float32), arr(g.max_distance, E, np.float32),
                           arr(g.status, E, np.uint8), g.weight_sigma, g.stretching_th)

    def close(self):
        if self._h:
            self.L.nrslam_b200_graph_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_options():
    o = Options()
    load().nrslam_b200_default_options(C.byref(o))
    return o


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


class Core:
    """One context = one GPU. Method names / argument meaning follow the reference functions."""

    def __init__(self, opt=None):
        self.L = load()
        self.opt = opt or default_options()
        self._ctx = C.c_void_p()
        self._children = []  # weak references to the KLT trackers living on this context
        rc = self.L.nrslam_b200_create(C.byref(self.opt), C.byref(self._ctx))
        if rc != 0:
            raise NrslamError(rc, "nrslam_b200_create failed")

    def _adopt(self, child):
        """Registers a child object (KLT, ShiTomasi, Pre, Triangulator) that holds a pointer to this context; dead
        references are dropped so the list does not grow over a sequence (point_reuse builds a tracker per frame)."""
        self._children = [r for r in self._children if r() is not None]
        self._children.append(weakref.ref(child))

    def close(self):
        if self._ctx:
            for ref in self._children:  # trackers hold a pointer to the context: destroy them first
                k = ref()
                if k is not None:
                    k.close()
            self._children = []
            self.L.nrslam_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise NrslamError(rc, (self.L.nrslam_b200_last_error(self._ctx) or b"").decode())
        return rc

    def device_info(self):
        d, s = C.c_int32(), C.c_int32()
        self._check(self.L.nrslam_b200_device_info(self._ctx, C.byref(d), C.byref(s)))
        return d.value, s.value

    def pose_only(self, cam, uv, X, pose):
        n = len(uv)
        uv, X = _f32(uv), _f32(X)
        pose = np.array(pose, np.float32)
        inl = np.zeros(n, np.uint8)
        st = Stats()
        rc = self._check(self.L.nrslam_b200_pose_only(self._ctx, C.byref(cam), n, ptr(uv, C.c_float),
                                                      ptr(X, C.c_float), ptr(pose, C.c_float), ptr(inl, C.c_uint8),
                                                      C.byref(st)))
        return dict(rc=rc, pose=pose, inliers=inl, stats=st.as_dict())

    def pose_deform(self, cam, uv, X_rest, point_vertex, vfs, graph, scale, pose, last_pos):
        """`graph` (abi.GraphArrays) and nothing else is updated in place; everything else is returned."""
        n = len(uv)
        M = graph.n_vertices
        uv, X_rest = _f32(uv), _f32(X_rest)
        pv = np.ascontiguousarray(point_vertex, np.int32)
        vfs = np.ascontiguousarray(vfs, np.int8)
        pose = np.array(pose, np.float32)
        last_pos = np.array(last_pos, np.float32)
        d = np.zeros((n, 3), np.float32)
        Xo = np.zeros((n, 3), np.float32)
        chi2 = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        med = C.c_float(0)
        lost = np.zeros(max(M, 1), np.int32)
        nl = C.c_int32(0)
        st = Stats()
        g = graph.struct()
        rc = self._check(self.L.nrslam_b200_pose_deform(
            self._ctx, C.byref(cam), n, ptr(uv, C.c_float), ptr(X_rest, C.c_float), ptr(pv, C.c_int32),
            ptr(vfs, C.c_int8), C.byref(g), C.c_float(scale), ptr(pose, C.c_float), ptr(last_pos, C.c_float),
            ptr(d, C.c_float), ptr(Xo, C.c_float), ptr(chi2, C.c_float), ptr(status, C.c_uint8), C.byref(med),
            ptr(lost, C.c_int32), C.byref(nl), C.byref(st)))
        return dict(rc=rc, pose=pose, deformation=d, X=Xo, chi2=chi2, status=status, median=med.value,
                    lost=lost[: nl.value].copy(), last_pos=last_pos, stats=st.as_dict())

    def track_pose_and_deform(self, cam, uv, X_rest, point_vertex, vfs, graph, scale, pose, last_pos):
        """Tracking::TrackCameraAndDeformation after the data association (tracking.cc:291-330): CameraPoseOptimization
        then CameraPoseAndDeformationOptimization seeded with its pose, as ONE call (nrslam_b200_track_pose_and_deform).
        Returns (pose_only result, pose_deform result) with the same keys as the two separate calls."""
        n = len(uv)
        M = graph.n_vertices
        uv, X_rest = _f32(uv), _f32(X_rest)
        pv = np.ascontiguousarray(point_vertex, np.int32)
        vfs = np.ascontiguousarray(vfs, np.int8)
        pose = np.array(pose, np.float32)
        last_pos = np.array(last_pos, np.float32)
        d = np.zeros((n, 3), np.float32)
        Xo = np.zeros((n, 3), np.float32)
        chi2 = np.zeros(n, np.float32)
        status = np.zeros(n, np.uint8)
        med = C.c_float(0)
        lost = np.zeros(max(M, 1), np.int32)
        nl = C.c_int32(0)
        pose0 = np.zeros(7, np.float32)
        inl0 = np.zeros(n, np.uint8)
        st0, st = Stats(), Stats()
        g = graph.struct()
        rc = self._check(self.L.nrslam_b200_track_pose_and_deform(
            self._ctx, C.byref(cam), n, ptr(uv, C.c_float), ptr(X_rest, C.c_float), ptr(pv, C.c_int32),
            ptr(vfs, C.c_int8), C.byref(g), C.c_float(scale), ptr(pose, C.c_float), ptr(last_pos, C.c_float),
            ptr(d, C.c_float), ptr(Xo, C.c_float), ptr(chi2, C.c_float), ptr(status, C.c_uint8), C.byref(med),
            ptr(lost, C.c_int32), C.byref(nl), ptr(pose0, C.c_float), ptr(inl0, C.c_uint8), C.byref(st0), C.byref(st)))
        r0 = dict(rc=rc, pose=pose0, inliers=inl0.astype(bool), stats=st0.as_dict())
        r1 = dict(rc=rc, pose=pose, deformation=d, X=Xo, chi2=chi2, status=status, median=med.value,
                  lost=lost[: nl.value].copy(), last_pos=last_pos, stats=st.as_dict())
        return r0, r1

    def local_ba(self, cam, kf_pose, obs_kf, obs_vertex, uv, X, graph, scale, iterations=0):
        F = len(kf_pose)
        O = len(obs_kf)
        kf_pose = np.array(kf_pose, np.float32)
        X = np.array(X, np.float32)
        ok = np.ascontiguousarray(obs_kf, np.int32)
        ov = np.ascontiguousarray(obs_vertex, np.int32)
        uv = _f32(uv)
        st = Stats()
        g = graph.struct()
        rc = self._check(self.L.nrslam_b200_local_ba(
            self._ctx, C.byref(cam), F, ptr(kf_pose, C.c_float), O, ptr(ok, C.c_int32), ptr(ov, C.c_int32),
            ptr(uv, C.c_float), ptr(X, C.c_float), C.byref(g), C.c_float(scale), int(iterations), C.byref(st)))
        return dict(rc=rc, kf_pose=kf_pose, X=X, stats=st.as_dict())

    # ---- landmark-sharded BA over several GPUs (one process + one Core per GPU; include/nrslam_b200.h)
    def shard_init(self, rank, world, max_rows, max_poses):
        """Allocates this rank's exchange buffer; returns its 64-byte CUDA IPC handle (bytes)."""
        h = (C.c_ubyte * IPC_HANDLE_BYTES)()
        self._check(self.L.nrslam_b200_shard_init(self._ctx, int(rank), int(world), int(max_rows), int(max_poses), h))
        self._shard = (int(rank), int(world))
        return bytes(h)

    def shard_attach(self, handles):
        """handles: the ranks' IPC handles in rank order (e.g. from torch.distributed.all_gather_object)."""
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self.L.nrslam_b200_shard_attach(self._ctx, buf))

    def local_ba_sharded(self, cam, kf_pose, obs_kf, obs_vertex, uv, X, graph, scale, iterations=0):
        """Collective: every rank passes the full window. Returns the poses (identical on every rank), X with this
        rank's observations optimised (others as passed in) and owner[n_obs]."""
        F = len(kf_pose)
        O = len(obs_kf)
        kf_pose = np.array(kf_pose, np.float32)
        X = np.array(X, np.float32)
        ok = np.ascontiguousarray(obs_kf, np.int32)
        ov = np.ascontiguousarray(obs_vertex, np.int32)
        uv = _f32(uv)
        owner = np.zeros(O, np.int32)
        st = Stats()
        g = graph.struct()
        rc = self._check(self.L.nrslam_b200_local_ba_sharded(
            self._ctx, C.byref(cam), F, ptr(kf_pose, C.c_float), O, ptr(ok, C.c_int32), ptr(ov, C.c_int32),
            ptr(uv, C.c_float), ptr(X, C.c_float), C.byref(g), C.c_float(scale), int(iterations),
            ptr(owner, C.c_int32), C.byref(st)))
        return dict(rc=rc, kf_pose=kf_pose, X=X, owner=owner, stats=st.as_dict())

    def resolve(self, which):
        """Re-run the device program of the last staged problem (0 pose_only, 1 pose_deform, 2 local_ba, 3 the
        lost-point stage of the last pose_deform call)."""
        st = Stats()
        self._check(self.L.nrslam_b200_resolve(self._ctx, int(which), C.byref(st)))
        return st.as_dict()

    def graph_get_edges(self, graph, vertex):
        g = graph.struct()
        out = np.zeros(graph.rowptr[vertex + 1] - graph.rowptr[vertex] + 1, np.int32)
        n = self.L.nrslam_b200_graph_get_edges(C.byref(g), int(vertex), ptr(out, C.c_int32), len(out))
        return out[:n].copy()

    def graph_update_vertex(self, graph, vertex, positions):
        g = graph.struct()
        positions = _f32(positions)
        return self.L.nrslam_b200_graph_update_vertex(C.byref(g), int(vertex), ptr(positions, C.c_float))

    def graph_get_edges_batch(self, graph, vertices, top_k=32):
        """RegularizationGraph::GetEdges for many vertices in one launch: (entries [n, top_k] CSR entry indices, -1
        padded; counts [n] = len(GetEdges(v)))."""
        g = graph.struct()
        v = np.ascontiguousarray(vertices, np.int32)
        ent = np.full((len(v), int(top_k)), -1, np.int32)
        cnt = np.zeros(len(v), np.int32)
        self._check(self.L.nrslam_b200_graph_get_edges_batch(self._ctx, C.byref(g), len(v), ptr(v, C.c_int32),
                                                             int(top_k), ptr(ent, C.c_int32), ptr(cnt, C.c_int32)))
        return ent, cnt

    def graph_update_vertices(self, graph, vertices, positions):
        """The UpdateVertex loop of CameraPoseAndDeformationOptimization (g2o_optimization.cc:458-474) on the device:
        graph attribute arrays updated in place, returns UpdateVertex's good-connection count per vertex."""
        g = graph.struct()
        v = np.ascontiguousarray(vertices, np.int32)
        positions = _f32(positions)
        good = np.zeros(len(v), np.int32)
        self._check(self.L.nrslam_b200_graph_update_vertices(self._ctx, C.byref(g), len(v), ptr(v, C.c_int32),
                                                             ptr(positions, C.c_float), ptr(good, C.c_int32)))
        return good


class KLT:
    """LucasKanadeTracker (modules/matching/lucas_kanade_tracker.h:55-70) on the GPU: SetReferenceImage, Track,
    Get/InsertPhotometricInformation, clear. One object per reference image; point i of every call is point i of
    set_reference (+ inserted ones), exactly like the reference (KLT index == frame index)."""

    def __init__(self, core, win=21, max_level=4, max_iters=10, eps=1e-4, min_eig=1e-4):
        self.core = core
        self.L = core.L
        self.win, self.max_level = win, max_level
        self._h = C.c_void_p()
        rc = self.L.nrslam_b200_klt_create(core._ctx, win, max_level, max_iters, C.c_float(eps), C.c_float(min_eig),
                                           C.byref(self._h))
        if rc != 0:
            raise NrslamError(rc, (self.L.nrslam_b200_last_error(core._ctx) or b"").decode())
        import weakref
        core._adopt(self)

    def close(self):
        if self._h:
            if self.core._ctx:  # the context is still alive
                self.L.nrslam_b200_klt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise NrslamError(rc, (self.L.nrslam_b200_last_error(self.core._ctx) or b"").decode())
        return rc

    def num_points(self):
        return self.L.nrslam_b200_klt_num_points(self._h)

    def set_reference(self, image, pts, mask=None):
        image = np.ascontiguousarray(image, np.uint8)
        pts = _f32(pts)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        return self._check(self.L.nrslam_b200_klt_set_reference(
            self._h, ptr(image, C.c_uint8), image.shape[1], image.shape[0], image.strides[0], len(pts),
            ptr(pts, C.c_float), ptr(m, C.c_uint8), 0 if m is None else m.strides[0]))

    def track(self, image, pts, status, use_initial_flow=False, min_ssim=0.7, mask=None):
        image = np.ascontiguousarray(image, np.uint8)
        pts = np.array(pts, np.float32)
        status = np.array(status, np.uint8)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        nt = C.c_int32(0)
        rc = self._check(self.L.nrslam_b200_klt_track(
            self._h, ptr(image, C.c_uint8), image.shape[1], image.shape[0], image.strides[0], len(pts),
            ptr(pts, C.c_float), ptr(status, C.c_uint8), int(use_initial_flow), C.c_float(min_ssim),
            ptr(m, C.c_uint8), 0 if m is None else m.strides[0], C.byref(nt)))
        return dict(rc=rc, pts=pts, status=status, n_tracked=nt.value)

    def retrack(self):
        ms = C.c_float(0)
        self._check(self.L.nrslam_b200_klt_retrack(self._h, C.byref(ms)))
        return ms.value

    def get_patch(self, idx):
        nl, a = self.max_level + 1, self.win * self.win
        gray = np.zeros((nl, a), np.int16)
        grad = np.zeros((nl, a, 2), np.int16)
        mean = np.zeros(nl, np.float32)
        mean2 = np.zeros(nl, np.float32)
        valid = np.zeros(nl, np.uint8)
        rc = self._check(self.L.nrslam_b200_klt_get_patch(self._h, int(idx), ptr(gray, C.c_int16),
                                                          ptr(grad, C.c_int16), ptr(mean, C.c_float),
                                                          ptr(mean2, C.c_float), ptr(valid, C.c_uint8)))
        return dict(rc=rc, gray=gray, grad=grad, mean=mean, mean2=mean2, valid=valid)

    def insert_patch(self, x, y, patch):
        return self._check(self.L.nrslam_b200_klt_insert_patch(
            self._h, C.c_float(x), C.c_float(y), ptr(np.ascontiguousarray(patch["gray"], np.int16), C.c_int16),
            ptr(np.ascontiguousarray(patch["grad"], np.int16), C.c_int16),
            ptr(np.ascontiguousarray(patch["mean"], np.float32), C.c_float),
            ptr(np.ascontiguousarray(patch["mean2"], np.float32), C.c_float),
            ptr(np.ascontiguousarray(patch["valid"], np.uint8), C.c_uint8)))

    def clear(self):
        return self._check(self.L.nrslam_b200_klt_clear(self._h))

    def debug_level(self, which, level, shape_hw):
        """Bordered pyramid level (diagnostics): returns (image u8, derivative int16 x2)."""
        h, w = shape_hw
        for _ in range(level):
            w, h = (w + 1) // 2, (h + 1) // 2
        img = np.zeros((h + 2 * self.win, w + 2 * self.win), np.uint8)
        der = np.zeros((h + 2 * self.win, w + 2 * self.win, 2), np.int16)
        ow, oh = C.c_int32(0), C.c_int32(0)
        self._check(self.L.nrslam_b200_klt_debug_level(self._h, int(which), int(level), ptr(img, C.c_uint8),
                                                       ptr(der, C.c_int16), C.byref(ow), C.byref(oh)))
        assert (ow.value, oh.value) == (w, h)
        return img, der


def shard_partition(world, kf_pose, obs_kf, obs_vertex, uv, X, graph, scale, opt=None):
    """Host-only view of the landmark partition local_ba_sharded uses (no GPU): owner[n_obs] and per-rank counts."""
    L = load()
    F, O = len(kf_pose), len(obs_kf)
    kf_pose = _f32(kf_pose)
    X = _f32(X)
    ok = np.ascontiguousarray(obs_kf, np.int32)
    ov = np.ascontiguousarray(obs_vertex, np.int32)
    uv = _f32(uv)
    owner = np.zeros(O, np.int32)
    n_own, n_halo, n_push = (np.zeros(world, np.int32) for _ in range(3))
    n_edges = np.zeros(3 * world, np.int32)
    g = graph.struct()
    rc = L.nrslam_b200_shard_partition(C.byref(opt) if opt is not None else None, int(world), F,
                                       ptr(kf_pose, C.c_float), O, ptr(ok, C.c_int32), ptr(ov, C.c_int32),
                                       ptr(uv, C.c_float), ptr(X, C.c_float), C.byref(g), C.c_float(scale),
                                       ptr(owner, C.c_int32), ptr(n_own, C.c_int32), ptr(n_halo, C.c_int32),
                                       ptr(n_push, C.c_int32), ptr(n_edges, C.c_int32))
    if rc != 0:
        raise NrslamError(rc, "shard_partition")
    return dict(owner=owner, n_own=n_own, n_halo=n_halo, n_push=n_push, n_edges=n_edges.reshape(world, 3))


def project_points(cam, pose, X):
    """CameraModel::Project of world points through a Sophus::SE3f pose, in fp32 like the reference
    (tracking.cc:401-407: `CameraTransformationWorld() * landmark` then `calibration_->Project`)."""
    pose = np.asarray(pose, np.float32)
    q, t = pose[:4], pose[4:]
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float32)
    Pc = (np.asarray(X, np.float32) @ R.T + t).astype(np.float32)
    p = np.array(cam.params[:], np.float32)
    if cam.model == 0:
        uv = np.stack([p[0] * Pc[:, 0] / Pc[:, 2] + p[2], p[1] * Pc[:, 1] / Pc[:, 2] + p[3]], 1)
    else:
        r2 = Pc[:, 0] * Pc[:, 0] + Pc[:, 1] * Pc[:, 1]
        th = np.arctan2(np.sqrt(r2), Pc[:, 2]).astype(np.float32)
        psi = np.arctan2(Pc[:, 1], Pc[:, 0]).astype(np.float32)
        t2 = th * th
        r = th + p[4] * th * t2 + p[5] * th * t2 * t2 + p[6] * th * t2 * t2 * t2 + p[7] * th * t2 * t2 * t2 * t2
        uv = np.stack([p[0] * r * np.cos(psi) + p[2], p[1] * r * np.sin(psi) + p[3]], 1)
    return uv.astype(np.float32), Pc


def pack_reuse_patches(patches, win=21):
    """Two finest levels of every map point's PhotometricInformation as the flat arrays nrslam_b200_point_reuse takes."""
    n, A = len(patches), win * win
    gray = np.zeros((n, 2, A), np.int16)
    grad = np.zeros((n, 2, A, 2), np.int16)
    mean = np.zeros((n, 2), np.float32)
    mean2 = np.zeros((n, 2), np.float32)
    valid = np.zeros((n, 2), np.uint8)
    for i, pt in enumerate(patches):
        gray[i] = np.asarray(pt["gray"][:2]).reshape(2, A)
        grad[i] = np.asarray(pt["grad"][:2]).reshape(2, A, 2)
        mean[i] = pt["mean"][:2]
        mean2[i] = pt["mean2"][:2]
        valid[i] = pt["valid"][:2]
    return gray, grad, mean, mean2, valid


def _point_reuse_call(fn, ctx_args, cam, pose, image, X_world, patches, in_frame, forced, mask, klt_max_iters, klt_eps,
                      klt_min_eig):
    image = np.ascontiguousarray(image, np.uint8)
    h, w = image.shape
    X = _f32(X_world).reshape(-1, 3)
    n = len(X)
    gray, grad, mean, mean2, valid = patches if isinstance(patches, tuple) else pack_reuse_patches(patches)
    inf = np.ascontiguousarray(in_frame, np.uint8)
    frc = None if forced is None else np.ascontiguousarray(forced, np.uint8)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    pose = _f32(pose)
    cand = np.zeros(max(n, 1), np.int32)
    seed = np.zeros((max(n, 1), 2), np.float32)
    uv = np.zeros((max(n, 1), 2), np.float32)
    st = np.zeros(max(n, 1), np.uint8)
    acc = np.zeros(max(n, 1), np.uint8)
    nc, nr = C.c_int32(0), C.c_int32(0)
    rc = fn(*ctx_args, C.byref(cam), ptr(pose, C.c_float), ptr(image, C.c_uint8), w, h, image.strides[0],
            ptr(m, C.c_uint8), 0 if m is None else m.strides[0], n, ptr(X, C.c_float), ptr(inf, C.c_uint8),
            ptr(frc, C.c_uint8), int(klt_max_iters), C.c_float(klt_eps), C.c_float(klt_min_eig), ptr(gray, C.c_int16),
            ptr(grad, C.c_int16), ptr(mean, C.c_float), ptr(mean2, C.c_float), ptr(valid, C.c_uint8),
            ptr(cand, C.c_int32), ptr(seed, C.c_float), ptr(uv, C.c_float), ptr(st, C.c_uint8), ptr(acc, C.c_uint8),
            C.byref(nc), C.byref(nr))
    k = nc.value
    return rc, dict(candidates=cand[:k].astype(np.int64), seeds=seed[:k].copy(), pts=uv[:k].copy(),
                    status=st[:k].copy(), accepted=acc[:k].astype(bool), n_reused=nr.value)


def point_reuse(core, cam, pose, image, X_world, patches, in_frame, forced=None, mask=None, klt_max_iters=10,
                klt_eps=1e-4, klt_min_eig=1e-4):
    """Tracking::PointReuse (modules/tracking/tracking.cc:394-506) through the C ABI (nrslam_b200_point_reuse):
    project the map points that are not in the frame (or that were reported lost), keep those inside the image with
    non-negative depth, track them with a fresh 2-level KLT fed from their stored patches (initial flow = projection,
    SSIM 0.75) and gate by the squared reprojection error (> 5.99 rejected). `patches[i]` is the map point's
    PhotometricInformation as returned by KLT.get_patch (or the tuple of pack_reuse_patches). Returns a dict:
    candidates (point indices, ascending), seeds, pts (tracked keypoints), status, accepted, n_reused."""
    rc, out = _point_reuse_call(core.L.nrslam_b200_point_reuse, (core._ctx,), cam, pose, image, X_world, patches,
                                in_frame, forced, mask, klt_max_iters, klt_eps, klt_min_eig)
    core._check(rc)
    return out


class ShiTomasi:
    """ShiTomasi feature extractor (modules/features/shi_tomasi.h:30-60; Feature::Extract, features/feature.h:34)."""

    def __init__(self, core, nms_window=7):
        self.core = core
        self.L = core.L
        self._h = C.c_void_p()
        rc = self.L.nrslam_b200_shi_create(core._ctx, int(nms_window), C.byref(self._h))
        if rc != 0:
            raise NrslamError(rc, (self.L.nrslam_b200_last_error(core._ctx) or b"").decode())
        import weakref
        core._adopt(self)
        self._shape = None

    def close(self):
        if self._h:
            if self.core._ctx:
                self.L.nrslam_b200_shi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def extract(self, image, existing=None, capacity=20000, want_scores=False):
        """Returns the NEW keypoints (raster order) and their class ids, like ShiTomasi::Extract appends them."""
        image = np.ascontiguousarray(image, np.uint8)
        ex = np.zeros((0, 2), np.float32) if existing is None else _f32(existing)
        xy = np.zeros((capacity, 2), np.float32)
        ids = np.zeros(capacity, np.int32)
        n = C.c_int32(0)
        rc = self.L.nrslam_b200_shi_extract(self._h, ptr(image, C.c_uint8), image.shape[1], image.shape[0],
                                            image.strides[0], ptr(ex, C.c_float), len(ex), ptr(xy, C.c_float),
                                            ptr(ids, C.c_int32), capacity, C.byref(n))
        if rc < 0:
            raise NrslamError(rc, (self.L.nrslam_b200_last_error(self.core._ctx) or b"").decode())
        m = min(n.value, capacity)
        out = dict(n=n.value, xy=xy[:m].copy(), ids=ids[:m].copy(), scores=None)
        if want_scores:
            sc = np.zeros(image.shape, np.float32)
            self.L.nrslam_b200_shi_debug_scores(self._h, ptr(sc, C.c_float))
            out["scores"] = sc
        return out


class Pre:
    """Per-frame image pre-processing on the GPU: System::ImageProcessing (SLAM/system.cc:189-201: RGB -> gray -> CLAHE)
    and Masker::mask (masking/masker.cc:80-92). Bit-exact with OpenCV 4."""

    def __init__(self, core, max_width=1440, max_height=1080):
        self.core = core
        self.L = core.L
        self._h = C.c_void_p()
        rc = self.L.nrslam_b200_pre_create(core._ctx, int(max_width), int(max_height), C.byref(self._h))
        if rc != 0:
            raise NrslamError(rc, (self.L.nrslam_b200_last_error(core._ctx) or b"").decode())
        core._adopt(self)

    def close(self):
        if getattr(self, "_h", None):
            if self.core._ctx:
                self.L.nrslam_b200_pre_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise NrslamError(rc, (self.L.nrslam_b200_last_error(self.core._ctx) or b"").decode())

    def image(self, rgb, clip_limit=3.0, tiles=(8, 8)):
        """rgb: (h, w, 3) uint8. Returns (gray, clahe) like ImageProcessing's im_gray and return value."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w = rgb.shape[:2]
        gray = np.empty((h, w), np.uint8)
        eq = np.empty((h, w), np.uint8)
        self._check(self.L.nrslam_b200_pre_image(self._h, ptr(rgb, C.c_uint8), w, h, 3 * w, C.c_float(clip_limit),
                                                 int(tiles[0]), int(tiles[1]), ptr(gray, C.c_uint8),
                                                 ptr(eq, C.c_uint8)))
        return gray, eq

    def mask(self, gray, filters, shape=None):
        """filters: list of ("bright", th) / ("border", rb, re, cb, ce) / ("predefined", mask). gray=None re-uses the
        gray image of the last image() call (pass its shape)."""
        if gray is not None:
            gray = np.ascontiguousarray(gray, np.uint8)
            h, w = gray.shape
        else:
            h, w = shape
        arr = (MaskFilter * max(len(filters), 1))()
        keep = []
        for i, f in enumerate(filters):
            if f[0] == "bright":
                arr[i].kind, arr[i].th = FILTER_BRIGHT, int(f[1])
            elif f[0] == "border":
                arr[i].kind = FILTER_BORDER
                arr[i].rb, arr[i].re, arr[i].cb, arr[i].ce = (int(v) for v in f[1:5])
            else:
                m = np.ascontiguousarray(f[1], np.uint8)
                keep.append(m)
                arr[i].kind, arr[i].mask = FILTER_PREDEFINED, ptr(m, C.c_uint8)
        out = np.empty((h, w), np.uint8)
        self._check(self.L.nrslam_b200_pre_mask(self._h, ptr(gray, C.c_uint8), w, h, arr, len(filters),
                                                ptr(out, C.c_uint8)))
        return out

    def last_ms(self):
        return float(self.L.nrslam_b200_pre_last_ms(self._h))

    def last_launches(self):
        return int(self.L.nrslam_b200_pre_last_launches(self._h))


class Triangulator:
    """Batched DeformableTriangulation (modules/optimization/g2o_optimization.h:34-37, .cc:559-814): one call per frame
    over all candidates of Mapping::LandmarkTriangulation (mapping/mapping.cc:88-113), one CTA per candidate."""

    def __init__(self, core):
        self.core = core
        self.L = core.L
        self._h = C.c_void_p()
        rc = self.L.nrslam_b200_tri_create(core._ctx, C.byref(self._h))
        if rc != 0:
            raise NrslamError(rc, (self.L.nrslam_b200_last_error(core._ctx) or b"").decode())
        import weakref
        core._adopt(self)

    def close(self):
        if self._h:
            if self.core._ctx:
                self.L.nrslam_b200_tri_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise NrslamError(rc, (self.L.nrslam_b200_last_error(self.core._ctx) or b"").decode())
        return rc

    def run(self, cam, track_ptr, track_uv, track_pose, n_neighbours, nb_pos, nb_valid, scale=1.0):
        """-> dict(position [n,3] f32, status [n] i32 (NRSLAM_B200_TRI_*), lm_iterations [n] i32, gpu_ms)."""
        track_ptr = np.ascontiguousarray(track_ptr, np.int32)
        n = len(track_ptr) - 1
        uv, pose = _f32(track_uv), _f32(track_pose)
        nnb = np.ascontiguousarray(n_neighbours, np.int32)
        pos, val = _f32(nb_pos), np.ascontiguousarray(nb_valid, np.uint8)
        out = np.zeros((n, 3), np.float32)
        status = np.full(n, -1, np.int32)
        iters = np.zeros(n, np.int32)
        self._check(self.L.nrslam_b200_tri_run(self._h, C.byref(cam), n, ptr(track_ptr, C.c_int32), ptr(uv, C.c_float),
                                               ptr(pose, C.c_float), ptr(nnb, C.c_int32), ptr(pos, C.c_float),
                                               ptr(val, C.c_uint8), C.c_float(scale), ptr(out, C.c_float),
                                               ptr(status, C.c_int32), ptr(iters, C.c_int32)))
        return dict(position=out, status=status, lm_iterations=iters, gpu_ms=self.L.nrslam_b200_tri_last_ms(self._h))

    def run_frame(self, cam, track_ptr, track_uv, track_pose, n_neighbours, nb_pos, nb_valid, rigid_ok,
                  rad_per_pixel, min_track=5, scale=1.0):
        """Mapping::LandmarkTriangulation's per-candidate compute and vote (mapping/mapping.cc:65-212) in one launch."""
        track_ptr = np.ascontiguousarray(track_ptr, np.int32)
        n = len(track_ptr) - 1
        uv, pose = _f32(track_uv), _f32(track_pose)
        nnb = np.ascontiguousarray(n_neighbours, np.int32)
        pos, val = _f32(nb_pos), np.ascontiguousarray(nb_valid, np.uint8)
        rok = np.ascontiguousarray(rigid_ok, np.uint8)
        dp, rp, sp = (np.zeros((n, 3), np.float32) for _ in range(3))
        ds, rs = np.full(n, -1, np.int32), np.full(n, -1, np.int32)
        sel = np.zeros(n, np.uint8)
        self._check(self.L.nrslam_b200_tri_run_frame(
            self._h, C.byref(cam), n, ptr(track_ptr, C.c_int32), ptr(uv, C.c_float), ptr(pose, C.c_float),
            ptr(nnb, C.c_int32), ptr(pos, C.c_float), ptr(val, C.c_uint8), ptr(rok, C.c_uint8),
            C.c_float(rad_per_pixel), int(min_track), C.c_float(scale), ptr(dp, C.c_float), ptr(ds, C.c_int32),
            ptr(rp, C.c_float), ptr(rs, C.c_int32), ptr(sp, C.c_float), ptr(sel, C.c_uint8)))
        return dict(deform_position=dp, deform_status=ds, rigid_position=rp, rigid_status=rs, selected_position=sp,
                    selected=sel, gpu_ms=self.L.nrslam_b200_tri_last_ms(self._h))

    def run_batch(self, batch):
        """Convenience over a synth.triangulation_batch dict."""
        return self.run(batch["cam"], batch["track_ptr"], batch["track_uv"], batch["track_pose"],
                        batch["n_neighbours"], batch["nb_pos"], batch["nb_valid"], batch.get("scale", 1.0))

    def rerun(self):
        """Re-run the kernel on the HBM-resident staged batch; device ms."""
        ms = C.c_float(0)
        self._check(self.L.nrslam_b200_tri_rerun(self._h, C.byref(ms)))
        return ms.value
