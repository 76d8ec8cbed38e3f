"""ctypes mirror of include/nrslam_b200.h (structs only). Shared by the product binding (this package) and by
the oracle loader in tests/ so both sides are fed byte-identical buffers."""
import ctypes as C

import numpy as np

TRACE = 64

TRACKED_WITH_3D, TRACKED, JUST_TRIANGULATED, BAD, OUT_IMAGE_BOUNDARIES, BAD_FEATURE = range(6)
EDGE_VERIFIED, EDGE_NEIGHBOR, EDGE_NEUTRAL, EDGE_BAD = range(4)


class Camera(C.Structure):
    _fields_ = [("model", C.c_int32), ("params", C.c_float * 8)]

    @staticmethod
    def pinhole(fx, fy, cx, cy):
        c = Camera()
        c.model = 0
        c.params[:4] = [fx, fy, cx, cy]
        return c

    @staticmethod
    def kb8(fx, fy, cx, cy, k0, k1, k2, k3):
        c = Camera()
        c.model = 1
        c.params[:] = [fx, fy, cx, cy, k0, k1, k2, k3]
        return c


class Options(C.Structure):
    _fields_ = [
        ("th_huber_2dof_sq", C.c_float), ("th_huber_3dof_sq", C.c_float), ("sigma_reprojection", C.c_float),
        ("sigma_position", C.c_float), ("sigma_spatial_factor", C.c_float), ("spring_k", C.c_float),
        ("regularizers_per_point", C.c_int32), ("pose_only_iterations", C.c_int32 * 3),
        ("pose_deform_iterations", C.c_int32 * 2), ("lost_iterations", C.c_int32), ("ba_iterations", C.c_int32),
        ("lm_max_trials", C.c_int32), ("lm_tau", C.c_double), ("pcg_rel_tol", C.c_double),
        ("pcg_max_iterations", C.c_int32), ("device", C.c_int32), ("grid_ctas", C.c_int32),
    ]


class Graph(C.Structure):
    _fields_ = [
        ("n_vertices", C.c_int32), ("n_edges", C.c_int32), ("rowptr", C.POINTER(C.c_int32)),
        ("col", C.POINTER(C.c_int32)), ("eid", C.POINTER(C.c_int32)), ("weight", C.POINTER(C.c_float)),
        ("first_distance", C.POINTER(C.c_float)), ("min_distance", C.POINTER(C.c_float)),
        ("max_distance", C.POINTER(C.c_float)), ("status", C.POINTER(C.c_uint8)), ("weight_sigma", C.c_float),
        ("stretching_th", C.c_float),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("lm_iterations", C.c_int32), ("lm_trials", C.c_int32), ("pcg_iterations", C.c_int32),
        ("n_sweeps", C.c_int32), ("n_chi2_passes", C.c_int32), ("n_reproj_edges", C.c_int32),
        ("n_pair_edges", C.c_int32), ("n_spring_edges", C.c_int32), ("n_damper_edges", C.c_int32),
        ("n_fixed_edges", C.c_int32), ("n_points", C.c_int32), ("n_poses", C.c_int32),
        ("kernel_launches", C.c_int32), ("n_trace", C.c_int32), ("chi2_trace", C.c_double * TRACE),
        ("lambda_final", C.c_double), ("gpu_ms", C.c_float), ("host_ms", C.c_float), ("stage_ms", C.c_float),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("grid_ctas", C.c_int32), ("block_threads", C.c_int32),
        ("direct_solves", C.c_int32), ("solve_failures", C.c_int32), ("factor_doubles", C.c_int64),
        ("update_doubles", C.c_int64), ("plan_reused", C.c_int32), ("reserved0", C.c_int32),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "chi2_trace"}
        d["chi2_trace"] = list(self.chi2_trace[: self.n_trace])
        return d


class MaskFilter(C.Structure):
    _fields_ = [("kind", C.c_int32), ("th", C.c_int32), ("rb", C.c_int32), ("re", C.c_int32), ("cb", C.c_int32),
                ("ce", C.c_int32), ("mask", C.POINTER(C.c_uint8))]


FILTER_BRIGHT, FILTER_BORDER, FILTER_PREDEFINED = range(3)


_PTR_TYPES = {}


def _pointer_type(ctype):
    t = _PTR_TYPES.get(ctype)
    if t is None:
        t = _PTR_TYPES[ctype] = C.POINTER(ctype)
    return t


def ptr(a, ctype):
    """numpy array -> typed ctypes pointer (None passes NULL)."""
    if a is None:
        return None
    # (ndarray.ctypes builds a helper object per call: 5 us; a frame passes two dozen pointers)
    p = C.cast(a.__array_interface__["data"][0], _pointer_type(ctype))
    p._keep = a  # a temporary passed straight into a call stays alive as long as its pointer
    return p


class GraphArrays:
    """Owns the numpy arrays behind a Graph struct (keeps them alive and exposes them for comparison)."""

    def __init__(self, rowptr, col, eid, weight, first_distance, min_distance, max_distance, status, weight_sigma,
                 stretching_th=1.1):
        self.rowptr = np.ascontiguousarray(rowptr, np.int32)
        self.col = np.ascontiguousarray(col, np.int32)
        self.eid = np.ascontiguousarray(eid, np.int32)
        self.weight = np.ascontiguousarray(weight, np.float32)
        self.first_distance = np.ascontiguousarray(first_distance, np.float32)
        self.min_distance = np.ascontiguousarray(min_distance, np.float32)
        self.max_distance = np.ascontiguousarray(max_distance, np.float32)
        self.status = np.ascontiguousarray(status, np.uint8)
        self.weight_sigma = float(weight_sigma)
        self.stretching_th = float(stretching_th)

    def copy(self):
        return GraphArrays(self.rowptr.copy(), self.col.copy(), self.eid.copy(), self.weight.copy(),
                           self.first_distance.copy(), self.min_distance.copy(), self.max_distance.copy(),
                           self.status.copy(), self.weight_sigma, self.stretching_th)

    @property
    def n_vertices(self):
        return len(self.rowptr) - 1

    @property
    def n_edges(self):
        return len(self.weight)

    def struct(self):
        g = Graph()
        g.n_vertices = self.n_vertices
        g.n_edges = self.n_edges
        g.rowptr = ptr(self.rowptr, C.c_int32)
        g.col = ptr(self.col, C.c_int32)
        g.eid = ptr(self.eid, C.c_int32)
        g.weight = ptr(self.weight, C.c_float)
        g.first_distance = ptr(self.first_distance, C.c_float)
        g.min_distance = ptr(self.min_distance, C.c_float)
        g.max_distance = ptr(self.max_distance, C.c_float)
        g.status = ptr(self.status, C.c_uint8)
        g.weight_sigma = self.weight_sigma
        g.stretching_th = self.stretching_th
        return g
