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
