"""The C-ABI library: builds, loads, exports every symbol include/nrslam_b200.h declares, and refuses to compute
without a device (no CPU fallback). Host bookkeeping entry points (graph) are checked against the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nrslam_b200 import abi, api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(api.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return api.load()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "nrslam_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nrslam_b200_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_abi_version_and_default_options(lib):
    assert lib.nrslam_b200_abi_version() == 2
    o = api.default_options()
    assert abs(o.th_huber_2dof_sq - 5.99) < 1e-6 and abs(o.th_huber_3dof_sq - 0.584) < 1e-6
    assert list(o.pose_only_iterations) == [10, 10, 10] and list(o.pose_deform_iterations) == [10, 10]
    assert o.ba_iterations == 5 and o.lost_iterations == 10 and o.regularizers_per_point == 10
    assert o.lm_max_trials == 10 and o.lm_tau == 1e-5 and abs(o.spring_k - 1.1) < 1e-6


def test_struct_layouts_match_the_header(lib):
    # sizes the C compiler produced, reported through a tiny probe: stats / options / graph are plain PODs
    assert C.sizeof(abi.Camera) == 36
    assert C.sizeof(abi.Options) % 8 == 0 and C.sizeof(abi.Stats) % 8 == 0
    assert C.sizeof(abi.Graph) == 8 + 8 * 8 + 8


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    ctx = C.c_void_p()
    rc = lib.nrslam_b200_create(None, C.byref(ctx))
    assert rc == -1 and not ctx.value  # NRSLAM_B200_ERR_NO_DEVICE
    with pytest.raises(api.NrslamError):
        api.Core()
    # compute entry points reject a null context instead of computing anything
    assert lib.nrslam_b200_pose_only(None, None, 0, None, None, None, None, None) < 0
    assert lib.nrslam_b200_resolve(None, 0, None) < 0


def test_graph_entry_points_match_oracle(lib, oracle):
    p = synth.tracking_problem("c1", n=300)
    g1, g2 = p["graph"].copy(), p["graph"].copy()
    g1.status[::5] = abi.EDGE_BAD
    g2.status[::5] = abi.EDGE_BAD
    rng = np.random.default_rng(0)
    pos = p["last_world_position"] + rng.normal(scale=0.05, size=p["last_world_position"].shape).astype(np.float32)

    class Host:  # product host functions without a context
        L = lib
    for v in range(0, g1.n_vertices, 13):
        s1, s2 = g1.struct(), g2.struct()
        out = np.zeros(64, np.int32)
        n = lib.nrslam_b200_graph_get_edges(C.byref(s1), v, abi.ptr(out, C.c_int32), 64)
        assert np.array_equal(out[:n], oracle.graph_get_edges(g2, v))
        a = lib.nrslam_b200_graph_update_vertex(C.byref(s1), v, abi.ptr(pos, C.c_float))
        b = oracle.graph_update_vertex(g2, v, pos)
        assert a == b
    for name in ("weight", "min_distance", "max_distance", "status"):
        assert np.array_equal(getattr(g1, name), getattr(g2, name)), name
