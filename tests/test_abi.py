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


def test_graph_store_add_edges_matches_the_reference_semantics(lib, oracle):
    """RegularizationGraph::AddEdge / SetSigma (map/regularization_graph.cc:33-55) through the owning store, against a
    dictionary restatement: distance = |relative position| in fp32, weight = exp(-d^2 / (2 sigma^2)), status NEUTRAL,
    min = max = first distance; inserting an existing pair again REPLACES its record; rows and row entries of the CSR
    view ascend. Then the view drives the host UpdateVertex / GetEdges entry points like any caller-built CSR."""
    rng = np.random.default_rng(5)
    M = 60
    pos = rng.normal(size=(M, 3)).astype(np.float32)
    sigma = np.float32(0.8)
    st = api.GraphStore(float(sigma), 1.1)
    ref = {}

    def weight(d, s):
        return np.float32(np.exp(-(d * d) / (np.float32(2) * s * s)))

    def add(pairs, s):
        v1 = np.array([p[0] for p in pairs], np.int32)
        v2 = np.array([p[1] for p in pairs], np.int32)
        rel = pos[v2] - pos[v1]
        assert st.add_edges(v1, v2, rel) == 0
        for a, b, r in zip(v1, v2, rel):
            d = np.float32(np.sqrt(np.float32(r[0] * r[0] + r[1] * r[1]) + np.float32(r[2] * r[2])))
            ref[(min(a, b), max(a, b))] = d

    # Map::InitializeRegularizationGraph: every pair of the first 25 landmarks (map.cc:139-167)
    add([(i, j) for i in range(25) for j in range(i + 1, 25)], sigma)
    g = st.arrays()
    assert g.n_vertices == 25 and g.n_edges == 300
    # Mapping: newly triangulated landmarks 25..59 connect to everything before them, in either argument order
    add([(j, i) if (i + j) % 2 else (i, j) for i in range(25, M) for j in range(0, i, 3)], sigma)
    g = st.arrays()
    assert g.n_vertices == M and g.n_edges == len(ref)
    # CSR view: rows ascending, entries ascending, both directions share one record
    for v in range(M):
        row = g.col[g.rowptr[v]:g.rowptr[v + 1]]
        assert np.all(np.diff(row) > 0)
        for p_ in range(g.rowptr[v], g.rowptr[v + 1]):
            e = g.eid[p_]
            d = ref[(min(v, g.col[p_]), max(v, g.col[p_]))]
            assert abs(g.first_distance[e] - d) <= 2e-7 * max(d, 1) and g.min_distance[e] == g.max_distance[e] == g.first_distance[e]
            assert abs(g.weight[e] - weight(g.first_distance[e], sigma)) <= 1e-6
            assert g.status[e] == abi.EDGE_NEUTRAL
    # the view behaves like a caller-built CSR: UpdateVertex writes straight into the store
    moved = pos + rng.normal(scale=0.3, size=pos.shape).astype(np.float32)
    g_ref = st.arrays()
    s1 = st.struct()
    for v in range(0, M, 7):
        a = lib.nrslam_b200_graph_update_vertex(C.byref(s1), v, abi.ptr(moved, C.c_float))
        assert a == oracle.graph_update_vertex(g_ref, v, moved)
    after = st.arrays()
    for name in ("weight", "min_distance", "max_distance", "status"):
        assert np.array_equal(getattr(after, name), getattr(g_ref, name)), name
    assert (after.max_distance > after.first_distance).any()
    # AddEdge on an existing pair replaces the record; SetSigma changes later weights only
    st.set_sigma(1.6)
    e_before = after.n_edges
    add([(3, 40)], np.float32(1.6))
    again = st.arrays()
    assert again.n_edges == e_before and again.weight_sigma == np.float32(1.6)
    row = list(again.col[again.rowptr[3]:again.rowptr[4]])
    e = again.eid[again.rowptr[3] + row.index(40)]
    assert again.min_distance[e] == again.max_distance[e] == again.first_distance[e]
    assert abs(again.weight[e] - weight(again.first_distance[e], np.float32(1.6))) <= 1e-6
    assert again.status[e] == abi.EDGE_NEUTRAL
    untouched = np.ones(again.n_edges, bool)
    untouched[e] = False
    assert np.array_equal(again.weight[untouched], after.weight[untouched])
    # argument errors modify nothing
    assert st.add_edges([1, 2], [1, 5], np.zeros((2, 3), np.float32)) < 0
    assert st.add_edges([-1], [5], np.zeros((1, 3), np.float32)) < 0
    assert st.arrays().n_edges == e_before
    st.close()
