"""Pins the oracle's exact linear solve against g2o's own known-answer vectors
(third_party/g2o/unit_test/solver/linear_solver_test.cpp:73-87, fixture sparse_system_helper.cpp:52-340;
extracted by tests/golden/make_g2o_kat.py). Tolerance is g2o's own: isApprox(1e-6)."""
import ctypes as C
import json
import os

import numpy as np

import oracle_lib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g2o_linear_solver_kat.json")


def _system():
    k = json.load(open(GOLD))
    rows, cols, vals = [], [], []
    for b in k["blocks"]:
        r, c, m = b["r"], b["c"], np.array(b["m"])
        for a in range(3):
            for d in range(3):
                if r == c and d < a:
                    continue
                rows.append(3 * r + a)
                cols.append(3 * c + d)
                vals.append(m[a, d])
    return k, np.array(rows, np.int32), np.array(cols, np.int32), np.array(vals, np.float64)


def _solve(rows, cols, vals, b, block):
    L = oracle_lib.lib()
    x = np.zeros(36)
    rc = L.orc_sparse_solve(36, block, len(vals), rows.ctypes.data_as(C.POINTER(C.c_int32)),
                            cols.ctypes.data_as(C.POINTER(C.c_int32)), vals.ctypes.data_as(C.POINTER(C.c_double)),
                            b.ctypes.data_as(C.POINTER(C.c_double)), x.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    return x


def test_sparse_cholesky_matches_g2o_golden_solution():
    k, rows, cols, vals = _system()
    b, xg = np.array(k["b"]), np.array(k["x"])
    for block in (3, 1):  # with / without block ordering, like the g2o typed test
        x = _solve(rows, cols, vals, b, block)
        assert np.linalg.norm(x - xg) <= 1e-6 * min(np.linalg.norm(x), np.linalg.norm(xg))


def test_sparse_cholesky_matches_g2o_golden_inverse():
    k, rows, cols, vals = _system()
    inv = np.array(k["inverse"])
    for col in (0, 7, 35):
        e = np.zeros(36)
        e[col] = 1.0
        x = _solve(rows, cols, vals, e, 3)
        assert np.allclose(x, inv[:, col], rtol=1e-9, atol=1e-15)
