"""Minimal stand-in for qpsolvers 1.8.0 -> quadprog (test infrastructure; see ../README.md).

Solves   min 1/2 x'Px + q'x   s.t.  A x = b   through the KKT system with a dense
LU solve.  The reference calls it with one equality row (scripts/helper.py:182).
"""
import numpy as np


def solve_qp(P, q, G=None, h=None, A=None, b=None, lb=None, ub=None, solver=None, **kw):
    assert G is None and h is None and lb is None and ub is None
    P = np.asarray(P, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64).reshape(-1)
    n = P.shape[0]
    if A is None:
        return np.linalg.solve(P, -q)
    A = np.asarray(A, dtype=np.float64)
    if A.ndim == 1:                       # qpsolvers reshapes a 1-D A to one row
        A = A.reshape(1, -1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    m = A.shape[0]
    K = np.zeros((n + m, n + m))
    K[:n, :n] = P
    K[:n, n:] = A.T
    K[n:, :n] = A
    rhs = np.concatenate([-q, b])
    return np.linalg.solve(K, rhs)[:n]
