"""Multi-segment gather windows (stencil operators, BASELINE config C3): the tiles of a wide 5-point-stencil operator
stage the few runs of columns they touch as separate bulk copies; products and Krylov iterations must agree with
scipy / the oracle exactly as the single-window and global-gather tiles do.

Reference interface under test: `jac_op!` products and `solve_two_mixed` of the `IterativeSolver`
(src/solve_linear_system.jl:119-150)."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def _opts(n, m, **kw):
    import fpsb200
    o = fpsb200.IterOpts()
    assert fpsb200._lib.lib().fpsb_iter_default_opts(ctypes.c_int64(n), ctypes.c_int64(m), ctypes.byref(o)) == 0
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _stencil(nx, ny, interleave):
    """A = [L  -I] of an nx x ny grid (nx = the stride between grid lines), columns interleaved (y_i, u_i) or blocked."""
    Tx = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(nx, nx))
    Ty = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(ny, ny))
    L = sp.kron(Ty, sp.identity(nx)) + sp.kron(sp.identity(ny), Tx)
    m = nx * ny
    A = sp.hstack([L, -sp.identity(m)]).tocoo()
    rng = np.random.default_rng(11)
    vals = A.data * (1.0 + 0.1 * rng.standard_normal(A.nnz))        # not symmetric, not constant: catches index mix-ups
    col = A.col.astype(np.int64)
    if interleave:
        perm = np.empty(2 * m, dtype=np.int64)
        perm[:m] = 2 * np.arange(m)
        perm[m:] = 2 * np.arange(m) + 1
        col = perm[col]
    return m, 2 * m, A.row.astype(np.int64), col, vals


@pytest.mark.parametrize("interleave", [True, False])
def test_stencil_tiles_are_multi_segment_and_exact(interleave):
    import torch
    import fpsb200
    m, n, jr, jc, vals = _stencil(1100, 40, interleave)
    A = sp.csr_matrix((vals, (jr, jc)), shape=(m, n))
    H = fpsb200.B200Handle(n, m, jr, jc)
    H.set_jac_values(vals)
    st = H.tile_stats()
    # nearly every tile of A and the stencil rows of A' need more than one segment; none is left on global gathers
    assert st["A"]["multi_segment"] >= 0.9 * st["A"]["tiles"], st
    assert st["A"]["windowed"] == st["A"]["tiles"], st
    assert st["At"]["windowed"] == st["At"]["tiles"], st
    assert st["At"]["multi_segment"] >= 0.4 * st["At"]["tiles"], st
    rng = np.random.default_rng(5)
    x, u = rng.standard_normal(n), rng.standard_normal(m)
    assert _rel(H.jprod(x), A @ x) < 1e-13
    assert _rel(H.jtprod(u), A.T @ u) < 1e-13
    X2, U2 = rng.standard_normal((2, n)), rng.standard_normal((2, m))
    Y2 = np.asarray(H.jprod2(X2.ravel())).reshape(2, m)
    Z2 = np.asarray(H.jtprod2(U2.ravel())).reshape(2, n)
    for k in range(2):
        assert _rel(Y2[k], A @ X2[k]) < 1e-13 and _rel(Z2[k], A.T @ U2[k]) < 1e-13
    # caller vectors that are only 8-byte aligned: the consumer groups stage the segments themselves
    xt = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
    xt[1:] = torch.as_tensor(x, device="cuda")
    y = H.jprod(xt[1:])
    assert _rel(y.cpu().numpy(), A @ x) < 1e-13
    ut = torch.zeros(m + 1, dtype=torch.float64, device="cuda")
    ut[1:] = torch.as_tensor(u, device="cuda")
    assert _rel(H.jtprod(ut[1:]).cpu().numpy(), A.T @ u) < 1e-13


def test_stencil_krylov_matches_oracle(oracle):
    """LSQR / CRAIG recurrences over the multi-segment tiles at fixed iteration counts (persistent loop kernel and
    launch-per-half-iteration kernels) against the CPU oracle, and the converged solve."""
    import os
    import fpsb200
    m, n, jr, jc, vals = _stencil(900, 36, True)
    A = sp.csr_matrix((vals, (jr, jc)), shape=(m, n))
    rng = np.random.default_rng(9)
    r1, r2 = rng.standard_normal(n), rng.standard_normal(m)
    kw = dict(ls_itmax=25, ln_itmax=25)
    ref = oracle.IterativeOracle(A, **kw).solve_two_mixed(1e-2, r1, r2)
    H = fpsb200.B200Handle(n, m, jr, jc)
    assert H.tile_stats()["A"]["multi_segment"] > 0
    H.iter_setup(_opts(n, m, **kw))
    H.set_jac_values(vals)
    got = H.iter_solve_two_mixed(1e-2, r1, r2)
    assert [s["niter"] for s in got[4]] == [s["niter"] for s in ref[4]]
    for a, b in zip(got[:4], ref[:4]):
        assert _rel(a, b) < 1e-10
    # least-squares pair (two LSQRs in lock step) and the extras
    r3 = rng.standard_normal(n)
    got = H.iter_solve_two_least_squares(1e-2, r1, r3)
    ref = oracle.IterativeOracle(A, **kw).solve_two_least_squares(1e-2, r1, r3)
    for a, b in zip(got[:4], ref[:4]):
        assert _rel(a, b) < 1e-10
    # converged at the reference tolerances
    H2 = fpsb200.B200Handle(n, m, jr, jc)
    H2.iter_setup(_opts(n, m))
    H2.set_jac_values(vals)
    full = H2.iter_solve_two_mixed(1e-2, r1, r2)
    oref = oracle.IterativeOracle(A).solve_two_mixed(1e-2, r1, r2)
    for s, o in zip(full[4], oref[4]):
        assert s["solved"] == o["solved"] and abs(s["niter"] - o["niter"]) <= 1
    for a, b in zip(full[:4], oref[:4]):
        assert _rel(a, b) < 1e-6


def test_stencil_partitioned_handle_world1(oracle):
    from fpsb200.partition import RowPartition, DistHandle
    m, n, jr, jc, vals = _stencil(800, 40, True)
    A = sp.csr_matrix((vals, (jr, jc)), shape=(m, n))
    rng = np.random.default_rng(2)
    r1, r2 = rng.standard_normal(n), rng.standard_normal(m)
    kw = dict(ls_itmax=20, ln_itmax=20)
    D = DistHandle(RowPartition(n, m, jr, jc, 1), 0, device=0, opts=_opts(n, m, **kw))
    D.set_jac_values(vals)
    assert _rel(D.jprod(r1), A @ r1) < 1e-13
    assert _rel(D.jtprod(r2), A.T @ r2) < 1e-13
    got = D.solve_two_mixed(1e-2, r1, r2)
    ref = oracle.IterativeOracle(A, **kw).solve_two_mixed(1e-2, r1, r2)
    assert [s["niter"] for s in got[4]] == [s["niter"] for s in ref[4]]
    for a, b in zip(got[:4], ref[:4]):
        assert _rel(a, b) < 1e-9
