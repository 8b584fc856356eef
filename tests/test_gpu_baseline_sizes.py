"""GPU parity at the FULL sizes of the BASELINE.json configs (through the C ABI, against the CPU oracle).

  headline / C4 shape  n = 1e6, m = 5e5, nnz = 1e7   Krylov path (solve_two_mixed):
        reference tolerances sqrt(eps): same `solved`, iterations +/-1, iterates <= 1e-6 (the level two
        sqrt(eps)-stopped iterations can agree to);  tightened tolerances (1e-13): the least-norm half (p2, q2)
        <= 1e-8 against the oracle run the same way and K-residual <= 1e-10 (north_star's bars); the
        least-squares half keeps Krylov.jl's own axtol / btol, which the reference's call site cannot tighten
  C2   n = 1e5, m = 5e4, 10 nnz/row                   LDLt path: solutions <= 1e-8, K-residual <= 1e-10
  C4   rank-deficient n = 1e5, m = 5e4, delta = 0     LDLt: identical mask of regularised pivots, factor 1e-10
  C3   Poisson control 512^2 (n = 524 288)            row-partitioned handle (world 1) vs oracle
Sizes are chosen so that the sequential oracle finishes each case in seconds."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

SQRT_EPS = float(np.sqrt(np.finfo(float).eps))


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _opts(n, m, **kw):
    import fpsb200
    o = fpsb200.IterOpts()
    assert fpsb200._lib.lib().fpsb_iter_default_opts(ctypes.c_int64(n), ctypes.c_int64(m), ctypes.byref(o)) == 0
    for k, v in kw.items():
        setattr(o, k, v)
    return o


@pytest.fixture(scope="module")
def headline():
    import fpsb200
    from fpsb200 import models
    n, m = 1_000_000, 500_000
    A = models.window_random_jacobian(m, n, 20, w=64, seed=1234)
    coo = A.tocoo()
    rng = np.random.default_rng(1234)
    H = fpsb200.B200Handle(n, m, coo.row.astype(np.int64), coo.col.astype(np.int64))
    H.set_jac_values(coo.data)
    return A, H, rng.standard_normal(n), rng.standard_normal(m)


def test_headline_krylov_reference_tolerances(oracle, headline):
    """BASELINE.json's metric config, the exact solve bench.py times: GPU vs the sequential oracle."""
    A, H, r1, r2 = headline
    n, m = A.shape[1], A.shape[0]
    H.iter_setup(None)
    got = H.iter_solve_two_mixed(0.0, r1, r2)
    ref = oracle.IterativeOracle(A).solve_two_mixed(0.0, r1, r2)
    for s, o in zip(got[4], ref[4]):
        assert s["solved"] == o["solved"] is True
        assert abs(s["niter"] - o["niter"]) <= 1
    for a, b in zip(got[:4], ref[:4]):
        assert _rel(a, b) < 1e-6
    p1, q1, p2, q2 = got[:4]
    assert _rel(p1 + A.T @ q1, r1) < 1e-6 and np.linalg.norm(A @ p1) < 1e-6 * np.linalg.norm(r1)
    assert np.linalg.norm(p2 + A.T @ q2) < 1e-6 * np.linalg.norm(r2) and _rel(A @ p2, r2) < 1e-6
    # the two-LSQR variant on the same operator
    r3 = np.random.default_rng(7).standard_normal(n)
    got = H.iter_solve_two_least_squares(0.0, r1, r3)
    ref = oracle.IterativeOracle(A).solve_two_least_squares(0.0, r1, r3)
    for s, o in zip(got[4], ref[4]):
        assert s["solved"] == o["solved"] and abs(s["niter"] - o["niter"]) <= 1
    for a, b in zip(got[:4], ref[:4]):
        assert _rel(a, b) < 1e-6


@pytest.mark.parametrize("n", [450_000, 500_000])
def test_ring_release_order_mid_sizes(oracle, n):
    """Regression (round 2): with 4 ring stages and 3 consumer groups a group that got two tiles ahead of its
    neighbours read a stage's previous contents (mbarrier parity one phase ahead).  It showed at n = 450-550 K with the
    early row sums of the persistent loop kernel, never at the headline size: repeated solves must be bitwise equal and
    agree with the oracle."""
    import fpsb200
    from fpsb200 import models
    m = n // 2
    A = models.window_random_jacobian(m, n, 20, w=64, seed=1234)
    coo = A.tocoo()
    rng = np.random.default_rng(1234)
    r1, r2 = rng.standard_normal(n), rng.standard_normal(m)
    H = fpsb200.B200Handle(n, m, coo.row.astype(np.int64), coo.col.astype(np.int64))
    H.set_jac_values(coo.data)
    outs = [H.iter_solve_two_mixed(0.0, r1, r2) for _ in range(4)]
    for o in outs[1:]:
        for a, b in zip(o[:4], outs[0][:4]):
            assert np.array_equal(a, b)
    ref = oracle.IterativeOracle(A).solve_two_mixed(0.0, r1, r2)
    for s, o in zip(outs[0][4], ref[4]):
        assert s["solved"] == o["solved"] is True and abs(s["niter"] - o["niter"]) <= 1
    for a, b in zip(outs[0][:4], ref[:4]):
        assert _rel(a, b) < 1e-6


def test_headline_krylov_tight_tolerances_1e8(oracle, headline):
    """north_star's floating-point bars on the Krylov path.  With the IterativeSolver tolerances tightened the
    least-norm half (CRAIG: p2, q2 -> the `v`, `w` blocks of ys / gs) converges to the solution itself on both
    sides: <= 1e-8 relative against the oracle and K-residual <= 1e-10.  The least-squares half (LSQR) keeps
    Krylov.jl's axtol = btol = sqrt(eps), which the reference's call site cannot override
    (src/solve_linear_system.jl:115-124 passes atol / rtol / itmax only): it stops at the same iteration as at
    the default tolerances on both sides, so its bar stays the sqrt(eps)-level 1e-6."""
    A, H, r1, r2 = headline
    n, m = A.shape[1], A.shape[0]
    kw = dict(ls_atol=1e-13, ls_rtol=1e-13, ln_atol=1e-13, ln_rtol=1e-13, ln_btol=1e-13)
    H.iter_setup(_opts(n, m, **kw))
    try:
        got = H.iter_solve_two_mixed(0.0, r1, r2)
    finally:
        H.iter_setup(None)
    ref = oracle.IterativeOracle(A, **kw).solve_two_mixed(0.0, r1, r2)
    for s, o in zip(got[4], ref[4]):
        assert s["solved"] == o["solved"] is True
        assert abs(s["niter"] - o["niter"]) <= 1
    assert got[4][1]["niter"] > 100                      # the tightened tolerances really were in force
    p1, q1, p2, q2 = got[:4]
    assert _rel(p2, ref[2]) < 1e-8 and _rel(q2, ref[3]) < 1e-8
    assert _rel(p1, ref[0]) < 1e-6 and _rel(q1, ref[1]) < 1e-6
    res2 = np.linalg.norm(np.r_[p2 + A.T @ q2, A @ p2 - r2]) / np.linalg.norm(r2)
    assert res2 < 1e-10


def test_c2_full_ldlt_vs_oracle(oracle):
    import fpsb200
    from fpsb200 import models
    from fpsb200.symbolic import order_dissection
    n, m = 100_000, 50_000
    qp = models.sparse_qp(n, m, nnz_per_row=10, w=64, seed=1234)
    A = qp.A.tocsr()
    coo = A.tocoo()
    jr, jc = coo.row.astype(np.int64), coo.col.astype(np.int64)
    rng = np.random.default_rng(1234)
    r1, r2, r3 = rng.standard_normal(n), rng.standard_normal(m), rng.standard_normal(n)
    H = fpsb200.B200Handle(n, m, jr, jc)
    H.set_jac_values(coo.data)
    H.ldlt_analyze(order_dissection(n, m, jr, jc))
    sym = H.ldlt_symbolic()
    lo = oracle.LDLtOracle(n, m, jr, jc, sym["P"])
    got = H.ldlt_solve_two_mixed(SQRT_EPS, r1, r2)
    ref = lo.solve_two_mixed(coo.data, SQRT_EPS, r1, r2)
    assert got[4] and ref[4]
    osym = lo.symbolic()
    for key in ("parent", "Lnz", "Lp", "Li"):
        assert np.array_equal(sym[key], osym[key]), key
    for a, b in zip(got[:4], ref[:4]):
        assert _rel(a, b) < 1e-8
    p1, q1, p2, q2 = got[:4]
    e1 = np.linalg.norm(np.r_[p1 + A.T @ q1 - r1, A @ p1 - SQRT_EPS * q1]) / np.linalg.norm(r1)
    e2 = np.linalg.norm(np.r_[p2 + A.T @ q2, A @ p2 - SQRT_EPS * q2 - r2]) / np.linalg.norm(r2)
    assert e1 < 1e-10 and e2 < 1e-10
    got = H.ldlt_solve_two_least_squares(r1, r3)
    ref = lo.solve_two_least_squares(r1, r3)
    for a, b in zip(got[:4], ref[:4]):
        assert _rel(a, b) < 1e-8


def test_c4_rank_deficient_1e5_same_regularised_pivots(oracle):
    """test/rank-deficient.jl at BASELINE C4's shape (1 % duplicated rows, delta = 0): the dynamic regularisation
    must replace exactly the same pivots on both sides."""
    import fpsb200
    from fpsb200 import models
    from fpsb200.symbolic import order_dissection
    n, m = 100_000, 50_000
    qp = models.sparse_qp(n, m, nnz_per_row=20, w=64, seed=1234, rank_deficient_frac=0.01)
    A = qp.A.tocsr()
    coo = A.tocoo()
    jr, jc = coo.row.astype(np.int64), coo.col.astype(np.int64)
    rng = np.random.default_rng(7)
    r1, r2 = rng.standard_normal(n), A @ rng.standard_normal(n)
    H = fpsb200.B200Handle(n, m, jr, jc)
    H.set_jac_values(coo.data)
    H.ldlt_analyze(order_dissection(n, m, jr, jc))
    lo = oracle.LDLtOracle(n, m, jr, jc, H.ldlt_symbolic()["P"])
    got = H.ldlt_solve_two_mixed(0.0, r1, r2)
    ref = lo.solve_two_mixed(coo.data, 0.0, r1, r2)
    assert got[4] and ref[4]
    Lx, D = H.ldlt_get_factor()
    oLx, oD = lo.numeric()
    reg, oreg = np.abs(D) == SQRT_EPS, np.abs(oD) == SQRT_EPS
    assert oreg.sum() >= 0.005 * m                        # the duplicated rows really trigger it
    assert np.array_equal(reg, oreg)
    assert np.array_equal(np.sign(D[reg]), np.sign(oD[oreg]))
    # a regularised pivot divides by sqrt(eps): what sits in those columns is roundoff / sqrt(eps), so the
    # factors are compared where the arithmetic is well conditioned and the solutions through their residuals
    assert _rel(D[~reg], oD[~oreg]) < 1e-8
    p1, q1, p2, q2 = got[:4]
    assert _rel(p1 + A.T @ q1, r1) < 1e-8 and _rel(A @ p2, r2) < 1e-6
    op1, oq1, op2, oq2 = ref[:4]
    assert _rel(p1, op1) < 1e-5 and _rel(p2, op2) < 1e-5


def test_c3_poisson_512_partitioned_handle_vs_oracle(oracle):
    """BASELINE config C3 (Poisson-constrained control, here a 512 x 512 grid) through the row-partitioned
    handle with one rank: products to 1e-13, fixed-iteration recurrences vs the oracle, and the converged
    solve at the reference tolerances."""
    import fpsb200
    from fpsb200 import models
    from fpsb200.partition import RowPartition, DistHandle
    N = 512
    A0 = models.poisson_control(N).A.tocsr()
    m, n = A0.shape
    perm = np.empty(n, dtype=np.int64)
    perm[:m] = 2 * np.arange(m)
    perm[m:] = 2 * np.arange(m) + 1
    coo = A0.tocoo()
    jr, jc, vals = coo.row.astype(np.int64), perm[coo.col], coo.data.astype(np.float64)
    A = sp.csr_matrix((vals, (jr, jc)), shape=(m, n))
    rng = np.random.default_rng(3)
    r1, r2 = rng.standard_normal(n), rng.standard_normal(m)
    kw = dict(ls_itmax=30, ln_itmax=30)
    D = DistHandle(RowPartition(n, m, jr, jc, 1), 0, device=0, opts=_opts(n, m, **kw))
    D.set_jac_values(vals)
    assert _rel(D.jprod(r1), A @ r1) < 1e-13
    assert _rel(D.jtprod(r2), A.T @ r2) < 1e-13
    delta = 1e-2
    got = D.solve_two_mixed(delta, r1, r2)
    ref = oracle.IterativeOracle(A, **kw).solve_two_mixed(delta, r1, r2)
    assert [s["niter"] for s in got[4]] == [s["niter"] for s in ref[4]]
    for a, b in zip(got[:4], ref[:4]):
        assert _rel(a, b) < 1e-9
    # the plain single-GPU handle on the same operator (what the partitioned record of bench.py compares with)
    H = fpsb200.B200Handle(n, m, jr, jc)
    H.iter_setup(_opts(n, m, **kw))
    H.set_jac_values(vals)
    one = H.iter_solve_two_mixed(delta, r1, r2)
    for a, b in zip(got[:4], one[:4]):
        assert _rel(a, b) < 1e-9
