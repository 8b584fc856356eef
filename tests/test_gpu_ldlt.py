"""GPU parity for the LDLt path through the C ABI: symbolic analysis bit-exact vs the oracle,
numeric factor and 2-RHS solves vs the oracle / a dense solve.

Bars (north_star / BASELINE.md §5): symbolic bit-exact; solve relative residual <= 1e-10;
multipliers (q) and p agree with the oracle to 1e-8 relative (observed ~1e-13)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

SQRT_EPS = float(np.sqrt(np.finfo(float).eps))


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _setup(A, P=None, **opts):
    import fpsb200
    coo = sp.coo_matrix(A)
    m, n = A.shape
    H = fpsb200.B200Handle(n, m, coo.row, coo.col)
    H.set_jac_values(coo.data)
    o = None
    if opts:
        o = fpsb200.LdltOpts()
        o.ldlt_tol, o.ldlt_r1, o.ldlt_r2 = SQRT_EPS, SQRT_EPS, -SQRT_EPS
        for k, v in opts.items():
            setattr(o, k, v)
    H.ldlt_analyze(P, 0, o)
    return H, coo


def _residuals(A, delta, r1, r2, p1, q1, p2, q2, kind):
    if kind == "mixed":
        e1 = np.r_[p1 + A.T @ q1 - r1, A @ p1 - delta * q1]
        e2 = np.r_[p2 + A.T @ q2, A @ p2 - delta * q2 - r2]
    else:
        e1 = np.r_[p1 + A.T @ q1 - r1, A @ p1 - delta * q1]
        e2 = np.r_[p2 + A.T @ q2 - r2, A @ p2 - delta * q2]
    return np.linalg.norm(e1) / np.linalg.norm(r1), np.linalg.norm(e2) / np.linalg.norm(r2)


CASES = [(1, 10, 9, 4, 0.0, None), (30, 60, 5, 8, 0.25, None), (30, 60, 5, 8, 0.0, "id"),
         (30, 60, 5, 8, 1e-2, "rand"), (500, 1000, 10, 32, 1e-2, None),
         (500, 1000, 10, 32, 0.0, "rand"), (5000, 10000, 10, 64, SQRT_EPS, None),
         (20000, 40000, 20, 64, 1e-8, None)]


@pytest.mark.parametrize("m,n,k,w,delta,pk", CASES)
def test_ldlt_mixed_and_least_squares(oracle, m, n, k, w, delta, pk):
    from fpsb200 import models
    A = models.window_random_jacobian(m, n, k, w=w, seed=17)
    rng = np.random.default_rng(5)
    P = None if pk is None else (np.arange(n + m) if pk == "id" else rng.permutation(n + m))
    H, coo = _setup(A, P)
    sym = H.ldlt_symbolic()
    lo = oracle.LDLtOracle(n, m, coo.row, coo.col, sym["P"])
    r1 = rng.standard_normal(n); r2 = rng.standard_normal(m); r3 = rng.standard_normal(n)
    p1, q1, p2, q2, ok = H.ldlt_solve_two_mixed(delta, r1, r2)
    op1, oq1, op2, oq2, ook = lo.solve_two_mixed(coo.data, delta, r1, r2)
    assert ok and ook
    # symbolic analysis: bit-exact
    osym = lo.symbolic()
    for key in ("P", "parent", "Lnz", "Lp", "Li"):
        assert np.array_equal(sym[key], osym[key]), key
    # numeric factor
    Lx, D = H.ldlt_get_factor()
    oLx, oD = lo.numeric()
    assert _rel(D, oD) < 1e-10 and _rel(Lx, oLx) < 1e-10
    # pivots replaced by +/-sqrt(eps) (dynamic regularisation, delta = 0) make the solve itself
    # ill-conditioned (cond ~ 1e8): the factors still agree to 1e-10 but the solutions only to ~1e-6
    perturbed = bool(np.any(np.abs(np.abs(oD) - SQRT_EPS) < 1e-12))
    bar = 1e-5 if perturbed else 1e-8
    for a, b in ((p1, op1), (q1, oq1), (p2, op2), (q2, oq2)):
        assert _rel(a, b) < bar
    if not perturbed:
        e1, e2 = _residuals(A, delta, r1, r2, p1, q1, p2, q2, "mixed")
        assert e1 < 1e-10 and e2 < 1e-10
    # solve-only path reuses the factorisation
    p1, q1, p2, q2, ok = H.ldlt_solve_two_least_squares(r1, r3)
    op1, oq1, op2, oq2, ook = lo.solve_two_least_squares(r1, r3)
    assert ok and ook
    for a, b in ((p1, op1), (q1, oq1), (p2, op2), (q2, oq2)):
        assert _rel(a, b) < bar


def test_ldlt_dense_ground_truth():
    from fpsb200 import models
    m, n, delta = 40, 90, 0.3
    A = models.window_random_jacobian(m, n, 7, w=10, seed=23)
    H, _ = _setup(A)
    Ad = A.toarray()
    K = np.block([[np.eye(n), Ad.T], [Ad, -delta * np.eye(m)]])
    rng = np.random.default_rng(0)
    r1 = rng.standard_normal(n); r2 = rng.standard_normal(m)
    p1, q1, p2, q2, ok = H.ldlt_solve_two_mixed(delta, r1, r2)
    s1 = np.linalg.solve(K, np.r_[r1, np.zeros(m)]); s2 = np.linalg.solve(K, np.r_[np.zeros(n), r2])
    assert ok
    assert _rel(np.r_[p1, q1], s1) < 1e-12 and _rel(np.r_[p2, q2], s2) < 1e-12


def test_ldlt_rank_deficient_dynamic_regularisation(oracle):
    """test/rank-deficient.jl analogue: duplicated Jacobian rows, delta = 0 -> the r2 = -sqrt(eps)
    pivot perturbation must fire identically on both sides."""
    from fpsb200 import models
    mdl = models.sparse_qp(2000, 1000, nnz_per_row=8, w=16, seed=3, rank_deficient_frac=0.02)
    A = mdl.A
    H, coo = _setup(A)
    sym = H.ldlt_symbolic()
    lo = oracle.LDLtOracle(2000, 1000, coo.row, coo.col, sym["P"])
    rng = np.random.default_rng(1)
    r1 = rng.standard_normal(2000); r2 = A @ rng.standard_normal(2000)
    got = H.ldlt_solve_two_mixed(0.0, r1, r2)
    ref = lo.solve_two_mixed(coo.data, 0.0, r1, r2)
    assert got[4] and ref[4]
    _, D = H.ldlt_get_factor(); _, oD = lo.numeric()
    reg = np.abs(np.abs(oD) - SQRT_EPS) < 1e-12
    assert reg.sum() >= 10                       # perturbed pivots exist
    assert np.array_equal(np.abs(np.abs(D) - SQRT_EPS) < 1e-12, reg)
    assert np.array_equal(np.sign(D), np.sign(oD))


def test_ldlt_failed_factorisation_passthrough():
    """D[k] == 0 with regularisation disabled -> factorized = false and `sol` holds the rhs
    (src/solve_linear_system.jl:242-251)."""
    A = sp.csr_matrix(np.array([[1.0, 0.0, 0.0], [1.0, 0.0, 0.0]]))   # duplicate rows, delta = 0
    H, _ = _setup(A, np.arange(5), ldlt_r1=0.0, ldlt_r2=0.0, ldlt_tol=0.0)
    r1 = np.array([1.0, 2.0, 3.0]); r2 = np.array([4.0, 5.0])
    p1, q1, p2, q2, ok = H.ldlt_solve_two_mixed(0.0, r1, r2)
    assert not ok
    assert np.array_equal(p1, r1) and not q1.any() and not p2.any() and np.array_equal(q2, r2)


def test_qdsolver_surface_golden(oracle):
    """Reference known answers through the plugin surface on the GPU (tests/golden/unit_test.json,
    from test/unit-test.jl:17-64, 100-126)."""
    import json, os
    import fpsb200
    from fpsb200 import models
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "unit_test.json")))
    for qds_name in ("ldlt", "iterative"):
        nlp = models.unit_test_model_sum(10)
        qds = fpsb200.qdsolver_correspondence[qds_name](nlp, 0.0)
        for approx in (1, 2):
            f = fpsb200.FletcherPenaltyNLP(nlp, 0.5, 0.0, 0.0, approx, qds=qds)
            tol = 1e-13 if qds_name == "ldlt" else 1e-7
            xf = np.array(G["G1"]["x"])
            assert abs(f.obj(xf) - G["G1"]["obj"]) < tol
            assert np.allclose(f.ys, G["G1"]["ys"], atol=tol)
            assert np.allclose(f.grad(xf), G["G1"]["grad"], atol=tol)
            xr = np.array(G["G2"]["x"])
            assert abs(f.obj(xr) - G["G2"]["obj"]) < 10 * tol
            assert np.allclose(f.ys, G["G2"]["ys"], atol=10 * tol)
            assert np.allclose(f.grad(xr), G["G2"]["grad"], atol=10 * tol)
            v = np.array(G["G4"]["v"])
            if approx == 2:
                assert np.allclose(f.hprod(xr, v), G["G4"]["hprod_val2"], atol=100 * tol)
        nlp = models.unit_test_model_rosenbrock_circle()
        qds = fpsb200.qdsolver_correspondence[qds_name](nlp, 0.0)
        f = fpsb200.FletcherPenaltyNLP(nlp, 0.5, 0.1, 0.25, 1, qds=qds)
        x = np.array(G["G3"]["x"])
        tol = 1e-13 if qds_name == "ldlt" else 1e-6
        assert abs(f.obj(x) - G["G3"]["obj"]) < 10 * tol
        assert np.allclose(f.gx, G["G3"]["gx"], atol=1e-12)
        assert np.allclose(f.ys, G["G3"]["ys"], atol=10 * tol)


def test_ldlt_dissection_ordering_and_pinned_host(oracle):
    """The B200-oriented ordering through the plugin surface (LDLtSolver(..., ordering="dissection"))
    gives the same solution as the oracle run with the same P; host buffers page-locked with
    fpsb_pin_host take the direct-DMA path and give identical results."""
    import fpsb200
    from fpsb200 import models
    mdl = models.sparse_qp(40000, 20000, nnz_per_row=10, w=32, seed=9)
    qds = fpsb200.LDLtSolver(mdl, 0.0, ordering="dissection")
    H = qds.handle
    info = H.ldlt_plan_info()
    assert info["nlevels"] < 400
    rows, cols = mdl.jac_structure()
    vals = mdl.jac_coord(None)
    H.set_jac_values(vals)
    rng = np.random.default_rng(2)
    r1 = rng.standard_normal(40000); r2 = rng.standard_normal(20000)
    p1, q1, p2, q2, ok = H.ldlt_solve_two_mixed(1e-3, r1, r2)
    assert ok
    sym = H.ldlt_symbolic()
    lo = oracle.LDLtOracle(40000, 20000, rows, cols, sym["P"])
    op1, oq1, op2, oq2, ook = lo.solve_two_mixed(vals, 1e-3, r1, r2)
    for a, b in ((p1, op1), (q1, oq1), (p2, op2), (q2, oq2)):
        assert _rel(a, b) < 1e-8
    A = mdl.A
    e1, e2 = _residuals(A, 1e-3, r1, r2, p1, q1, p2, q2, "mixed")
    assert e1 < 1e-10 and e2 < 1e-10
    # pinned host buffers: same numbers through the direct-DMA path
    bufs = [np.array(vals), np.array(r1), np.array(r2)]
    for b in bufs:
        H.pin_host(b)
    H.set_jac_values(bufs[0])
    pp1, qq1, pp2, qq2, ok = H.ldlt_solve_two_mixed(1e-3, bufs[1], bufs[2])
    assert ok and np.array_equal(pp1, p1) and np.array_equal(qq2, q2)
    got = H.iter_solve_two_mixed(1e-3, bufs[1], bufs[2])
    ref = H.iter_solve_two_mixed(1e-3, r1, r2)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[3], ref[3])
    for b in bufs:
        H.unpin_host(b)
