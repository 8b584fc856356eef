"""CPU: pins the oracle (oracle/fps_oracle.c) before anything is checked against it.

  * every known answer the reference's tests hold for the path (tests/golden/unit_test.json,
    from test/unit-test.jl) — through the same FletcherPenaltyNLP host logic the GPU path uses;
  * an independent dense numpy.linalg.solve on K for all six solve_two_* methods;
  * scipy.sparse.linalg lsqr as a third opinion on the Krylov restatements.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import fpsb200
from fpsb200 import models
from oracle_qds import OracleIterative, OracleLDLt

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "unit_test.json")))
SQRT_EPS = float(np.sqrt(np.finfo(float).eps))


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("qds_cls,tol", [(OracleLDLt, 1e-13), (OracleIterative, 1e-7)])
@pytest.mark.parametrize("approx", [1, 2])
def test_golden_unit_test_sum_model(oracle, qds_cls, tol, approx):
    """test/unit-test.jl:16-76 and :154-214 (G1, G2, G4)."""
    nlp = models.unit_test_model_sum(10)
    f = fpsb200.FletcherPenaltyNLP(nlp, 0.5, 0.0, 0.0, approx, qds=qds_cls(nlp, 0.0))
    assert np.isnan(f.fx)
    xf = np.array(G["G1"]["x"])
    assert abs(f.obj(xf) - G["G1"]["obj"]) < max(tol, 1e-14)
    assert abs(f.fx - G["G1"]["fx"]) < 1e-14
    assert np.allclose(f.gx, G["G1"]["gx"], atol=1e-14)
    assert np.allclose(f.ys, G["G1"]["ys"], atol=max(tol, 1e-14))
    assert np.allclose(f.cx, G["G1"]["cx"], atol=1e-14)
    assert np.allclose(f.grad(xf), G["G1"]["grad"], atol=max(tol, 1e-14))
    xr = np.array(G["G2"]["x"])
    assert np.array_equal(nlp.cons(xr), G["G2"]["cx"])
    assert abs(f.obj(xr) - G["G2"]["obj"]) < 10 * tol
    assert np.allclose(f.grad(xr), G["G2"]["grad"], atol=10 * tol)
    v = np.array(G["G4"]["v"])
    # Val(2) closed form (unit-test.jl:191, 202); for this model (linear constraint, so the
    # solve_two_extras terms vanish) Val(1) gives the same product
    bar = 100 * tol if approx == 2 else 1e-6
    assert np.allclose(f.hprod(xf, v), G["G4"]["hprod_val2"], atol=bar)
    assert np.allclose(f.hprod(xr, v), G["G4"]["hprod_val2"], atol=bar)


@pytest.mark.parametrize("qds_cls,tol", [(OracleLDLt, 1e-13), (OracleIterative, 1e-6)])
def test_golden_rosenbrock_circle_regularised(oracle, qds_cls, tol):
    """test/unit-test.jl:78-152 (G3): sigma, rho, delta = 0.5, 0.1, 0.25."""
    nlp = models.unit_test_model_rosenbrock_circle()
    f = fpsb200.FletcherPenaltyNLP(nlp, 0.5, 0.1, 0.25, 1, qds=qds_cls(nlp, 0.0))
    x = np.array(G["G3"]["x"])
    assert abs(f.obj(x) - G["G3"]["obj"]) < max(10 * tol, 1e-14)
    assert abs(f.fx - G["G3"]["fx"]) < 1e-14
    assert np.allclose(f.gx, G["G3"]["gx"], atol=1e-13)
    assert np.allclose(f.ys, G["G3"]["ys"], atol=max(tol, 1e-14))
    assert np.allclose(f.cx, G["G3"]["cx"], atol=1e-14)
    fo, go = f.objgrad(x)
    assert abs(fo - f.obj(x)) < 1e-12 and np.allclose(go, f.grad(x))


def _problem(m=40, n=90, seed=3):
    A = models.window_random_jacobian(m, n, 7, w=10, seed=seed)
    rng = np.random.default_rng(seed)
    return A, rng.standard_normal(n), rng.standard_normal(m), rng.standard_normal(n)


@pytest.mark.parametrize("delta", [0.0, 0.25, 1e-4])
def test_oracle_vs_dense_solve(oracle, delta):
    A, r1, r2, r3 = _problem()
    m, n = A.shape
    Ad = A.toarray()
    K = np.block([[np.eye(n), Ad.T], [Ad, -delta * np.eye(m)]])
    s1 = np.linalg.solve(K, np.r_[r1, np.zeros(m)])
    s2 = np.linalg.solve(K, np.r_[np.zeros(n), r2])
    s3 = np.linalg.solve(K, np.r_[r3, np.zeros(m)])
    coo = A.tocoo()
    for P in (np.arange(n + m), np.random.default_rng(0).permutation(n + m)):
        lo = oracle.LDLtOracle(n, m, coo.row, coo.col, P)
        p1, q1, p2, q2, ok = lo.solve_two_mixed(coo.data, delta, r1, r2)
        assert ok
        # delta = 0: a constraint pivoted before all of its variables has D = -0 and is replaced by
        # -sqrt(eps) (dynamic regularisation) => O(sqrt(eps)) perturbation of the solution
        _, D = lo.numeric()
        bar = 1e-6 if np.any(np.abs(np.abs(D) - SQRT_EPS) < 1e-12) else 1e-11
        assert _rel(np.r_[p1, q1], s1) < bar and _rel(np.r_[p2, q2], s2) < bar
        p1, q1, p2, q2, ok = lo.solve_two_least_squares(r1, r3)
        assert _rel(np.r_[p1, q1], s1) < bar and _rel(np.r_[p2, q2], s3) < bar
    it = oracle.IterativeOracle(A)
    p1, q1, p2, q2, st = it.solve_two_mixed(delta, r1, r2)
    assert st[0]["solved"] and st[1]["solved"]
    assert _rel(np.r_[p1, q1], s1) < 1e-5 and _rel(np.r_[p2, q2], s2) < 1e-5
    p1, q1, p2, q2, st = it.solve_two_least_squares(delta, r1, r3)
    assert _rel(np.r_[p1, q1], s1) < 1e-5 and _rel(np.r_[p2, q2], s3) < 1e-5
    tau = max(delta, 1e-14)
    S = Ad @ Ad.T + tau * np.eye(m)
    u1r, u2r = np.linalg.solve(S, Ad @ r1), np.linalg.solve(S, r2)
    for u1, u2, st in (it.solve_two_extras(delta, r1, r2), lo.solve_two_extras(delta, r1, r2)):
        assert _rel(u1, u1r) < 1e-5 and _rel(u2, u2r) < 1e-4


def test_krylov_restatements_vs_scipy(oracle):
    A, r1, r2, _ = _problem(60, 150, 5)
    x, st = oracle.lsqr(A, r2, lam=0.0, atol=1e-12, rtol=1e-12)
    xs = spla.lsqr(A, r2, atol=1e-14, btol=1e-14)[0]
    assert _rel(A @ x, A @ xs) < 1e-6
    S = (A @ A.T + 0.1 * sp.identity(60)).tocsr()
    x, st = oracle.minres_normal(A, r2, 0.1, atol=1e-12, rtol=1e-12)
    assert st["solved"] and _rel(S @ x, r2) < 1e-8
    x, y, st = oracle.craig(A, r2, 0.0, atol=1e-12, rtol=1e-12, btol=1e-12)
    assert st["solved"] and _rel(A @ x, r2) < 1e-9 and _rel(x, A.T @ y) < 1e-9
    x, st = oracle.cgls(A, r1, lam=0.3, transpose=True, atol=1e-12, rtol=1e-12)
    xd = np.linalg.solve(A.toarray() @ A.toarray().T + 0.3 * np.eye(60), A.toarray() @ r1)
    assert _rel(x, xd) < 1e-8


def test_dynamic_regularisation_rule(oracle):
    """|D[k]| < tol -> D[k] = sign(r) max(|D[k] + r|, |r|), r1 for the first n_d original indices,
    r2 otherwise; D == 0 without regularisation -> not factorised."""
    A = sp.csr_matrix(np.array([[1.0, 0.0, 0.0], [1.0, 0.0, 0.0]]))
    coo = A.tocoo()
    lo = oracle.LDLtOracle(3, 2, coo.row, coo.col, np.arange(5))
    out = lo.solve_two_mixed(coo.data, 0.0, np.ones(3), np.ones(2))
    assert out[4]
    _, D = lo.numeric()
    assert np.allclose(D[:3], 1.0) and D[3] == -1.0 and D[4] == -SQRT_EPS
    lo = oracle.LDLtOracle(3, 2, coo.row, coo.col, np.arange(5), ldlt_tol=0.0, ldlt_r1=0.0, ldlt_r2=0.0)
    r1, r2 = np.array([1.0, 2.0, 3.0]), np.array([4.0, 5.0])
    p1, q1, p2, q2, ok = lo.solve_two_mixed(coo.data, 0.0, r1, r2)
    assert not ok and np.array_equal(p1, r1) and np.array_equal(q2, r2) and not q1.any() and not p2.any()


def test_sparse_coo_to_csc_semantics(oracle):
    """sparse(I, J, V): duplicates summed, explicit zeros kept, rows ascending (App. B4)."""
    I = [2, 0, 2, 1, 0]; J = [1, 0, 1, 1, 0]; V = [1.0, 0.0, 2.5, 0.0, 0.0]
    Cp, Ci, Cx, slot = oracle.coo_to_csc(3, I, J, V)
    assert Cp.tolist() == [0, 1, 3, 3] and Ci.tolist() == [0, 1, 2] and Cx.tolist() == [0.0, 0.0, 3.5]
