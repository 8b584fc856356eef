"""GPU: throughput mode (BASELINE config C5) — batches of independent small instances through
fpsb_batch_solve_two, checked instance by instance against the oracle's LDLt path (natural
ordering) and a dense solve; includes rank-deficient and failing instances."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SQRT_EPS = float(np.sqrt(np.finfo(float).eps))


@pytest.mark.parametrize("n,m,delta", [(2, 1, 0.0), (3, 2, 0.25), (10, 3, 1e-2), (10, 1, 0.0), (20, 12, 1e-3)])
def test_batch_mixed_and_least_squares_vs_oracle(oracle, n, m, delta):
    import fpsb200
    rng = np.random.default_rng(n * 100 + m)
    ninst = 257
    A = rng.standard_normal((ninst, m, n))
    r1 = rng.standard_normal((ninst, n)); r2 = rng.standard_normal((ninst, m)); r3 = rng.standard_normal((ninst, n))
    p1, q1, p2, q2, ok = fpsb200.batch_solve_two(A, delta, r1, r2, "mixed")
    P1, Q1, P2, Q2, ok2 = fpsb200.batch_solve_two(A, delta, r1, r3, "least_squares")
    assert ok.all() and ok2.all()
    rows, cols = np.meshgrid(np.arange(m), np.arange(n), indexing="ij")
    for i in range(0, ninst, 16):
        lo = oracle.LDLtOracle(n, m, rows.ravel(), cols.ravel(), np.arange(n + m))
        o = lo.solve_two_mixed(A[i].ravel(), delta, r1[i], r2[i])
        for a, b in ((p1[i], o[0]), (q1[i], o[1]), (p2[i], o[2]), (q2[i], o[3])):
            assert np.allclose(a, b, rtol=1e-10, atol=1e-12)
        o = lo.solve_two_least_squares(r1[i], r3[i])
        for a, b in ((P1[i], o[0]), (Q1[i], o[1]), (P2[i], o[2]), (Q2[i], o[3])):
            assert np.allclose(a, b, rtol=1e-10, atol=1e-12)
        K = np.block([[np.eye(n), A[i].T], [A[i], -delta * np.eye(m)]])
        s = np.linalg.solve(K, np.r_[r1[i], np.zeros(m)])
        assert np.allclose(np.r_[p1[i], q1[i]], s, rtol=1e-9, atol=1e-11)


def test_batch_rank_deficient_and_failed_instances(oracle):
    """HS61-like rank deficiency at x0 = 0 (test/rank-deficient.jl:23-29): J(0) = [3 0 0; 4 0 0]."""
    import fpsb200
    J = np.array([[3.0, 0.0, 0.0], [4.0, 0.0, 0.0]])
    A = np.stack([J, np.array([[1.0, 2.0, 3.0], [0.0, 1.0, 1.0]])])
    r1 = np.array([[-33.0, 16.0, -24.0]] * 2); r2 = np.array([[-7.0, -11.0]] * 2)
    p1, q1, p2, q2, ok = fpsb200.batch_solve_two(A, 0.0, r1, r2)
    assert ok.all()
    rows, cols = np.meshgrid(np.arange(2), np.arange(3), indexing="ij")
    for i in range(2):
        lo = oracle.LDLtOracle(3, 2, rows.ravel(), cols.ravel(), np.arange(5))
        o = lo.solve_two_mixed(A[i].ravel(), 0.0, r1[i], r2[i])
        assert np.allclose(q1[i], o[1], rtol=1e-9) and np.allclose(p2[i], o[2], rtol=1e-9)
    # regularisation off: the rank-deficient instance fails and returns its right-hand sides
    opts = fpsb200.LdltOpts(); opts.ldlt_tol = 0.0; opts.ldlt_r1 = 0.0; opts.ldlt_r2 = 0.0
    A2 = np.stack([np.array([[1.0, 0.0, 0.0], [1.0, 0.0, 0.0]]), A[1]])
    p1, q1, p2, q2, ok = fpsb200.batch_solve_two(A2, 0.0, r1, r2, opts=opts)
    assert not ok[0] and ok[1]
    assert np.array_equal(p1[0], r1[0]) and np.array_equal(q2[0], r2[0]) and not q1[0].any() and not p2[0].any()


def test_batch_large_throughput_shape():
    """4096 instances of the largest C5 shape run in one launch and solve their systems."""
    import fpsb200
    rng = np.random.default_rng(0)
    A = rng.standard_normal((4096, 3, 10)); r1 = rng.standard_normal((4096, 10)); r2 = rng.standard_normal((4096, 3))
    p1, q1, p2, q2, ok = fpsb200.batch_solve_two(A, 1e-3, r1, r2)
    assert ok.all()
    res = p1 + np.einsum("imn,im->in", A, q1) - r1
    assert np.abs(res).max() < 1e-10
