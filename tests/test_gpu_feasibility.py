"""GPU: the feasibility restoration's normal step on the fused kernels: `TR_fused` (MINRES slot of
solve_two_extras + one product with J') against the host-loop `TR_lsmr` of src/feasibility.jl:208-235, and a whole
`feasibility_step` (src/feasibility.jl:21-189) driven by it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _curved(n=400, m=150, seed=7):
    from fpsb200 import models
    qp = models.sparse_qp(n, m, nnz_per_row=6, w=24, seed=seed)
    rng = np.random.default_rng(seed + 1)
    return models.CurvedQPModel(qp.Q, qp.q, qp.A, qp.b, 0.3 * rng.standard_normal(m))


def _qds(kind, nlp):
    import fpsb200
    if kind == "ldlt":
        return fpsb200.LDLtSolver(nlp, 0.0)
    return fpsb200.IterativeSolver(nlp, 0.0, ne_atol=1e-12, ne_rtol=1e-12)


@pytest.mark.parametrize("kind", ["ldlt", "iterative"])
def test_tr_fused_matches_tr_lsmr(kind):
    from fpsb200.fps_solve import TR_fused, TR_lsmr
    cm = _curved()
    qds = _qds(kind, cm)
    rng = np.random.default_rng(3)
    z = 0.5 * rng.standard_normal(400)
    cz = cm.cons(z)
    ncz = float(np.linalg.norm(cz))
    # radius inactive: both are the minimum-norm solution of J d = -c (J has full row rank)
    d0, Jd0, inf0, ok0 = TR_lsmr(cm, z, cz, 1e-8, 1e6, ncz)
    d1, Jd1, inf1, ok1 = TR_fused(qds, cm, z, cz, 1e-8, 1e6, ncz)
    assert ok1 and not inf1
    assert np.linalg.norm(Jd1 + cz) <= 1e-6 * ncz
    assert np.linalg.norm(d1 - d0) <= 1e-5 * np.linalg.norm(d0)
    assert np.linalg.norm(Jd1 - cm.jprod(z, d1)) <= 1e-12 * np.linalg.norm(Jd1)
    # radius active: on the boundary, along the same direction, the linearised residual decreases
    Delta = 0.25 * float(np.linalg.norm(d1))
    d2, Jd2, inf2, _ = TR_fused(qds, cm, z, cz, 1e-8, Delta, ncz)
    assert abs(np.linalg.norm(d2) - Delta) <= 1e-12 * Delta and not inf2
    assert np.linalg.norm(Jd2 + cz) < ncz
    assert abs(d2 @ d1 / (np.linalg.norm(d2) * np.linalg.norm(d1)) - 1.0) < 1e-12


def test_tr_fused_device_tensors():
    import torch
    import fpsb200
    from fpsb200.fps_solve import TR_fused
    cm = _curved()
    dcm = fpsb200.DeviceCurvedQP(cm)
    qds = _qds("iterative", dcm)
    rng = np.random.default_rng(5)
    z = 0.5 * rng.standard_normal(400)
    zd = torch.tensor(z, device="cuda")
    czd = dcm.cons(zd)
    d, Jd, infeasible, ok = TR_fused(qds, dcm, zd, czd, 1e-8, 1e6, float(czd.norm()))
    assert d.is_cuda and Jd.is_cuda and ok and not infeasible
    href = TR_fused(_qds("iterative", cm), cm, z, cm.cons(z), 1e-8, 1e6, float(czd.norm()))
    assert np.linalg.norm(d.cpu().numpy() - href[0]) <= 1e-9 * np.linalg.norm(href[0])


@pytest.mark.parametrize("kind", ["ldlt", "iterative"])
def test_feasibility_step_on_fused_kernels(kind):
    from fpsb200.fps_solve import GNSolver, feasibility_step
    cm = _curved()
    rng = np.random.default_rng(8)
    x = rng.standard_normal(400)
    cx = cm.cons(x)
    for fused in (False, True):
        z, cz, ncz, status = feasibility_step(GNSolver(fused=fused), cm, x, cx, float(np.linalg.norm(cx)), 1e-7, 1e-7,
                                              qds=_qds(kind, cm) if fused else None)
        assert status == "success" and ncz <= 1e-7
        assert np.linalg.norm(cm.cons(z) - cz) <= 1e-12 * max(1.0, np.linalg.norm(cz))
