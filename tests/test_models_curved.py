"""CPU: the curved-constraint test model (models.CurvedQPModel) that the device-resident Val(1) Hessian product is checked
on (tests/test_gpu_device_nlp.py): its Jacobian, hprod and ghjvprod against finite differences."""
import numpy as np


def _curved(n=400, m=150, seed=7):
    from fpsb200 import models
    qp = models.sparse_qp(n, m, nnz_per_row=6, w=24, seed=seed)
    rng = np.random.default_rng(seed + 1)
    return models.CurvedQPModel(qp.Q, qp.q, qp.A, qp.b, 0.3 * rng.standard_normal(m))


def test_curved_model_derivatives_are_consistent():
    """The curved-constraint test model itself (host): Jacobian, hprod and ghjvprod against finite differences."""
    cm = _curved(60, 20, 3)
    rng = np.random.default_rng(0)
    x, v, g, y = rng.standard_normal(60), rng.standard_normal(60), rng.standard_normal(60), rng.standard_normal(20)
    t = 1e-6
    assert np.linalg.norm((cm.cons(x + t * v) - cm.cons(x - t * v)) / (2 * t) - cm.jprod(x, v)) < 1e-8
    r, c = cm.jac_structure()
    J = np.zeros((20, 60)); np.add.at(J, (r, c), cm.jac_coord(x))
    assert np.linalg.norm(J @ v - cm.jprod(x, v)) < 1e-12 and np.linalg.norm(J.T @ y - cm.jtprod(x, y)) < 1e-12
    lag = lambda z: cm.grad(z) + cm.jtprod(z, y)
    assert np.linalg.norm((lag(x + t * v) - lag(x - t * v)) / (2 * t) - cm.hprod(x, y, v)) < 1e-7
    jg = lambda z: cm.jprod(z, g)
    assert np.linalg.norm((jg(x + t * v) - jg(x - t * v)) / (2 * t) - cm.ghjvprod(x, g, v)) < 1e-8
