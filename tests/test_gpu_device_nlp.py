"""GPU: device-resident FletcherPenaltyNLP (fpsb_fp_* fused combination kernels + FPSB_DEVICE solves,
SURVEY §8 f1) against the host mirror of src/model-Fletcherpenaltynlp.jl on the same model."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("solver", ["ldlt", "iterative"])
@pytest.mark.parametrize("sigma,rho,delta", [(1.0, 0.0, 0.0), (10.0, 2.0, 1e-2)])
def test_device_nlp_matches_host_mirror(solver, sigma, rho, delta):
    import torch
    import fpsb200
    from fpsb200 import models
    qp = models.sparse_qp(400, 150, nnz_per_row=6, w=24, seed=7)
    dqp = fpsb200.DeviceSparseQP(qp)
    mk = (lambda nlp: fpsb200.LDLtSolver(nlp, 0.0)) if solver == "ldlt" else (lambda nlp: fpsb200.IterativeSolver(nlp, 0.0, ls_atol=1e-13, ls_rtol=1e-13, ln_atol=1e-13, ln_rtol=1e-13, ln_btol=1e-13))
    host = fpsb200.FletcherPenaltyNLP(qp, sigma, rho, delta, 2, qds=mk(qp))
    dev = fpsb200.DeviceFletcherPenaltyNLP(dqp, sigma, rho, delta, qds=mk(dqp))
    rng = np.random.default_rng(1)
    tol = 1e-9 if solver == "ldlt" else 1e-6
    rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    for _ in range(2):
        x = rng.standard_normal(400); v = rng.standard_normal(400)
        xd = torch.tensor(x, device="cuda"); vd = torch.tensor(v, device="cuda")
        assert abs(dev.obj(xd) - host.obj(x)) <= tol * max(1.0, abs(host.obj(x)))
        gd = dev.grad(xd)
        assert gd.is_cuda and rel(gd.cpu().numpy(), host.grad(x)) < tol
        assert rel(dev.ys.cpu().numpy(), host.ys) < tol and rel(dev.gs.cpu().numpy(), host.gs) < tol
        Hd = dev.hprod(xd, vd, obj_weight=0.5)
        assert Hd.is_cuda and rel(Hd.cpu().numpy(), host.hprod(x, v, obj_weight=0.5)) < tol
    # memo: same x -> no new solve ; different x -> new key
    k1 = dev._hash(xd); k2 = dev._hash(xd.clone()); k3 = dev._hash(xd + 1e-16 * 0 + torch.roll(xd, 1) * 0 + 0)
    assert k1 == k2 == k3
    assert dev._hash(torch.roll(xd, 1)) != k1        # a permutation of the same values changes the key
    n_before = dev.handle.launch_count()
    dev.obj(xd)
    assert dev.handle.launch_count() - n_before <= 2  # hash + obj reduction only, no solve


def test_fp_obj_proximal_term_and_edge_cases():
    import ctypes as C
    import torch
    import fpsb200
    from fpsb200 import _lib, models
    qp = models.sparse_qp(50, 20, nnz_per_row=4, w=8, seed=3)
    H = fpsb200.IterativeSolver(qp, 0.0).handle
    L = _lib.lib()
    rng = np.random.default_rng(0)
    c = rng.standard_normal(20); ys = rng.standard_normal(20); x = rng.standard_normal(50); xk = rng.standard_normal(50)
    t = lambda a: torch.tensor(a, device="cuda")
    cd, yd, xd, kd = t(c), t(ys), t(x), t(xk)
    phi = C.c_double()
    p = lambda z: C.c_void_p(z.data_ptr())
    assert L.fpsb_fp_obj(H.h, C.c_double(3.0), C.c_double(0.7), C.c_double(0.3), p(cd), p(yd), p(xd), p(kd), C.byref(phi)) == 0
    ref = 3.0 - c @ ys + 0.35 * (c @ c) + 0.15 * np.linalg.norm(x - xk) ** 2
    assert abs(phi.value - ref) < 1e-12 * max(1, abs(ref))
    assert L.fpsb_fp_obj(H.h, C.c_double(3.0), C.c_double(0.0), C.c_double(0.0), p(cd), p(yd), None, None, C.byref(phi)) == 0
    assert abs(phi.value - (3.0 - c @ ys)) < 1e-12
    assert L.fpsb_fp_grad(H.h, C.c_double(1.0), C.c_double(1.0), C.c_double(0.0), p(xd), p(xd), p(xd), p(xd), None, None, None, p(xd)) != 0   # rho > 0 without J'c


def _curved(n=400, m=150, seed=7):
    from fpsb200 import models
    qp = models.sparse_qp(n, m, nnz_per_row=6, w=24, seed=seed)
    rng = np.random.default_rng(seed + 1)
    return models.CurvedQPModel(qp.Q, qp.q, qp.A, qp.b, 0.3 * rng.standard_normal(m))


@pytest.mark.parametrize("solver", ["ldlt", "iterative"])
@pytest.mark.parametrize("sigma,rho,delta", [(1.0, 0.0, 0.0), (10.0, 2.0, 1e-2)])
def test_device_val1_hprod_matches_host_mirror(solver, sigma, rho, delta):
    """Val(1) (exact Hessian: ghjvprod + solve_two_extras, src/model-Fletcherpenaltynlp.jl:572-634) device-resident on a
    model with curved constraints, against the host mirror driving the same GPU solvers."""
    import torch
    import fpsb200
    cm = _curved()
    dcm = fpsb200.DeviceCurvedQP(cm)
    tight = dict(ls_atol=1e-13, ls_rtol=1e-13, ln_atol=1e-13, ln_rtol=1e-13, ln_btol=1e-13, ne_atol=1e-13, ne_rtol=1e-13)
    mk = (lambda nlp: fpsb200.LDLtSolver(nlp, 0.0)) if solver == "ldlt" else (lambda nlp: fpsb200.IterativeSolver(nlp, 0.0, **tight))
    host = fpsb200.FletcherPenaltyNLP(cm, sigma, rho, delta, 1, qds=mk(cm))
    dev = fpsb200.DeviceFletcherPenaltyNLP(dcm, sigma, rho, delta, 1, qds=mk(dcm))
    rng = np.random.default_rng(4)
    # solve_two_extras is iterative on both paths (LDLt: CGLS + MINRES at the reference's sqrt(eps) tolerances,
    # src/solve_linear_system.jl:45-77), so two runs from inputs that differ in the last bits agree to the solve tolerance
    tol = 1e-7 if solver == "ldlt" else 1e-5
    rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    for _ in range(2):
        x = 0.5 * rng.standard_normal(400); v = rng.standard_normal(400)
        xd = torch.tensor(x, device="cuda"); vd = torch.tensor(v, device="cuda")
        assert abs(dev.obj(xd) - host.obj(x)) <= tol * max(1.0, abs(host.obj(x)))
        assert rel(dev.grad(xd).cpu().numpy(), host.grad(x)) < tol
        Hd = dev.hprod(xd, vd, obj_weight=0.5)
        Hh = host.hprod(x, v, obj_weight=0.5)
        assert Hd.is_cuda and rel(Hd.cpu().numpy(), Hh) < tol
        # the Val(1) terms are really there: the Val(2) product of the same model differs
        dev2 = fpsb200.DeviceFletcherPenaltyNLP(dcm, sigma, rho, delta, 2, qds=dev.qdsolver)
        assert rel(dev2.hprod(xd, vd, obj_weight=0.5).cpu().numpy(), Hh) > 1e-4
