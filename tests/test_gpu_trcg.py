"""GPU: the device-resident Steihaug–Toint CG of the trust-region subsolver (fpsb_trcg_init / fpsb_trcg_step, SURVEY §8
f3) against the host loop `_steihaug` on the same quadratic models: interior convergence, boundary exit, negative
curvature, bound-masked variables."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _handle(n):
    import fpsb200
    m = 3
    rows = np.array([0, 1, 2], dtype=np.int64)
    cols = np.array([0, 1, 2], dtype=np.int64)
    return fpsb200.B200Handle(n, m, rows, cols)


@pytest.mark.parametrize("case", ["interior", "boundary", "negative_curvature", "masked", "zero_gradient"])
def test_device_steihaug_matches_host_loop(case):
    import torch
    from fpsb200.fps_solve import _steihaug, _steihaug_device
    n = 5000
    rng = np.random.default_rng(12)
    diag = 1.0 + rng.random(n) * 9.0
    if case == "negative_curvature":
        diag[::7] = -0.5
    U = rng.standard_normal((n, 3)) / np.sqrt(n)
    g = rng.standard_normal(n)
    if case == "zero_gradient":
        g[:] = 0.0
    radius = {"interior": 1e6, "boundary": 0.5, "negative_curvature": 50.0, "masked": 1e6, "zero_gradient": 1.0}[case]
    free = None
    if case == "masked":
        free = (rng.random(n) > 0.3) * 1.0
    hv_h = lambda v: diag * v + U @ (U.T @ v)
    dg, dU = torch.tensor(diag, device="cuda"), torch.tensor(U, device="cuda")
    hv_d = lambda v: dg * v + dU @ (dU.T @ v)
    tol = 1e-8 * np.linalg.norm(g)
    s0, pred0, np0 = _steihaug(hv_h, g, radius, tol, 200, free)
    H = _handle(n)
    fd = None if free is None else torch.tensor(free, device="cuda")
    s1, pred1, np1 = _steihaug_device(H, hv_d, torch.tensor(g, device="cuda"), radius, tol, 200, fd)
    assert s1.is_cuda
    s1 = s1.cpu().numpy()
    assert np1 == np0
    if case == "zero_gradient":
        assert np1 == 0 and not s1.any() and pred1 == 0.0
        return
    assert np.linalg.norm(s1 - s0) <= 1e-10 * np.linalg.norm(s0)
    assert abs(pred1 - pred0) <= 1e-10 * abs(pred0)
    if case in ("boundary", "negative_curvature"):
        assert abs(np.linalg.norm(s1) - radius) <= 1e-10 * radius
    if case == "masked":
        assert not s1[free == 0.0].any()
    # the model decrease reported is the decrease of the quadratic at s
    q = g @ s1 + 0.5 * s1 @ hv_h(s1)
    assert abs(-q - pred1) <= 1e-9 * abs(pred1)


def test_trunk_uses_the_device_cg_and_agrees_with_the_host_loop(monkeypatch):
    """A whole device-resident `trunk` run on the penalty model of a sparse QP: same minimiser with the fused CG and with
    the host-loop CG (FPSB_TRCG_HOST=1)."""
    import torch
    import fpsb200
    from fpsb200 import models
    from fpsb200.fps_solve import trunk
    qp = models.sparse_qp(300, 100, nnz_per_row=5, w=20, seed=2)
    res = []
    for host_cg in ("0", "1"):
        monkeypatch.setenv("FPSB_TRCG_HOST", host_cg)
        dqp = fpsb200.DeviceSparseQP(qp)
        dev = fpsb200.DeviceFletcherPenaltyNLP(dqp, 10.0, 1.0, 0.0, qds=fpsb200.LDLtSolver(dqp, 0.0))
        n0 = dev.handle.launch_count()
        out = trunk(dev, torch.zeros(300, dtype=torch.float64, device="cuda"), atol=1e-8, rtol=1e-8)
        assert out.optimal
        res.append((out.x.cpu().numpy(), out.iter, out.cg_iter))
    assert np.linalg.norm(res[0][0] - res[1][0]) <= 1e-7 * max(1.0, np.linalg.norm(res[1][0]))
    assert abs(res[0][1] - res[1][1]) <= 1 and abs(res[0][2] - res[1][2]) <= 2
