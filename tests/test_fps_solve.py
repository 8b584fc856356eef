"""CPU: the `fps_solve` surface (src/FletcherPenaltySolver.jl:127-207, src/algo.jl, src/parameters.jl,
src/feasibility.jl) on the reference's own solver tests — test/test-2.jl, test/rank-deficient.jl:22-36,
docs/src/fine-tuneFPS.md:25-33 — with the 2-RHS solves done by the CPU oracle (test-only QDSolver subtypes
in tests/oracle_qds.py).  The GPU twin of this file runs the same problems through libfpsb200.so."""
import math
import warnings

import numpy as np
import pytest

import fpsb200
from fpsb200 import models
import importlib
F = importlib.import_module("fpsb200.fps_solve")      # the module (the package also exports the function)

SOLUTIONS = {  # known minimisers quoted by the reference's tests
    "rosenbrock_sum": [-1.612771347383541, 2.612771347383541],     # test/test-2.jl:9
    "simple": [0.1] * 10,                                          # test/test-2.jl:39 (n x = 1)
    "spurious": [1.0],                                             # test/test-2.jl:257
    "hs28": [0.5, -0.5, 0.5],
    "hs6": [1.0, 1.0],
}


@pytest.fixture(scope="module")
def qds(oracle):
    import oracle_qds
    return {"ldlt": oracle_qds.OracleLDLt, "iterative": oracle_qds.OracleIterative}


@pytest.mark.parametrize("solver", ["ldlt", "iterative"])
@pytest.mark.parametrize("ha", [1, 2])
@pytest.mark.parametrize("name", models.REFERENCE_TEST_PROBLEMS)
def test_reference_solver_tests(qds, name, solver, ha):
    nlp = models.reference_test_problem(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        stats = F.fps_solve(nlp, nlp.meta.x0, hessian_approx=ha, qds_solver=qds[solver])
    scale = max(np.linalg.norm(nlp.meta.x0), 1.0)
    assert stats.status == "first_order"                       # the assertions of test/test-2.jl:15-17
    if name not in ("hs61", "hs28", "readme_eq"):              # test/test-2.jl asserts these two, the others only the status
        assert stats.dual_feas < 1e-6 * scale
        assert stats.primal_feas < 1e-6 * scale
    assert stats.iter >= 1 and stats.multipliers.shape == (nlp.meta.ncon,)
    if name in SOLUTIONS:
        assert np.linalg.norm(stats.solution - np.array(SOLUTIONS[name])) < 1e-5
    c = nlp.cons(stats.solution) - nlp.meta.lcon
    assert np.abs(c).max() < 1e-6 * scale


def test_keywords_and_defaults():
    a = F.AlgoData()
    se = math.sqrt(np.finfo(float).eps)
    # src/parameters.jl:69-93
    assert (a.sigma_0, a.sigma_update, a.rho_0, a.rho_update, a.delta_update) == (1e3, 2.0, 1.0, 2.0, 10.0)
    assert a.sigma_max == a.rho_max == a.delta_max == a.lagrange_bound == a.subpb_unbounded_threshold == 1 / se
    assert a.delta_0 == se and a.eta_1 == 0.0 and a.eta_update == 1.0 and a.Delta == 0.95
    assert a.subsolver_max_iter == 20000 and a.hessian_approx == 2 and a.atol_sub(0.3) == 0.3
    b = F.AlgoData(**F._ascii_kwargs({"σ_0": 5.0, "ρ_update": 3.0, "δ_0": 1e-2, "Δ": 0.5}))
    assert (b.sigma_0, b.rho_update, b.delta_0, b.Delta) == (5.0, 3.0, 1e-2, 0.5)
    g = F.GNSolver()
    assert (g.eta1, g.eta2, g.sigma1, g.sigma2, g.Delta0, g.bad_steps_lim) == (1e-3, 0.66, 0.25, 2.0, 1.0, 3)
    assert set(fpsb200.qdsolver_correspondence) == {"iterative", "ldlt"}       # src/parameters.jl:197


def test_maximisation_and_inequalities_are_rejected(qds):
    nlp = models.reference_test_problem("hs6")
    nlp.meta.minimize = False
    with pytest.raises(ValueError):
        F.fps_solve(nlp, qds_solver=qds["ldlt"])                                # src/FletcherPenaltySolver.jl:135
    nlp = models.reference_test_problem("readme_ineq")
    s = F.FPSSSolver(nlp, qds_solver=qds["ldlt"])
    with pytest.raises(ValueError):
        F.solve(s)                                                              # src/algo.jl:40-43


def test_slack_model():
    nlp = models.reference_test_problem("readme_ineq")
    s = models.SlackModel(nlp)
    assert s.meta.nvar == 3 and s.meta.ncon == 1
    assert s.meta.lvar[2] == 0.0 and s.meta.uvar[2] == 1.0 and s.meta.lcon[0] == s.meta.ucon[0] == 0.0
    x = np.array([0.3, -0.7, 0.2])
    assert np.allclose(s.cons(x), nlp.cons(x[:2]) - 0.2)
    r, c = s.jac_structure()
    J = np.zeros((1, 3)); J[r, c] = s.jac_coord(x)
    assert np.allclose(J, [[x[1], x[0], -1.0]])
    assert np.allclose(s.grad(x), np.append(nlp.grad(x[:2]), 0.0))


def test_callback_and_user_stop(qds):
    nlp = models.reference_test_problem("hs7")
    seen = []

    def cb(model, solver, stats):
        seen.append(stats.iter)
        if stats.iter >= 1:
            stats.status = "user"                                               # src/algo.jl:111

    stats = F.fps_solve(nlp, qds_solver=qds["ldlt"], callback=cb)
    assert stats.status == "user" and stats.iter == 1 and seen[0] == -1


def test_restart_with_another_initial_guess(qds):
    """test/restart.jl:1-27 with an equality problem this subsolver handles."""
    nlp = models.reference_test_problem("hs6")
    solver = F.FPSSSolver(nlp, qds_solver=qds["ldlt"])
    s1 = F.solve(solver)
    assert s1.status == "first_order" and np.allclose(s1.solution, [1.0, 1.0], atol=1e-6)
    nlp.meta.x0[:] = [3.0, -2.0]
    solver.reset()
    s2 = F.solve(solver)
    assert s2.status == "first_order" and np.allclose(s2.solution, [1.0, 1.0], atol=1e-6)


def test_penalty_updates():
    class M:
        sigma, rho, delta, shahx = 1e3, 1.0, 0.0, 7
    meta = F.AlgoData()
    m = M()
    F._update_parameters(meta, m, feas=True)                    # src/algo.jl:366-381: rho only when infeasible
    assert (m.sigma, m.rho, m.shahx) == (2e3, 1.0, 7)           # the memo survives the update (reference, App. D-1)
    m2 = M()
    F._update_parameters(F.AlgoData(refresh_memo_on_update=True), m2, feas=True)
    assert (m2.sigma, m2.shahx) == (2e3, 0)                     # opt-in: drop it
    F._update_parameters(meta, m, feas=False)
    assert (m.sigma, m.rho) == (4e3, 2.0)
    F._update_parameters_unbdd(meta, m, feas=False)             # :388-395: delta starts at delta_0, then x10
    assert m.delta == meta.delta_0
    F._update_parameters_unbdd(meta, m, feas=False)
    assert m.delta == 10 * meta.delta_0


def test_lsmr_with_radius_and_feasibility_step():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((6, 9)); b = rng.standard_normal(6)
    x, ok = F._lsmr_radius(lambda v: A @ v, lambda u: A.T @ u, b, 9, 0.0, atol=1e-12, btol=1e-12)
    assert ok and np.linalg.norm(x - np.linalg.pinv(A) @ b) < 1e-8
    xr, ok = F._lsmr_radius(lambda v: A @ v, lambda u: A.T @ u, b, 9, 0.1)
    assert ok and abs(np.linalg.norm(xr) - 0.1) < 1e-12            # stopped on the trust-region boundary
    # feasibility_step: from an infeasible point of HS7 to ||c|| <= 1e-8   (src/feasibility.jl:21-189)
    nlp = models.reference_test_problem("hs7")
    x0 = np.array([2.0, 2.0]); c0 = nlp.cons(x0)
    z, cz, ncz, status = F.feasibility_step(F.GNSolver(), nlp, x0, c0, np.linalg.norm(c0), 1e-8, 1e-8)
    assert status == "success" and ncz <= 1e-8 and np.linalg.norm(nlp.cons(z)) <= 1e-8


def test_trunk_on_an_unconstrained_function_and_with_bounds():
    class Rosen:
        def obj(self, x): return (x[0] - 1) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2
        def grad(self, x): return np.array([2 * (x[0] - 1) - 400 * x[0] * (x[1] - x[0] ** 2), 200 * (x[1] - x[0] ** 2)])
        def objgrad(self, x): return self.obj(x), self.grad(x)
        def hprod(self, x, v):
            return np.array([[2 - 400 * x[1] + 1200 * x[0] ** 2, -400 * x[0]], [-400 * x[0], 200.0]]) @ v
    out = F.trunk(Rosen(), np.array([-1.2, 1.0]), atol=1e-9, rtol=1e-9)
    assert out.optimal and np.linalg.norm(out.x - 1.0) < 1e-6
    out = F.trunk(Rosen(), np.array([-1.2, 1.0]), atol=1e-9, rtol=1e-9, lvar=np.array([-2.0, -2.0]), uvar=np.array([0.5, 2.0]))
    assert out.optimal and abs(out.x[0] - 0.5) < 1e-8 and abs(out.x[1] - 0.25) < 1e-6
    out = F.trunk(Rosen(), np.array([-1.2, 1.0]), max_iter=2)
    assert out.iteration_limit and not out.optimal


def test_consistent_gradient_is_the_derivative_of_obj(qds):
    """The reference's grad! passes +ys to hprod! (src/model-Fletcherpenaltynlp.jl:382): kept by default,
    `consistent_gradient=True` gives the exact derivative of obj (checked by central differences)."""
    nlp = models.reference_test_problem("hs26")
    x = nlp.meta.x0 + 0.1
    h = 1e-6
    for flag, expect_exact in ((True, True), (False, False)):
        fp = fpsb200.FletcherPenaltyNLP(nlp, 1e3, 1.0, 0.0, 2, qds=qds["ldlt"](nlp), consistent_gradient=flag)
        gfd = np.array([(fp.obj(x + h * e) - fp.obj(x - h * e)) / (2 * h) for e in np.eye(3)])
        err = np.abs(fp.grad(x) - gfd).max() / np.abs(gfd).max()
        assert (err < 1e-7) == expect_exact


def test_inequalities_go_through_the_slack_model(qds):
    """README.md:59-65 (0 <= x1 x2 - 1 <= 1): SlackModel + bounded subproblem; the statistics come back in the
    original variables (src/FletcherPenaltySolver.jl:138-184).  With a slack at its bound the penalty function is
    not exact, so feasibility only improves as sigma grows — the run ends on sigma_max (:186-187), feasible to 1e-5."""
    nlp = models.reference_test_problem("readme_ineq")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        stats = F.fps_solve(nlp, qds_solver=qds["ldlt"], consistent_gradient=True)
    assert stats.solution.shape == (2,)
    c = nlp.cons(stats.solution)[0]
    assert -1e-5 <= c <= 1.0 + 1e-5
    assert stats.status in ("first_order", "unknown") and stats.solver_specific["sigma"] > 1e3
    assert abs(stats.objective - 360.3798) < 1e-2


def test_unconstrained_problem(qds):
    """test/solvertest.jl:1-8 (`unconstrained_nlp`): ncon = 0, the penalty function is f itself."""
    A = np.array
    f = lambda x: (x[0] - 1.0) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2
    g = lambda x: A([2 * (x[0] - 1) - 400 * x[0] * (x[1] - x[0] ** 2), 200 * (x[1] - x[0] ** 2)])
    H = lambda x: A([[2 - 400 * x[1] + 1200 * x[0] ** 2, -400 * x[0]], [-400 * x[0], 200.0]])
    for key in ("ldlt", "iterative"):
        nlp = models.CallableModel(f, g, lambda x: np.zeros(0), lambda x: np.zeros((0, 2)), H, lambda x, j: np.zeros((2, 2)),
                                   A([-1.2, 1.0]), 0, name="rosenbrock")
        stats = F.fps_solve(nlp, qds_solver=qds[key])
        assert stats.status == "first_order" and np.linalg.norm(stats.solution - 1.0) < 1e-3
        assert stats.multipliers.shape == (0,) and stats.primal_feas == 0.0


def test_stopping_statuses(qds):
    """`status_stopping_to_stats` outcomes other than :first_order (src/algo.jl:269): iteration limit of the outer
    loop, time limit, sigma_max exceeded -> fail_sub_pb -> :unknown (:186-187)."""
    nlp = models.reference_test_problem("hs61")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = F.fps_solve(nlp, qds_solver=qds["ldlt"], atol=1e-14, rtol=1e-14, max_iter=0)
        assert s.status == "max_iter" and s.iter == 1
        s = F.fps_solve(models.reference_test_problem("hs61"), qds_solver=qds["ldlt"], atol=1e-14, rtol=1e-14, max_time=0.0)
        assert s.status == "max_time"
        # a penalty parameter that may not grow: the first failed feasibility check ends the run as :unknown
        s = F.fps_solve(models.reference_test_problem("readme_ineq"), qds_solver=qds["ldlt"], consistent_gradient=True,
                        **{"σ_max": 1.5e3})
        assert s.status == "unknown" and s.solver_specific["sigma"] > 1.5e3


def test_stats_fields_and_multipliers_sign(qds):
    """stats.multipliers = -ys (src/algo.jl:137): with L = f + λ'c the KKT residual g + J'λ vanishes at the solution."""
    nlp = models.reference_test_problem("hs7")
    s = F.fps_solve(nlp, qds_solver=qds["iterative"])
    x, lam = s.solution, s.multipliers
    J = nlp.jac_coord(x).reshape(1, 2)
    assert np.linalg.norm(nlp.grad(x) + J.T @ lam) < 1e-5
    assert s.elapsed_time >= 0 and s.objective == pytest.approx(-np.sqrt(3), abs=1e-6)
    assert set(s.solver_specific) >= {"sigma", "rho", "delta", "restoration", "feasibility", "solver"}


def _mgh01feas(rhs2):
    """test/test-2.jl:290-322: f = 0, one linear row -x1 = -1 and c(x) = 10 (x2 - x1^2) = rhs2 (rows: linear first)."""
    A = np.array
    return models.CallableModel(lambda x: 0.0, lambda x: np.zeros(2),
                                lambda x: A([-x[0], 10 * (x[1] - x[0] ** 2)]),
                                lambda x: A([[-1.0, 0.0], [-20 * x[0], 10.0]]),
                                lambda x: np.zeros((2, 2)),
                                lambda x, j: np.zeros((2, 2)) if j == 0 else A([[-20.0, 0.0], [0.0, 0.0]]),
                                [-1.2, 1.0], 2, lcon=[-1.0, rhs2], ucon=[-1.0, rhs2], lin=[0], name="mgh01feas")


@pytest.mark.parametrize("solver", ["ldlt", "iterative"])
@pytest.mark.parametrize("rhs2,x2", [(0.0, 1.0), (1.0, 1.1)])
def test_explicit_linear_constraints(qds, solver, rhs2, x2):
    """"Problems with explicit linear constraints" (test/test-2.jl:290-322): the linear row stays a constraint of the
    subproblem (kept by null-space projections through a second QDSolver), only the nonlinear row is penalised."""
    nlp = _mgh01feas(rhs2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        stats = F.fps_solve(nlp, explicit_linear_constraints=True, qds_solver=qds[solver])
    assert stats.status == "first_order"
    assert np.linalg.norm(nlp.cons(stats.solution) - nlp.meta.lcon) <= 1e-10          # the reference's four assertions
    assert stats.dual_feas <= 1e-10 and stats.primal_feas <= 1e-10
    assert np.linalg.norm(stats.solution - np.array([1.0, x2])) <= 1e-9
    assert stats.multipliers.shape == (2,)


def test_explicit_linear_constraints_model_and_projector(qds):
    """The pieces: the penalty model only sees the nonlinear rows (ys, cx of size nnln, hprod with zero multipliers on
    the linear rows), the projector returns (I - A'(AA')^-1 A) v, (AA')^-1 A v and the minimum-norm correction."""
    import oracle_qds
    nlp = models.CallableModel(lambda x: float(x @ x), lambda x: 2 * x,
                               lambda x: np.array([x[0] + 2 * x[1] - x[2], x[0] * x[1] - 1.0, x[1] - x[3]]),
                               lambda x: np.array([[1.0, 2.0, -1.0, 0.0], [x[1], x[0], 0.0, 0.0], [0.0, 1.0, 0.0, -1.0]]),
                               lambda x: 2 * np.eye(4),
                               lambda x, j: np.array([[0, 1.0, 0, 0], [1.0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]]) if j == 1 else np.zeros((4, 4)),
                               [1.0, 2.0, 0.5, -1.0], 3, lcon=[0.5, 0.0, -0.25], ucon=[0.5, 0.0, -0.25], lin=[0, 2], name="mixed-rows")
    assert nlp.meta.nlin == 2 and nlp.meta.nln == [1]
    x = np.array([0.3, -0.7, 1.1, 0.2])
    r, c = nlp.jac_lin_structure()
    Al = np.zeros((2, 4)); np.add.at(Al, (r, c), nlp.jac_lin_coord(x))
    assert np.allclose(Al, [[1, 2, -1, 0], [0, 1, 0, -1]]) and np.allclose(nlp.cons_lin(x), Al @ x)
    model = fpsb200.FletcherPenaltyNLP(nlp, 2.0, 1.0, 0.0, 2, qds=oracle_qds.OracleLDLt(nlp, 0.0, explicit_linear_constraints=True),
                                       explicit_linear_constraints=True)
    assert model.npen == 1
    model.obj(x)
    assert model.cx.shape == (1,) and abs(model.cx[0] - (x[0] * x[1] - 1.0)) < 1e-15 and model.ys.shape == (1,)
    v = np.array([1.0, -2.0, 0.5, 3.0])
    assert np.allclose(model._hprod_nln(x, np.array([3.0]), v, obj_weight=0.0), 3.0 * np.array([v[1], v[0], 0, 0]))
    assert np.all(np.isfinite(model.grad(x))) and np.all(np.isfinite(model.hprod(x, v)))
    P = F.LinearConstraintProjector(nlp, lambda rows: oracle_qds.OracleLDLt(rows, 0.0))
    pv, lam = P.project(v)
    G = Al @ Al.T
    assert np.allclose(pv, v - Al.T @ np.linalg.solve(G, Al @ v), atol=1e-10) and np.allclose(lam, np.linalg.solve(G, Al @ v), atol=1e-10)
    d = P.correct(P.residual(x))
    assert np.allclose(Al @ (x - d), [0.5, -0.25], atol=1e-10) and np.allclose(d, Al.T @ np.linalg.solve(G, Al @ x - [0.5, -0.25]), atol=1e-10)
    # a whole solve with objective, curved and linear rows.  consistent_gradient: with the reference's gradient formula
    # (+ys where the derivative of obj has -ys) the projected subproblem stalls away from c = 0 on this problem
    kw = dict(qds_solver=oracle_qds.OracleLDLt, consistent_gradient=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nlp.counters.__init__()
        stats = F.fps_solve(nlp, explicit_linear_constraints=True, **kw)
        nlp.counters.__init__()
        ref = F.fps_solve(nlp, **kw)
    assert stats.status == ref.status == "first_order"
    assert np.linalg.norm(stats.solution - ref.solution) <= 1e-5
    assert np.abs(Al @ stats.solution - [0.5, -0.25]).max() <= 1e-10                  # the linear rows hold to rounding
    assert np.allclose(stats.multipliers[[0, 2]], ref.multipliers[[0, 2]], rtol=1e-4)  # same sign convention in both modes


def test_explicit_linear_constraints_rejects_what_it_cannot_do(qds):
    nlp = _mgh01feas(0.0)
    nlp.meta.ucon = np.array([0.0, 0.0])                     # a linear inequality
    with pytest.raises(NotImplementedError):
        F.FPSSSolver(nlp, explicit_linear_constraints=True, qds_solver=qds["ldlt"])
