"""FletcherPenaltyNLP — caller side of the hot path (mirror of src/model-Fletcherpenaltynlp.jl).

Only what surrounds the 2-RHS solves is here: the memoised `_compute_ys_gs!` (:234-252), `obj`
(:352-370), `grad!` (:372-401), `objgrad!` (:403-437) and the two `hprod!` methods (:521-634).
`hess_coord!` / `hess_structure!` (dense pinv debug path, :439-519) are out of scope.
The solves themselves go through qdsolver.solve_two_* -> libfpsb200.so.
"""
import numpy as np

from .qdsolver import (LDLtSolver, solve_two_extras, solve_two_least_squares, solve_two_mixed)


class FletcherPenaltyNLP:
    """FletcherPenaltyNLP(nlp, sigma, rho, delta, hessian_approx; qds = LDLtSolver(nlp, 0.0))."""

    def __init__(self, nlp, sigma=1.0, rho=0.0, delta=0.0, hessian_approx=2, x0=None, *, qds=None,
                 explicit_linear_constraints=False, consistent_gradient=False):
        assert hessian_approx in (1, 2)
        # Reference quirk: grad! hands +ys to hprod! (src/model-Fletcherpenaltynlp.jl:382, 415)
        # where the derivative of obj needs H(x, -ys), the matrix hprod! itself uses (:538-539).  The two only
        # differ when the constraints have curvature and c(x) != 0.  False = the reference's formula
        # (drop-in parity), True = the exact gradient of obj.
        self.consistent_gradient = consistent_gradient
        self.nlp = nlp
        self.explicit_linear_constraints = explicit_linear_constraints
        nvar = nlp.meta.nvar
        npen = nlp.meta.nnln if explicit_linear_constraints else nlp.meta.ncon
        self.nvar, self.npen = nvar, npen
        self.x0 = nlp.meta.x0 if x0 is None else x0
        self.shahx = 0
        self.fx = float("nan")
        self.cx = np.empty(npen)
        self.gx = np.empty(nvar)
        self.ys = np.empty(npen)
        self.gs = np.empty(nvar)
        self.xk = np.zeros(nvar)
        self.v = np.empty(nvar)
        self.w = np.empty(npen)
        self.sigma, self.rho, self.delta, self.eta = sigma, rho, delta, 0.0
        self.qdsolver = qds if qds is not None else LDLtSolver(nlp, 0.0)
        self.hessian_approx = hessian_approx
        self.neval = dict(obj=0, grad=0, hprod=0)

    # δ is read at solve time under the reference's field name
    # ------------------------------------------------------------------------------------------
    def cons_norhs(self, x):
        """cons(nlp, x) - lcon   (src/model-Fletcherpenaltynlp.jl:260-269)."""
        nlp = self.nlp
        if nlp.meta.ncon == 0:
            return np.empty(0)
        if self.explicit_linear_constraints:
            return nlp.cons_nln(x) - nlp.meta.lcon[nlp.meta.nln]
        return nlp.cons(x) - nlp.meta.lcon

    def _hprod_nln(self, x, y, v, obj_weight=1.0):
        """hprod_nln!: multipliers of the penalised (nonlinear) rows only   (:277-290)."""
        if self.explicit_linear_constraints and self.nlp.meta.ncon > 0:
            lag_mul = np.zeros(self.nlp.meta.ncon)
            lag_mul[list(self.nlp.meta.nln)] = y
            return self.nlp.hprod(x, lag_mul, v, obj_weight=obj_weight)
        return self.nlp.hprod(x, y, v, obj_weight=obj_weight)

    def _ghjvprod_nln(self, x, g, v):
        out = self.nlp.ghjvprod(x, g, v)
        return np.asarray(out)[list(self.nlp.meta.nln)] if self.explicit_linear_constraints else out

    def linear_system2(self, x):
        """p1, q1, p2, q2 = solve_two_mixed(nlp, x, gx, cx)   (:215-227)."""
        return solve_two_mixed(self, x, self.gx, self.cx)

    def _compute_ys_gs(self, x):
        """memoised on hash(x) only (reference quirk D-1)   (:234-252)."""
        shahx = hash(np.ascontiguousarray(x, dtype=np.float64).tobytes())
        if shahx != self.shahx:
            self.shahx = shahx
            self.fx = self.nlp.obj(x)
            self.gx = self.nlp.grad(x)
            self.cx = self.cons_norhs(x)
            p1, q1, p2, q2 = self.linear_system2(x)
            self.gs = p1 + self.sigma * p2
            self.ys = q1 + self.sigma * q2
            self.v = np.array(p2, copy=True)
            self.w = np.array(q2, copy=True)
        return self.gs, self.ys, self.v, self.w

    # ------------------------------------------------------------------------------------------
    def obj(self, x):
        self.neval["obj"] += 1
        self._compute_ys_gs(x)
        c = self.cx
        fx = self.fx - c @ self.ys + self.rho / 2 * (c @ c)
        if self.eta > 0.0:
            fx += self.eta / 2 * np.linalg.norm(x - self.xk) ** 2
        return fx

    def grad(self, x):
        self.neval["grad"] += 1
        gs, ys, v, w = self._compute_ys_gs(x)
        c = self.cx
        Hsv = self._hprod_nln(x, -ys if self.consistent_gradient else ys, v, obj_weight=1.0)
        Sstw = self._hprod_nln(x, w, gs, obj_weight=0.0)
        gx = gs - Hsv + self.sigma * v + Sstw
        if self.rho > 0.0:
            gx = gx + self.rho * self._jtprod(x, c)
        if self.eta > 0.0:
            gx = gx + self.eta * (x - self.xk)
        return gx

    def objgrad(self, x):
        gx = self.grad(x)
        self.neval["obj"] += 1
        c = self.cx
        fx = self.fx - c @ self.ys
        if self.rho > 0.0:
            fx += self.rho / 2 * (c @ c)
        if self.eta > 0.0:
            fx += self.eta / 2 * np.linalg.norm(x - self.xk) ** 2
        return fx, gx

    def _jprod(self, x, v):
        """jprod! on the device-resident Jacobian (refreshed by the last solve_two_mixed at x);
        QDSolver subtypes without a device handle fall back to the user model's jprod!."""
        h = getattr(self.qdsolver, "handle", None)
        if h is None:
            if self.explicit_linear_constraints:
                r, c = self.nlp.jac_nln_structure()
                out = np.zeros(self.npen)
                np.add.at(out, r, self.nlp.jac_nln_coord(x) * np.asarray(v)[c])
                return out
            return self.nlp.jprod(x, v)
        return h.jprod(np.ascontiguousarray(v, dtype=np.float64))

    def _jtprod(self, x, u):
        h = getattr(self.qdsolver, "handle", None)
        if h is None:
            if self.explicit_linear_constraints:
                r, c = self.nlp.jac_nln_structure()
                out = np.zeros(self.nvar)
                np.add.at(out, c, self.nlp.jac_nln_coord(x) * np.asarray(u)[r])
                return out
            return self.nlp.jtprod(x, u)
        return h.jtprod(np.ascontiguousarray(u, dtype=np.float64))

    def hprod(self, x, v, obj_weight=1.0):
        self.neval["hprod"] += 1
        sigma, rho = self.sigma, self.rho
        gs, ys, _, _ = self._compute_ys_gs(x)
        c = self.cx
        mys = -ys
        Hsv = self._hprod_nln(x, mys, v, obj_weight=1.0)
        p1, _, p2, _ = solve_two_least_squares(self, x, v, Hsv)
        p2 = np.array(p2, copy=True)           # must survive solve_two_extras
        Ptv = v - p1
        HsPtv = self._hprod_nln(x, mys, Ptv, obj_weight=1.0)
        if self.hessian_approx == 2:
            Hv = p2 - HsPtv + 2 * sigma * Ptv
            if rho > 0.0:
                Jv = self._jprod(x, v)
                JtJv = self._jtprod(x, Jv)
                Hcv = self._hprod_nln(x, c, v, obj_weight=0.0)
                Hv = Hv + Hcv + rho * JtJv
        else:
            Ssv = self._ghjvprod_nln(x, gs, v)
            invJtJJv, invJtJSsv = solve_two_extras(self, x, v, Ssv)
            JtinvJtJSsv = self._jtprod(x, invJtJSsv)
            Hv = p2 - HsPtv + 2 * sigma * Ptv - JtinvJtJSsv
            SsinvJtJJv = self._hprod_nln(x, invJtJJv, gs, obj_weight=0.0)
            Hv = Hv - SsinvJtJJv
            if rho > 0.0:
                Jv = self._jprod(x, v)
                JtJv = self._jtprod(x, Jv)
                Hcv = self._hprod_nln(x, c, v, obj_weight=0.0)
                Hv = Hv + rho * (Hcv + JtJv)
        if self.eta > 0.0:
            Hv = Hv + self.eta * v
        return obj_weight * Hv
