"""Device-resident FletcherPenaltyNLP: the same evaluation as fletcher_nlp.py
(src/model-Fletcherpenaltynlp.jl:234-252, 352-437, 521-570) with x, g, c, ys, gs, v, w and every
intermediate living in HBM.  The 2-RHS solves are called with device pointers (FPSB_DEVICE), the
vector combinations either side of them are the fused fpsb_fp_* kernels of libfpsb200, and the memo
key is computed on the device (fpsb_fp_hash) instead of hash(x) on the host.

The user model must evaluate on the device: `obj(x) -> float`, `grad(x)`, `cons(x)`, `jac_coord(x)`,
`hprod(x, y, v, obj_weight)` taking / returning torch CUDA float64 tensors (DeviceSparseQP below is
the synthetic model of BASELINE configs C2 / C4, DeviceCurvedQP its variant with curved constraints).  Both Hessian
variants are device-resident: Val(2), and Val(1) (round 2) which adds `ghjvprod(x, g, v)` on the device model and the
`solve_two_extras` pair with device pointers (src/model-Fletcherpenaltynlp.jl:572-634).
"""
import ctypes as C

import numpy as np

from . import _lib
from .models import AbstractNLPModel, NLPModelMeta
from .qdsolver import LDLtSolver, solve_two_extras, solve_two_least_squares, solve_two_mixed


def _dp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class DeviceSparseQP(AbstractNLPModel):
    """min 1/2 x' diag(Q) x + q'x  s.t.  A x = b  with all data on the GPU (the model side is user code:
    it uses torch for its own arithmetic)."""

    def __init__(self, host_qp, device="cuda"):
        import torch
        super().__init__()
        self.host = host_qp
        A = host_qp.A
        self.meta = NLPModelMeta(A.shape[1], A.shape[0], x0=np.zeros(A.shape[1]), nnzj=A.nnz, name=host_qp.meta.name + "-device")
        self._rows, self._cols = host_qp.jac_structure()
        f64 = dict(dtype=torch.float64, device=device)
        self.Q = torch.tensor(host_qp.Q, **f64)
        self.q = torch.tensor(host_qp.q, **f64)
        self.b = torch.tensor(host_qp.b, **f64)
        self._vals = torch.tensor(host_qp._vals, **f64)          # the values of A in the COO order of jac_structure
        self._device = device
        self._H = None          # the model's own operator handle for c(x) = A x - b (hand-written SpMV, no cuSPARSE)

    def _op(self):
        if self._H is None:
            import torch
            from .qdsolver import B200Handle
            d = torch.device(self._device)
            idx = d.index if d.index is not None else torch.cuda.current_device()
            self._H = B200Handle(self.meta.nvar, self.meta.ncon, self._rows, self._cols, device=idx)
            self._H.set_jac_values(self._vals)
        return self._H

    def obj(self, x):
        return float(0.5 * (x * self.Q * x).sum() + self.q @ x)

    def grad(self, x):
        return self.Q * x + self.q

    def cons(self, x):
        return self._op().jprod(x) - self.b

    def jac_structure(self):
        return self._rows, self._cols

    def jac_coord(self, x):
        return self._vals

    def hprod(self, x, y, v, obj_weight=1.0):
        return obj_weight * self.Q * v


class DeviceCurvedQP(DeviceSparseQP):
    """models.CurvedQPModel with all data on the GPU: curved constraints, so `hprod` depends on y and `ghjvprod` is not
    zero (the model side is user code: torch for its own arithmetic, the repo's SpMV for A x)."""

    def __init__(self, host_model, device="cuda"):
        import torch
        super().__init__(host_model, device=device)
        f64 = dict(dtype=torch.float64, device=device)
        self.d = torch.tensor(host_model.d, **f64)
        self._k = torch.tensor(host_model._k, dtype=torch.int64, device=device)
        self._first = torch.tensor(host_model._first, dtype=torch.int64, device=device)

    def cons(self, x):
        return self._op().jprod(x) + 0.5 * self.d * x[self._k] ** 2 - self.b

    def jac_coord(self, x):
        vals = self._vals.clone()
        vals[self._first] += self.d * x[self._k]
        return vals

    def hprod(self, x, y, v, obj_weight=1.0):
        out = obj_weight * self.Q * v
        out.index_add_(0, self._k, y * self.d * v[self._k])
        return out

    def ghjvprod(self, x, g, v):
        return g[self._k] * self.d * v[self._k]


class DeviceFletcherPenaltyNLP:
    """FletcherPenaltyNLP(nlp, sigma, rho, delta, Val(1) | Val(2); qds = ...) with device-resident state."""

    def __init__(self, nlp, sigma=1.0, rho=0.0, delta=0.0, hessian_approx=2, *, qds=None, device="cuda",
                 consistent_gradient=False):
        import torch
        assert hessian_approx in (1, 2), "hessian_approx is Val(1) or Val(2)"
        self.consistent_gradient = consistent_gradient      # see FletcherPenaltyNLP
        self.torch = torch
        self.nlp = nlp
        self.explicit_linear_constraints = False
        self.nvar, self.npen = nlp.meta.nvar, nlp.meta.ncon
        self.sigma, self.rho, self.delta, self.eta = sigma, rho, delta, 0.0
        self.qdsolver = qds if qds is not None else LDLtSolver(nlp, 0.0)
        self.handle = self.qdsolver.handle
        self.hessian_approx = int(hessian_approx)
        f64 = dict(dtype=torch.float64, device=device)
        n, m = self.nvar, self.npen
        self.key = None
        self.fx = float("nan")
        self.gx = self.cx = None
        self.gs, self.v, self.xk = (torch.empty(n, **f64) for _ in range(3))
        self.xk.zero_()
        self.ys, self.w = (torch.empty(m, **f64) for _ in range(2))
        self.neval = dict(obj=0, grad=0, hprod=0)
        self._lib = _lib.lib()

    @property
    def shahx(self):
        return self.key

    @shahx.setter
    def shahx(self, value):          # `shahx = 0` drops the memo (what reinit! does to the sub-state)
        self.key = None if not value else value

    def _hash(self, x):
        self.handle._follow_torch_stream(x)      # every fpsb_fp_* call is ordered against torch's current stream
        k = C.c_uint64()
        _lib.check(self._lib.fpsb_fp_hash(self.handle.h, _dp(x), C.byref(k)), "fpsb_fp_hash")
        return k.value

    def _compute_ys_gs(self, x):
        key = self._hash(x)
        if key != self.key:
            self.key = key
            self.fx = self.nlp.obj(x)
            self.gx = self.nlp.grad(x)
            self.cx = self.nlp.cons(x)          # lcon = 0 for the equality models handled here
            p1, q1, p2, q2 = solve_two_mixed(self, x, self.gx, self.cx)
            _lib.check(self._lib.fpsb_fp_ys_gs(self.handle.h, C.c_double(self.sigma), _dp(p1), _dp(q1), _dp(p2), _dp(q2),
                                               _dp(self.gs), _dp(self.ys), _dp(self.v), _dp(self.w)), "fpsb_fp_ys_gs")
        return self.gs, self.ys, self.v, self.w

    def obj(self, x):
        self.neval["obj"] += 1
        self._compute_ys_gs(x)
        phi = C.c_double()
        _lib.check(self._lib.fpsb_fp_obj(self.handle.h, C.c_double(self.fx), C.c_double(self.rho), C.c_double(self.eta),
                                         _dp(self.cx), _dp(self.ys), _dp(x), _dp(self.xk), C.byref(phi)), "fpsb_fp_obj")
        return phi.value

    def grad(self, x):
        self.neval["grad"] += 1
        gs, ys, v, w = self._compute_ys_gs(x)
        Hsv = self.nlp.hprod(x, -ys if self.consistent_gradient else ys, v, obj_weight=1.0)
        Sstw = self.nlp.hprod(x, w, gs, obj_weight=0.0)
        Jtc = self.handle.jtprod(self.cx) if self.rho > 0.0 else None
        g = self.torch.empty_like(x)
        _lib.check(self._lib.fpsb_fp_grad(self.handle.h, C.c_double(self.sigma), C.c_double(self.rho), C.c_double(self.eta),
                                          _dp(gs), _dp(Hsv), _dp(v), _dp(Sstw), _dp(Jtc), _dp(x), _dp(self.xk), _dp(g)),
                   "fpsb_fp_grad")
        return g

    def objgrad(self, x):
        g = self.grad(x)
        return self.obj(x), g

    def hprod(self, x, v, obj_weight=1.0):
        self.neval["hprod"] += 1
        gs, ys, _, _ = self._compute_ys_gs(x)
        mys = -ys
        Hsv = self.nlp.hprod(x, mys, v, obj_weight=1.0)
        p1, _, p2, _ = solve_two_least_squares(self, x, v, Hsv)
        Ptv = self.torch.empty_like(v)
        _lib.check(self._lib.fpsb_fp_ptv(self.handle.h, _dp(v), _dp(p1), _dp(Ptv)), "fpsb_fp_ptv")
        HsPtv = self.nlp.hprod(x, mys, Ptv, obj_weight=1.0)
        Hcv = JtJv = None
        if self.rho > 0.0:
            JtJv = self.handle.jtprod(self.handle.jprod(v))
            Hcv = self.nlp.hprod(x, self.cx, v, obj_weight=0.0)
        Hv = self.torch.empty_like(v)
        if self.hessian_approx == 1:
            # the exact Hessian's two extra terms (src/model-Fletcherpenaltynlp.jl:603-634): ghjvprod on the device model,
            # the MINRES / LSQR (or LDLt) pair of solve_two_extras with device pointers, one product with A'
            Ssv = self.nlp.ghjvprod(x, gs, v)
            invJtJJv, invJtJSsv = solve_two_extras(self, x, v, Ssv)
            JtinvJtJSsv = self.handle.jtprod(invJtJSsv)
            SsinvJtJJv = self.nlp.hprod(x, invJtJJv, gs, obj_weight=0.0)
            _lib.check(self._lib.fpsb_fp_hprod1(self.handle.h, C.c_double(self.sigma), C.c_double(self.rho), C.c_double(self.eta),
                                                C.c_double(obj_weight), _dp(p2), _dp(HsPtv), _dp(Ptv), _dp(JtinvJtJSsv),
                                                _dp(SsinvJtJJv), _dp(Hcv), _dp(JtJv), _dp(v), _dp(Hv)), "fpsb_fp_hprod1")
            return Hv
        _lib.check(self._lib.fpsb_fp_hprod2(self.handle.h, C.c_double(self.sigma), C.c_double(self.rho), C.c_double(self.eta),
                                            C.c_double(obj_weight), _dp(p2), _dp(HsPtv), _dp(Ptv), _dp(Hcv), _dp(JtJv), _dp(v),
                                            _dp(Hv)), "fpsb_fp_hprod2")
        return Hv
