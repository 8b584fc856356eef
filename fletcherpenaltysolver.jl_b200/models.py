"""Minimal NLPModels-style model interface (the *user model* side of the boundary) and the
synthetic problems named in BASELINE.json.

The reference consumes the user's model only through the NLPModels API
(jac_structure!/jac_coord!/jprod!/jtprod!/hprod!/ghjvprod!/obj/grad!/cons!);
these classes provide that API in Python with 0-based COO structures.  They are inputs to the
hot path, not part of it.
"""
import numpy as np
import scipy.sparse as sp


class NLPModelMeta:
    def __init__(self, nvar, ncon=0, x0=None, lcon=None, ucon=None, nnzj=0, name="generic",
                 lvar=None, uvar=None, lin=(), minimize=True):
        self.nvar, self.ncon, self.nnzj, self.name = nvar, ncon, nnzj, name
        self.x0 = np.zeros(nvar) if x0 is None else np.asarray(x0, dtype=np.float64)
        self.lcon = np.zeros(ncon) if lcon is None else np.asarray(lcon, dtype=np.float64)
        self.ucon = np.zeros(ncon) if ucon is None else np.asarray(ucon, dtype=np.float64)
        self.lvar = np.full(nvar, -np.inf) if lvar is None else np.asarray(lvar, dtype=np.float64)
        self.uvar = np.full(nvar, np.inf) if uvar is None else np.asarray(uvar, dtype=np.float64)
        self.lin = list(lin)
        self.nlin = len(self.lin)
        self.nln = [i for i in range(ncon) if i not in set(self.lin)]
        self.nnln = len(self.nln)
        self.minimize = minimize


class Counters:
    FIELDS = ("neval_obj", "neval_grad", "neval_cons", "neval_jac", "neval_jprod", "neval_jtprod",
              "neval_hprod", "neval_jhprod")

    def __init__(self):
        for f in self.FIELDS:
            setattr(self, f, 0)


class AbstractNLPModel:
    meta: NLPModelMeta

    def __init__(self):
        self.counters = Counters()

    # the subclass supplies: obj, grad, cons, jac_structure, jac_coord, hprod, ghjvprod
    def jprod(self, x, v):
        r, c = self.jac_structure()
        vals = self.jac_coord(x)
        out = np.zeros(self.meta.ncon)
        np.add.at(out, r, vals * v[c])
        return out

    def jtprod(self, x, u):
        r, c = self.jac_structure()
        vals = self.jac_coord(x)
        out = np.zeros(self.meta.nvar)
        np.add.at(out, c, vals * u[r])
        return out

    def ghjvprod(self, x, g, v):
        """(g' H_j(x) v)_j for each constraint j."""
        return np.zeros(self.meta.ncon)

    # linear / nonlinear rows (NLPModels' cons_lin!, cons_nln!, jac_lin_structure!, jac_nln_coord!, ...): generic
    # versions that filter the rows of cons / jac_structure / jac_coord by meta.lin / meta.nln
    def _rows_of(self, which):
        key = tuple(int(i) for i in which)
        cache = self.__dict__.setdefault("_rowsel", {})
        if key not in cache:
            r, c = self.jac_structure()
            r, c = np.asarray(r, dtype=np.int64), np.asarray(c, dtype=np.int64)
            pos = -np.ones(max(self.meta.ncon, 1), dtype=np.int64)
            pos[list(key)] = np.arange(len(key))
            sel = np.nonzero(pos[r] >= 0)[0] if len(r) else np.zeros(0, dtype=np.int64)
            cache[key] = (sel, pos[r[sel]], c[sel])
        return cache[key]

    def jac_lin_structure(self):
        _, r, c = self._rows_of(self.meta.lin)
        return r, c

    def jac_nln_structure(self):
        _, r, c = self._rows_of(self.meta.nln)
        return r, c

    def jac_lin_coord(self, x):
        return np.asarray(self.jac_coord(x), dtype=np.float64)[self._rows_of(self.meta.lin)[0]]

    def jac_nln_coord(self, x):
        return np.asarray(self.jac_coord(x), dtype=np.float64)[self._rows_of(self.meta.nln)[0]]

    def cons_lin(self, x):
        return np.asarray(self.cons(x), dtype=np.float64)[list(self.meta.lin)]

    def cons_nln(self, x):
        return np.asarray(self.cons(x), dtype=np.float64)[list(self.meta.nln)]


class CallableModel(AbstractNLPModel):
    """Small dense model from callables (the role ADNLPModel plays in the reference's tests).

    jac(x) returns the dense ncon x nvar Jacobian; its structure is the full dense COO.
    hess_f(x) and hess_c(x, j) return dense Hessians of the objective / j-th constraint.
    """

    def __init__(self, f, grad, cons, jac, hess_f, hess_c, x0, ncon, lcon=None, ucon=None,
                 name="callable", lin=()):
        super().__init__()
        nvar = len(x0)
        self._f, self._g, self._c, self._J, self._Hf, self._Hc = f, grad, cons, jac, hess_f, hess_c
        rows, cols = np.meshgrid(np.arange(ncon), np.arange(nvar), indexing="ij")
        self._rows = rows.ravel().astype(np.int64)
        self._cols = cols.ravel().astype(np.int64)
        self.meta = NLPModelMeta(nvar, ncon, x0=x0, lcon=lcon, ucon=ucon if ucon is not None else lcon,
                                 nnzj=nvar * ncon, name=name, lin=lin)

    def obj(self, x):
        self.counters.neval_obj += 1
        return float(self._f(x))

    def grad(self, x):
        self.counters.neval_grad += 1
        return np.asarray(self._g(x), dtype=np.float64)

    def cons(self, x):
        self.counters.neval_cons += 1
        return np.asarray(self._c(x), dtype=np.float64)

    def jac_structure(self):
        return self._rows, self._cols

    def jac_coord(self, x):
        self.counters.neval_jac += 1
        return np.asarray(self._J(x), dtype=np.float64).reshape(self.meta.ncon, self.meta.nvar).ravel()

    def hprod(self, x, y, v, obj_weight=1.0):
        self.counters.neval_hprod += 1
        H = obj_weight * np.asarray(self._Hf(x), dtype=np.float64)
        for j in range(self.meta.ncon):
            H = H + y[j] * np.asarray(self._Hc(x, j), dtype=np.float64)
        return H @ v

    def ghjvprod(self, x, g, v):
        return np.array([g @ (np.asarray(self._Hc(x, j)) @ v) for j in range(self.meta.ncon)])


def unit_test_model_sum(n=10):
    """f = x'x, c = sum(x) - 1  (test/unit-test.jl:17)."""
    return CallableModel(lambda x: x @ x, lambda x: 2 * x, lambda x: np.array([x.sum() - 1.0]),
                         lambda x: np.ones((1, n)), lambda x: 2 * np.eye(n),
                         lambda x, j: np.zeros((n, n)), np.zeros(n), 1, name="xtx-sum")


def unit_test_model_rosenbrock_circle():
    """f = (x1-1)^2 + 100 (x2 - x1^2)^2, c = x1^2 + x2^2 - 1  (test/unit-test.jl:79-85)."""
    f = lambda x: (x[0] - 1) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2
    g = lambda x: np.array([2 * (x[0] - 1) - 400 * x[0] * (x[1] - x[0] ** 2), 200 * (x[1] - x[0] ** 2)])
    c = lambda x: np.array([x[0] ** 2 + x[1] ** 2 - 1])
    J = lambda x: np.array([[2 * x[0], 2 * x[1]]])
    Hf = lambda x: np.array([[2 - 400 * x[1] + 1200 * x[0] ** 2, -400 * x[0]], [-400 * x[0], 200.0]])
    Hc = lambda x, j: 2 * np.eye(2)
    return CallableModel(f, g, c, J, Hf, Hc, np.zeros(2), 1, name="rosenbrock-circle")


class SparseQPModel(AbstractNLPModel):
    """min 1/2 x' diag(Q) x + q'x  s.t.  A x = b   (BASELINE configs C2 / C4).

    The Jacobian is constant; jac_coord returns the CSR values in COO order of (row, col)."""

    def __init__(self, Qdiag, q, A, b, name="sparse-qp"):
        super().__init__()
        A = sp.csr_matrix(A)
        A.sort_indices()
        self.A = A
        self.Q, self.q, self.b = np.asarray(Qdiag, float), np.asarray(q, float), np.asarray(b, float)
        coo = A.tocoo()
        self._rows = coo.row.astype(np.int64)
        self._cols = coo.col.astype(np.int64)
        self._vals = coo.data.astype(np.float64)
        m, n = A.shape
        self.meta = NLPModelMeta(n, m, x0=np.zeros(n), nnzj=A.nnz, name=name)

    def obj(self, x):
        self.counters.neval_obj += 1
        return float(0.5 * x @ (self.Q * x) + self.q @ x)

    def grad(self, x):
        self.counters.neval_grad += 1
        return self.Q * x + self.q

    def cons(self, x):
        self.counters.neval_cons += 1
        return self.A @ x - self.b

    def jac_structure(self):
        return self._rows, self._cols

    def jac_coord(self, x):
        self.counters.neval_jac += 1
        return self._vals

    def jprod(self, x, v):
        return self.A @ v

    def jtprod(self, x, u):
        return self.A.T @ u

    def hprod(self, x, y, v, obj_weight=1.0):
        self.counters.neval_hprod += 1
        return obj_weight * self.Q * v


class CurvedQPModel(SparseQPModel):
    """SparseQPModel with curved constraints: c_i(x) = (A x)_i + d_i/2 x_{k_i}^2 - b_i, k_i the column of the first
    stored entry of row i — so the Jacobian keeps A's sparsity pattern (its entry (i, k_i) becomes a_{i k_i} + d_i x_{k_i})
    while the constraint Hessians d_i e_k e_k' make `hprod` depend on y and `ghjvprod` nonzero: the smallest large sparse
    model that exercises the Val(1) Hessian product (src/model-Fletcherpenaltynlp.jl:572-634) with real terms."""

    def __init__(self, Qdiag, q, A, b, d, name="curved-qp"):
        super().__init__(Qdiag, q, A, b, name=name)
        self.d = np.asarray(d, float)
        assert np.all(np.diff(self.A.indptr) > 0), "every row needs an entry"
        self._first = self.A.indptr[:-1].astype(np.int64)         # position of (i, k_i) in the CSR / COO order
        self._k = self.A.indices[self._first].astype(np.int64)

    def cons(self, x):
        self.counters.neval_cons += 1
        return self.A @ x + 0.5 * self.d * x[self._k] ** 2 - self.b

    def jac_coord(self, x):
        self.counters.neval_jac += 1
        vals = self._vals.copy()
        vals[self._first] += self.d * x[self._k]
        return vals

    def jprod(self, x, v):
        return self.A @ v + self.d * x[self._k] * v[self._k]

    def jtprod(self, x, u):
        out = self.A.T @ u
        np.add.at(out, self._k, self.d * x[self._k] * u)
        return out

    def hprod(self, x, y, v, obj_weight=1.0):
        self.counters.neval_hprod += 1
        out = obj_weight * self.Q * v
        np.add.at(out, self._k, y * self.d * v[self._k])
        return out

    def ghjvprod(self, x, g, v):
        return g[self._k] * self.d * v[self._k]


# --------------------------------------------------------------------------------------------------
# synthetic generators (seeded)
# --------------------------------------------------------------------------------------------------
def window_random_jacobian(m, n, nnz_per_row, w=64, seed=1234, dtype=np.float64):
    """m x n CSR, `nnz_per_row` entries per row with columns drawn without replacement from the
    window |j - (n/m) i| <= w, values N(0,1).  Locality keeps the LDL' fill bounded (SURVEY §7
    hard part 3) and the gathers cache friendly."""
    rng = np.random.default_rng(seed)
    ratio = n / m
    width = 2 * w + 1
    assert nnz_per_row <= width
    centers = np.minimum(np.maximum((np.arange(m) * ratio).astype(np.int64), w), n - 1 - w)
    # draw distinct offsets per row: argsort of random keys, take the first nnz_per_row
    keys = rng.random((m, width), dtype=np.float32)
    offs = np.argpartition(keys, nnz_per_row - 1, axis=1)[:, :nnz_per_row].astype(np.int64) - w
    cols = centers[:, None] + offs
    cols.sort(axis=1)
    vals = rng.standard_normal((m, nnz_per_row)).astype(dtype)
    indptr = np.arange(0, (m + 1) * nnz_per_row, nnz_per_row, dtype=np.int64)
    return sp.csr_matrix((vals.ravel(), cols.ravel(), indptr), shape=(m, n))


def sparse_qp(n, m, nnz_per_row=10, w=64, seed=1234, rank_deficient_frac=0.0):
    """C2 (n=1e5, m=5e4, 10 nnz/row) / C4 (n=1e6, m=5e5, 20 nnz/row, 1% duplicated rows)."""
    rng = np.random.default_rng(seed + 1)
    A = window_random_jacobian(m, n, nnz_per_row, w, seed)
    if rank_deficient_frac > 0:
        A = A.tolil(copy=False) if False else A
        k = max(1, int(rank_deficient_frac * m))
        dst = rng.choice(m, size=k, replace=False)
        src = (dst + 1 + rng.integers(0, 8, size=k)) % m
        indptr, indices, data = A.indptr, A.indices.copy(), A.data.copy()
        per = indptr[1] - indptr[0]
        for d, s in zip(dst, src):
            indices[d * per:(d + 1) * per] = indices[s * per:(s + 1) * per]
            data[d * per:(d + 1) * per] = data[s * per:(s + 1) * per]
        A = sp.csr_matrix((data, indices, indptr), shape=(m, n))
    Q = 1.0 + rng.random(n)
    q = rng.standard_normal(n)
    xstar = rng.standard_normal(n)
    b = A @ xstar
    return SparseQPModel(Q, q, A, b, name=f"sparse-qp-n{n}-m{m}")


def poisson_control(N, alpha=1e-2):
    """C3: 5-point Laplacian L on an N x N interior grid of (-1,1)^2, A = [L  -I]
    (m = N^2, n = 2 N^2), tracking objective 1/2|y - yd|^2 + alpha/2 |u|^2, c = L y - u - h
    (docs/assets/example.jl:47-58 is the model for the data)."""
    hgrid = 2.0 / (N + 1)
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(N, N))
    I = sp.identity(N)
    L = (sp.kron(I, T) + sp.kron(T, I)) / hgrid ** 2
    m = N * N
    A = sp.hstack([L, -sp.identity(m)]).tocsr()
    xs = -1 + hgrid * (1 + np.arange(N))
    X1, X2 = np.meshgrid(xs, xs, indexing="ij")
    omega = np.pi - 1.0 / 8
    hvec = (-np.sin(omega * X1) * np.sin(omega * X2)).ravel()
    yd = (-X1 ** 2).ravel()
    Q = np.concatenate([np.ones(m), alpha * np.ones(m)])
    q = np.concatenate([-yd, np.zeros(m)])
    return SparseQPModel(Q, q, A, hvec, name=f"poisson-control-{N}")


# --------------------------------------------------------------------------------------------------
# the reference's own small test problems (BASELINE configs C1 / C5), derivatives written by hand
# (the reference gets them from ADNLPModels' automatic differentiation)
# --------------------------------------------------------------------------------------------------
def _cm(f, g, c, J, Hf, Hc, x0, ncon, lcon=None, ucon=None, name=""):
    return CallableModel(f, g, c, J, Hf, Hc, np.asarray(x0, dtype=np.float64), ncon, lcon=lcon, ucon=ucon, name=name)


def reference_test_problem(name):
    """One of the problems `fps_solve` is tested on in the reference:
    test/test-2.jl (rosenbrock_sum, simple, hs6, hs7, hs8, hs9, hs26, hs27, unbounded_quad_penalty, spurious, flt),
    test/rank-deficient.jl:23-29 (hs61), docs/src/tutorial.md (hs28), docs/src/fine-tuneFPS.md:25-33
    (readme_eq: Rosenbrock with x1 x2 = 1) and README.md:59-65 (readme_ineq: 0 <= x1 x2 - 1 <= 1)."""
    A = np.array
    Z = lambda n: (lambda x, j: np.zeros((n, n)))
    rosen_f = lambda x: (x[0] - 1.0) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2
    rosen_g = lambda x: A([2 * (x[0] - 1) - 400 * x[0] * (x[1] - x[0] ** 2), 200 * (x[1] - x[0] ** 2)])
    rosen_H = lambda x: A([[2 - 400 * x[1] + 1200 * x[0] ** 2, -400 * x[0]], [-400 * x[0], 200.0]])
    if name == "rosenbrock_sum":
        return _cm(rosen_f, rosen_g, lambda x: A([x[0] + x[1]]), lambda x: A([[1.0, 1.0]]), rosen_H, Z(2),
                   [-1.2, 1.0], 1, lcon=[1.0], name=name)
    if name == "simple":
        return unit_test_model_sum(10)
    if name == "hs6":
        return _cm(lambda x: (1 - x[0]) ** 2, lambda x: A([-2 * (1 - x[0]), 0.0]),
                   lambda x: A([10 * (x[1] - x[0] ** 2)]), lambda x: A([[-20 * x[0], 10.0]]),
                   lambda x: A([[2.0, 0.0], [0.0, 0.0]]), lambda x, j: A([[-20.0, 0.0], [0.0, 0.0]]),
                   [-1.2, 1.0], 1, name=name)
    if name == "hs7":
        return _cm(lambda x: np.log(1 + x[0] ** 2) - x[1], lambda x: A([2 * x[0] / (1 + x[0] ** 2), -1.0]),
                   lambda x: A([(1 + x[0] ** 2) ** 2 + x[1] ** 2 - 4]),
                   lambda x: A([[4 * x[0] * (1 + x[0] ** 2), 2 * x[1]]]),
                   lambda x: A([[2 * (1 - x[0] ** 2) / (1 + x[0] ** 2) ** 2, 0.0], [0.0, 0.0]]),
                   lambda x, j: A([[4 + 12 * x[0] ** 2, 0.0], [0.0, 2.0]]), [2.0, 2.0], 1, name=name)
    if name == "hs8":
        return _cm(lambda x: -1.0, lambda x: np.zeros(2), lambda x: A([x[0] ** 2 + x[1] ** 2 - 25, x[0] * x[1] - 9]),
                   lambda x: A([[2 * x[0], 2 * x[1]], [x[1], x[0]]]), lambda x: np.zeros((2, 2)),
                   lambda x, j: 2 * np.eye(2) if j == 0 else A([[0.0, 1.0], [1.0, 0.0]]), [2.0, 1.0], 2, name=name)
    if name == "hs9":
        a, b = np.pi / 12, np.pi / 16
        return _cm(lambda x: np.sin(a * x[0]) * np.cos(b * x[1]),
                   lambda x: A([a * np.cos(a * x[0]) * np.cos(b * x[1]), -b * np.sin(a * x[0]) * np.sin(b * x[1])]),
                   lambda x: A([4 * x[0] - 3 * x[1]]), lambda x: A([[4.0, -3.0]]),
                   lambda x: A([[-a * a * np.sin(a * x[0]) * np.cos(b * x[1]), -a * b * np.cos(a * x[0]) * np.sin(b * x[1])],
                                [-a * b * np.cos(a * x[0]) * np.sin(b * x[1]), -b * b * np.sin(a * x[0]) * np.cos(b * x[1])]]),
                   Z(2), [0.0, 0.0], 1, name=name)
    if name == "hs26":
        def Hf(x):
            t = 12 * (x[1] - x[2]) ** 2
            return A([[2.0, -2.0, 0.0], [-2.0, 2 + t, -t], [0.0, -t, t]])
        return _cm(lambda x: (x[0] - x[1]) ** 2 + (x[1] - x[2]) ** 4,
                   lambda x: A([2 * (x[0] - x[1]), -2 * (x[0] - x[1]) + 4 * (x[1] - x[2]) ** 3, -4 * (x[1] - x[2]) ** 3]),
                   lambda x: A([(1 + x[1] ** 2) * x[0] + x[2] ** 4 - 3]),
                   lambda x: A([[1 + x[1] ** 2, 2 * x[0] * x[1], 4 * x[2] ** 3]]), Hf,
                   lambda x, j: A([[0.0, 2 * x[1], 0.0], [2 * x[1], 2 * x[0], 0.0], [0.0, 0.0, 12 * x[2] ** 2]]),
                   [-2.6, 2.0, 2.0], 1, name=name)
    if name == "hs27":
        return _cm(lambda x: 0.01 * (x[0] - 1) ** 2 + (x[1] - x[0] ** 2) ** 2,
                   lambda x: A([0.02 * (x[0] - 1) - 4 * x[0] * (x[1] - x[0] ** 2), 2 * (x[1] - x[0] ** 2), 0.0]),
                   lambda x: A([x[0] + x[2] ** 2 + 1.0]), lambda x: A([[1.0, 0.0, 2 * x[2]]]),
                   lambda x: A([[0.02 - 4 * x[1] + 12 * x[0] ** 2, -4 * x[0], 0.0], [-4 * x[0], 2.0, 0.0], [0.0, 0.0, 0.0]]),
                   lambda x, j: np.diag([0.0, 0.0, 2.0]), [2.0, 2.0, 2.0], 1, name=name)
    if name == "unbounded_quad_penalty":
        return _cm(lambda x: x[0] ** 3 * x[1] ** 3, lambda x: A([3 * x[0] ** 2 * x[1] ** 3, 3 * x[0] ** 3 * x[1] ** 2]),
                   lambda x: A([x[0] ** 2 + x[1] ** 2 - 1]), lambda x: A([[2 * x[0], 2 * x[1]]]),
                   lambda x: A([[6 * x[0] * x[1] ** 3, 9 * x[0] ** 2 * x[1] ** 2], [9 * x[0] ** 2 * x[1] ** 2, 6 * x[0] ** 3 * x[1]]]),
                   lambda x, j: 2 * np.eye(2), [0.0, 0.0], 1, name=name)
    if name == "spurious":
        return _cm(lambda x: 0.0, lambda x: np.zeros(1), lambda x: A([x[0] ** 3 + x[0] - 2.0]),
                   lambda x: A([[3 * x[0] ** 2 + 1]]), lambda x: np.zeros((1, 1)), lambda x, j: A([[6 * x[0]]]),
                   [0.0], 1, name=name)
    if name == "flt":
        return _cm(lambda x: (x[1] - 1) ** 2, lambda x: A([0.0, 2 * (x[1] - 1)]), lambda x: A([x[0] ** 2, x[0] ** 3]),
                   lambda x: A([[2 * x[0], 0.0], [3 * x[0] ** 2, 0.0]]), lambda x: np.diag([0.0, 2.0]),
                   lambda x, j: np.diag([2.0, 0.0]) if j == 0 else np.diag([6 * x[0], 0.0]), [1.0, 0.0], 2, name="FLT")
    if name == "hs61":
        return _cm(lambda x: 4 * x[0] ** 2 + 2 * x[1] ** 2 + 2 * x[2] ** 2 - 33 * x[0] + 16 * x[1] - 24 * x[2],
                   lambda x: A([8 * x[0] - 33, 4 * x[1] + 16, 4 * x[2] - 24]),
                   lambda x: A([3 * x[0] - 2 * x[1] ** 2 - 7, 4 * x[0] - x[2] ** 2 - 11]),
                   lambda x: A([[3.0, -4 * x[1], 0.0], [4.0, 0.0, -2 * x[2]]]), lambda x: np.diag([8.0, 4.0, 4.0]),
                   lambda x, j: np.diag([0.0, -4.0, 0.0]) if j == 0 else np.diag([0.0, 0.0, -2.0]),
                   [0.0, 0.0, 0.0], 2, name=name)
    if name == "hs28":
        return _cm(lambda x: (x[0] + x[1]) ** 2 + (x[1] + x[2]) ** 2,
                   lambda x: A([2 * (x[0] + x[1]), 2 * (x[0] + x[1]) + 2 * (x[1] + x[2]), 2 * (x[1] + x[2])]),
                   lambda x: A([x[0] + 2 * x[1] + 3 * x[2] - 1]), lambda x: A([[1.0, 2.0, 3.0]]),
                   lambda x: A([[2.0, 2.0, 0.0], [2.0, 4.0, 2.0], [0.0, 2.0, 2.0]]), Z(3), [-4.0, 1.0, 1.0], 1, name=name)
    if name in ("readme_eq", "readme_ineq"):
        f = lambda x: 100 * (x[1] - x[0] ** 2) ** 2 + (x[0] - 1) ** 2
        return _cm(f, rosen_g, lambda x: A([x[0] * x[1] - 1]), lambda x: A([[x[1], x[0]]]), rosen_H,
                   lambda x, j: A([[0.0, 1.0], [1.0, 0.0]]), [-1.2, 1.0], 1, lcon=[0.0],
                   ucon=[0.0] if name == "readme_eq" else [1.0], name=name)
    raise KeyError(name)


REFERENCE_TEST_PROBLEMS = ("rosenbrock_sum", "simple", "hs6", "hs7", "hs8", "hs9", "hs26", "hs27",
                           "unbounded_quad_penalty", "spurious", "flt", "hs61", "hs28", "readme_eq")


class SlackModel(AbstractNLPModel):
    """Equalities + bounded slacks for a model with inequalities: c_j(x) - s_j = 0, lcon_j <= s_j <= ucon_j
    (the role of NLPModelsModifiers.SlackModel at src/FletcherPenaltySolver.jl:140-143).  Variables [x; s],
    one slack per inequality row in row order."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        me = model.meta
        self.ineq = np.flatnonzero(np.asarray(me.lcon) < np.asarray(me.ucon))
        ns, n = len(self.ineq), me.nvar
        lcon, ucon = np.array(me.lcon, dtype=np.float64), np.array(me.ucon, dtype=np.float64)
        lvar = np.concatenate([me.lvar, lcon[self.ineq]])
        uvar = np.concatenate([me.uvar, ucon[self.ineq]])
        lcon[self.ineq] = 0.0
        ucon[self.ineq] = 0.0
        r, c = model.jac_structure()
        self._rows = np.concatenate([np.asarray(r, dtype=np.int64), self.ineq.astype(np.int64)])
        self._cols = np.concatenate([np.asarray(c, dtype=np.int64), n + np.arange(ns, dtype=np.int64)])
        self.meta = NLPModelMeta(n + ns, me.ncon, x0=np.concatenate([me.x0, np.zeros(ns)]), lcon=lcon, ucon=ucon,
                                 nnzj=len(self._rows), name=me.name + "-slack", lvar=lvar, uvar=uvar)
        self.n = n

    def obj(self, x):
        self.counters.neval_obj += 1
        return self.model.obj(x[:self.n])

    def grad(self, x):
        self.counters.neval_grad += 1
        return np.concatenate([self.model.grad(x[:self.n]), np.zeros(len(self.ineq))])

    def cons(self, x):
        self.counters.neval_cons += 1
        c = np.array(self.model.cons(x[:self.n]), dtype=np.float64)
        c[self.ineq] -= x[self.n:]
        return c

    def jac_structure(self):
        return self._rows, self._cols

    def jac_coord(self, x):
        self.counters.neval_jac += 1
        return np.concatenate([self.model.jac_coord(x[:self.n]), -np.ones(len(self.ineq))])

    def hprod(self, x, y, v, obj_weight=1.0):
        self.counters.neval_hprod += 1
        return np.concatenate([self.model.hprod(x[:self.n], y, v[:self.n], obj_weight=obj_weight),
                               np.zeros(len(self.ineq))])

    def ghjvprod(self, x, g, v):
        return self.model.ghjvprod(x[:self.n], g[:self.n], v[:self.n])
