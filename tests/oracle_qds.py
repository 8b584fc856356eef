"""Test-only QDSolver subtypes backed by the CPU oracle (lives under tests/, never shipped)."""
import numpy as np

from fpsb200.qdsolver import QDSolver
from oracle import oracle as O


class OracleLDLt(QDSolver):
    def __init__(self, nlp, _zero=0.0, P=None, **kw):
        self.rows, self.cols = nlp.jac_structure()
        n, m = nlp.meta.nvar, nlp.meta.ncon
        kw = {k: v for k, v in kw.items() if k in ("ldlt_tol", "ldlt_r1", "ldlt_r2")}   # like the reference's kwargs...
        self.o = O.LDLtOracle(n, m, self.rows, self.cols, np.arange(n + m) if P is None else P, **kw)
        self.handle = None

    def solve_two_mixed(self, fpnlp, x, rhs1, rhs2):
        return self.o.solve_two_mixed(fpnlp.nlp.jac_coord(x), fpnlp.delta, rhs1, rhs2)[:4]

    def solve_two_least_squares(self, fpnlp, x, rhs1, rhs2):
        return self.o.solve_two_least_squares(rhs1, rhs2)[:4]

    def solve_two_extras(self, fpnlp, x, rhs1, rhs2):
        self.o.jvals = np.asarray(fpnlp.nlp.jac_coord(x), dtype=np.float64)
        return self.o.solve_two_extras(fpnlp.delta, rhs1, rhs2)[:2]


class OracleIterative(QDSolver):
    def __init__(self, nlp, _zero=0.0, **kw):
        self.nlp = nlp
        self.kw = {k: v for k, v in kw.items() if k.split("_")[0] in ("ls", "ln", "ne")}
        self.o = None
        self.handle = None

    def _refresh(self, fpnlp, x):
        import scipy.sparse as sp
        r, c = fpnlp.nlp.jac_structure()
        A = sp.csr_matrix((fpnlp.nlp.jac_coord(x), (r, c)), shape=(fpnlp.nlp.meta.ncon, fpnlp.nlp.meta.nvar))
        self.o = O.IterativeOracle(A, **self.kw)

    def solve_two_mixed(self, fpnlp, x, rhs1, rhs2):
        self._refresh(fpnlp, x)
        return self.o.solve_two_mixed(fpnlp.delta, rhs1, rhs2)[:4]

    def solve_two_least_squares(self, fpnlp, x, rhs1, rhs2):
        return self.o.solve_two_least_squares(fpnlp.delta, rhs1, rhs2)[:4]

    def solve_two_extras(self, fpnlp, x, rhs1, rhs2):
        return self.o.solve_two_extras(fpnlp.delta, rhs1, rhs2)[:2]
