"""Test-only QDSolver subtypes backed by the CPU oracle (lives under tests/, never shipped)."""
import numpy as np

from fpsb200.qdsolver import QDSolver, _jac_values, _npen, _structure
from oracle import oracle as O


class OracleLDLt(QDSolver):
    def __init__(self, nlp, _zero=0.0, P=None, explicit_linear_constraints=False, **kw):
        self.explicit_linear_constraints = explicit_linear_constraints
        self.rows, self.cols = _structure(nlp, explicit_linear_constraints)
        n, m = nlp.meta.nvar, _npen(nlp, explicit_linear_constraints)
        kw = {k: v for k, v in kw.items() if k in ("ldlt_tol", "ldlt_r1", "ldlt_r2")}   # like the reference's kwargs...
        self.o = O.LDLtOracle(n, m, self.rows, self.cols, np.arange(n + m) if P is None else P, **kw)
        self.handle = None

    def solve_two_mixed(self, fpnlp, x, rhs1, rhs2):
        return self.o.solve_two_mixed(_jac_values(fpnlp, x), fpnlp.delta, rhs1, rhs2)[:4]

    def solve_two_least_squares(self, fpnlp, x, rhs1, rhs2):
        return self.o.solve_two_least_squares(rhs1, rhs2)[:4]

    def solve_two_extras(self, fpnlp, x, rhs1, rhs2):
        self.o.jvals = np.asarray(_jac_values(fpnlp, x), dtype=np.float64)
        return self.o.solve_two_extras(fpnlp.delta, rhs1, rhs2)[:2]


class OracleIterative(QDSolver):
    def __init__(self, nlp, _zero=0.0, explicit_linear_constraints=False, **kw):
        self.explicit_linear_constraints = explicit_linear_constraints
        self.nlp = nlp
        self.kw = {k: v for k, v in kw.items() if k.split("_")[0] in ("ls", "ln", "ne")}
        self.o = None
        self.handle = None

    def _refresh(self, fpnlp, x):
        import scipy.sparse as sp
        r, c = _structure(fpnlp.nlp, fpnlp.explicit_linear_constraints)
        A = sp.csr_matrix((_jac_values(fpnlp, x), (r, c)),
                          shape=(_npen(fpnlp.nlp, fpnlp.explicit_linear_constraints), fpnlp.nlp.meta.nvar))
        self.o = O.IterativeOracle(A, **self.kw)

    def solve_two_mixed(self, fpnlp, x, rhs1, rhs2):
        self._refresh(fpnlp, x)
        return self.o.solve_two_mixed(fpnlp.delta, rhs1, rhs2)[:4]

    def solve_two_least_squares(self, fpnlp, x, rhs1, rhs2):
        return self.o.solve_two_least_squares(fpnlp.delta, rhs1, rhs2)[:4]

    def solve_two_extras(self, fpnlp, x, rhs1, rhs2):
        return self.o.solve_two_extras(fpnlp.delta, rhs1, rhs2)[:2]
