"""BASELINE config C4: rank-deficient Jacobian stress at n = 1e6, m = 5e5, nnz = 1e7 (1 % of the rows are copies of
other rows), delta = 0 -> LDLt refactorisation with the reference's dynamic regularisation (r2 = -sqrt(eps) on the
pivots of the lower-right block that fall below tol), 2-RHS solve, and the Krylov path on the same operator."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200
from fpsb200 import models
from fpsb200.symbolic import order_dissection

n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1_000_000, 500_000)
qp = models.sparse_qp(n, m, nnz_per_row=20, w=64, seed=1234, rank_deficient_frac=0.01)
A = qp.A
coo = A.tocoo()
rng = np.random.default_rng(7)
r1, r2 = rng.standard_normal(n), A @ rng.standard_normal(n)        # rhs2 in the range of A: a consistent system
H = fpsb200.B200Handle(n, m, coo.row.astype(np.int64), coo.col.astype(np.int64))
H.set_jac_values(torch.tensor(coo.data, device="cuda"))
t = time.time(); P = order_dissection(n, m, coo.row.astype(np.int64), coo.col.astype(np.int64)); H.ldlt_analyze(P); t_an = time.time() - t
d1, d2 = torch.tensor(r1, device="cuda"), torch.tensor(r2, device="cuda")
best = 1e30
for _ in range(3):
    H.timer_start(); out = H.ldlt_solve_two_mixed(0.0, d1, d2); best = min(best, H.timer_stop())
p1, q1, p2, q2 = (o.cpu().numpy() for o in out[:4])
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
L, D = H.ldlt_get_factor()[:2] if hasattr(H, "ldlt_get_factor") else (None, None)
se = float(np.sqrt(np.finfo(float).eps))
line = {"config": "C4 rank-deficient LDLt", "n": n, "m": m, "nnz": int(A.nnz), "duplicated_rows": int(0.01 * m), "delta": 0.0,
        "analysis_host_s": round(t_an, 2), "refactor_plus_solve_ms": round(best, 2), "factorized": bool(out[4]),
        "res1_rel": rel(p1 + A.T @ q1, r1), "Ap1_over_rhs1": float(np.linalg.norm(A @ p1) / np.linalg.norm(r1)),
        "res2_rel": rel(A @ p2, r2), "p2_plus_Atq2_over_rhs2": float(np.linalg.norm(p2 + A.T @ q2) / np.linalg.norm(r2)),
        "q_norms": [float(np.linalg.norm(q1)), float(np.linalg.norm(q2))]}
if D is not None:
    D = np.asarray(D)
    line["pivots_regularised_to_minus_sqrt_eps"] = int(np.sum(D == -se))
    line["pivots_regularised_to_plus_sqrt_eps"] = int(np.sum(D == se))
print(json.dumps(line), flush=True)
H.iter_setup(None)
for _ in range(2):
    o2 = H.iter_solve_two_mixed(0.0, d1, d2)
    ms, nl = H.iter_last_profile()
P1, Q1, P2, Q2 = (o.cpu().numpy() for o in o2[:4])
print(json.dumps({"config": "C4 rank-deficient Krylov", "krylov_loop_ms": round(ms, 2), "step_launches": int(nl),
                  "stats": [{k: s[k] for k in ("niter", "solved", "status")} for s in o2[4]],
                  "res1_rel": rel(P1 + A.T @ Q1, r1), "Ap1_over_rhs1": float(np.linalg.norm(A @ P1) / np.linalg.norm(r1)),
                  "res2_rel": rel(A @ P2, r2), "p1_vs_ldlt_rel": rel(P1, p1), "p2_vs_ldlt_rel": rel(P2, p2)}), flush=True)
