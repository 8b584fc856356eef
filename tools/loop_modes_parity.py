"""Single-GPU solve_two_mixed on the headline operator at size n: saves the four outputs (run once per FPSB_LOOP mode, then
compare with --compare).  python tools/loop_modes_parity.py --size N --tag T ;  python tools/loop_modes_parity.py --compare A B"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=500_000)
ap.add_argument("--tag", default="x")
ap.add_argument("--compare", nargs=2)
a = ap.parse_args()
if a.compare:
    A, B = np.load(f"/tmp/lmp_{a.compare[0]}.npz"), np.load(f"/tmp/lmp_{a.compare[1]}.npz")
    rel = [float(np.linalg.norm(A[k] - B[k]) / np.linalg.norm(B[k])) for k in ("p1", "q1", "p2", "q2")]
    print(a.compare, "iters", A["it"].tolist(), B["it"].tolist(), "rel diff", ["%.2e" % r for r in rel])
    sys.exit(0)
import torch
import fpsb200, bench
n, m = a.size, a.size // 2
A, jr, jc, vals, _, _ = bench.make_workload(n, m, 20, 64, 1234)
rng = np.random.default_rng(1234)
g1, g2 = rng.standard_normal(n), rng.standard_normal(m)
H = fpsb200.B200Handle(n, m, jr, jc)
H.iter_setup(None)
H.set_jac_values(vals)
nrep = int(os.environ.get("LMP_REPS", "3"))
outs = [H.iter_solve_two_mixed(0.0, g1, g2) for _ in range(nrep)]
o = outs[-1]
rep = max(float(np.abs(outs[0][k] - outs[-1][k]).max()) for k in range(4))
print("per-solve max |diff| to the last solve:", ["%.1e" % max(float(np.abs(x[k] - o[k]).max()) for k in range(4)) for x in outs],
      "iters", [[x[4][0]["niter"], x[4][1]["niter"]] for x in outs])
np.savez(f"/tmp/lmp_{a.tag}.npz", p1=o[0], q1=o[1], p2=o[2], q2=o[3], it=np.array([o[4][0]["niter"], o[4][1]["niter"]]))
# true residuals of the two systems K [p; q] = rhs
At = A.T.tocsr()
r1a = np.linalg.norm(o[0] + At @ o[1] - g1) / np.linalg.norm(g1); r1b = np.linalg.norm(A @ o[0]) / np.linalg.norm(g1)
r2a = np.linalg.norm(o[2] + At @ o[3]) / np.linalg.norm(g2); r2b = np.linalg.norm(A @ o[2] - g2) / np.linalg.norm(g2)
print(a.tag, "n", n, "iters", [o[4][0]["niter"], o[4][1]["niter"]], "repeat diff %.1e" % rep, "residuals %.2e %.2e %.2e %.2e" % (r1a, r1b, r2a, r2b),
      {k: v for k, v in os.environ.items() if k.startswith("FPSB_")})
