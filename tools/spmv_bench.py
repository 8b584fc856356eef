"""Micro-benchmark used for ncu captures: full-size (n=1M, m=500K, nnz=10M) SpMV + one fused solve."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200, bench
n, m, k, w = 1_000_000, 500_000, 20, 64
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, k, w, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol)
H.iter_setup(None)
H.set_jac_values(torch.tensor(vals, device="cuda"))
d1 = torch.tensor(r1, device="cuda"); d2 = torch.tensor(r2, device="cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    H.jprod(d1); H.jtprod(d2)
H.timer_start()
for _ in range(reps):
    H.jprod(d1)
t1 = H.timer_stop() / reps
H.timer_start()
for _ in range(reps):
    H.jtprod(d2)
t2 = H.timer_stop() / reps
out = H.iter_solve_two_mixed(0.0, d1, d2)
print("spmv A %.1f us  At %.1f us  solve iters %s" % (1e3 * t1, 1e3 * t2, [s["niter"] for s in out[4]]))
