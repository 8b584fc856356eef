"""Where does a resident bench step go? wall / event times of set_jac_values and the solve."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200, bench
n, m, k, w = 1_000_000, 500_000, 20, 64
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, k, w, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol)
H.iter_setup(None)
dv = torch.tensor(vals, device="cuda"); d1 = torch.tensor(r1, device="cuda"); d2 = torch.tensor(r2, device="cuda")
for _ in range(5):
    H.set_jac_values(dv); H.iter_solve_two_mixed(0.0, d1, d2)
ts = {"set_jac": 0.0, "solve": 0.0}; loop = 0.0; N = 20
torch.cuda.synchronize(); t00 = time.perf_counter()
for _ in range(N):
    t0 = time.perf_counter(); H.set_jac_values(dv); t1 = time.perf_counter(); out = H.iter_solve_two_mixed(0.0, d1, d2); t2 = time.perf_counter()
    ts["set_jac"] += t1 - t0; ts["solve"] += t2 - t1; loop += H.iter_last_profile()[0]
tot = time.perf_counter() - t00
print("per step: total %.3f ms | set_jac_values %.3f ms | solve call %.3f ms (Krylov loop on device %.3f ms)" % (1e3 * tot / N, 1e3 * ts["set_jac"] / N, 1e3 * ts["solve"] / N, loop / N))
