"""Debug: one solve_two_mixed through the persistent loop kernel at a given size; prints status / timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200, bench
n = int(sys.argv[1]); m = n // 2
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, 20, 64, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol)
H.iter_setup(None)
H.set_jac_values(torch.tensor(vals, device="cuda"))
d1 = torch.tensor(r1, device="cuda"); d2 = torch.tensor(r2, device="cuda")
t = time.perf_counter()
try:
    out = H.iter_solve_two_mixed(0.0, d1, d2)
    print("n", n, "ok iters", [s["niter"] for s in out[4]], "%.3f s" % (time.perf_counter() - t), H.iter_last_profile())
except Exception as e:
    print("n", n, "FAILED after %.3f s:" % (time.perf_counter() - t), str(e)[-120:])
