"""Times the three Krylov 2-RHS solves at the headline size (device-resident inputs) with the loop profile."""
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
d3 = torch.tensor(np.random.default_rng(7).standard_normal(n), device="cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
def run(name, f):
    for _ in range(3):
        out = f()
    H.timer_start()
    prof = []
    for _ in range(reps):
        out = f()
        prof.append(H.iter_last_profile())
    ms = H.timer_stop() / reps
    st = out[-1]
    print("%-28s %7.3f ms  iters %s  loop %.3f ms  half-iterations %s" % (name, ms, [s["niter"] for s in st],
          np.mean([p[0] for p in prof]), prof[-1][1]))
run("solve_two_mixed", lambda: H.iter_solve_two_mixed(0.0, d1, d2))
run("solve_two_least_squares", lambda: H.iter_solve_two_least_squares(0.0, d1, d3))
run("solve_two_extras", lambda: H.iter_solve_two_extras(0.0, d1, d2))
H.timer_start()
for _ in range(20):
    y = H.jprod(d1)
print("jprod  %.1f us" % (H.timer_stop() / 20 * 1e3))
H.timer_start()
for _ in range(20):
    y = H.jtprod(d2)
print("jtprod %.1f us" % (H.timer_stop() / 20 * 1e3))
