"""Debug: event-timed solve_two_least_squares vs solve_two_mixed at full size."""
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
for name, fn in (("mixed", lambda: H.iter_solve_two_mixed(0.0, d1, d2)), ("lsq", lambda: H.iter_solve_two_least_squares(0.0, d1, d3))):
    for rep in range(4):
        t0 = time.perf_counter(); H.timer_start(); out = fn(); ms = H.timer_stop(); wall = 1e3 * (time.perf_counter() - t0)
        lms, nl = H.iter_last_profile()
        print(name, "event %.3f ms wall %.3f ms loop %.3f ms launches %d iters %s" % (ms, wall, lms, nl, [s["niter"] for s in out[4]]))

# what bench.py does between the headline measurement and this extra: asynchronous products, a second handle that comes and goes
def lsq(tag):
    for rep in range(3):
        H.timer_start(); out = H.iter_solve_two_least_squares(0.0, d1, d3); ms = H.timer_stop()
        lms, nl = H.iter_last_profile()
        print(tag, "lsq event %.3f ms loop %.3f ms launches %d iters %s" % (ms, lms, nl, [s["niter"] for s in out[4]]))
for _ in range(20):
    y = H.jprod(d1)
for _ in range(20):
    y = H.jtprod(d2)
lsq("after async products:")
H2 = fpsb200.B200Handle(n, m, jrow, jcol)
H2.iter_setup(None)
H2.set_jac_values(torch.tensor(vals, device="cuda"))
H2.iter_solve_two_mixed(0.0, d1, d2)
lsq("second handle alive:")
H2.close(); del H2
lsq("second handle closed:")
for name, fn in (("mixed", lambda: H.iter_solve_two_mixed(0.0, d1, d2)),):
    for rep in range(3):
        H.timer_start(); out = fn(); ms = H.timer_stop()
        print(name, "event %.3f ms" % ms)
