"""Per-call device time of iter_solve_two_least_squares at the headline size, before / after an NVML sampling episode
(diagnostic for bench.py's `extra.iter_solve_two_least_squares`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import fpsb200
n, m, k, w = 1_000_000, 500_000, 20, 64
A, jrow, jcol, vals, rhs1, rhs2 = bench.make_workload(n, m, k, w, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol, device=0)
H.iter_setup(None)
dev = torch.device("cuda", 0)
d_vals = torch.tensor(vals, device=dev); d_r1 = torch.tensor(rhs1, device=dev); d_r2 = torch.tensor(rhs2, device=dev)
d_r3 = torch.tensor(np.random.default_rng(7).standard_normal(n), device=dev)
H.set_jac_values(d_vals)
def series(tag, cnt=8):
    ts = []
    for _ in range(cnt):
        t0 = time.perf_counter()
        H.timer_start(); o = H.iter_solve_two_least_squares(0.0, d_r1, d_r3); ms = H.timer_stop()
        ts.append((round(ms, 2), round(1e3 * (time.perf_counter() - t0), 2), round(H.iter_last_profile()[0], 2)))
    print(tag, "(device ms, wall ms, loop ms):", ts, flush=True)
for _ in range(3):
    H.iter_solve_two_mixed(0.0, d_r1, d_r2)
series("plain")
s = bench.make_clock_sampler(torch, 0)
s.start(); time.sleep(0.2); print(s.stop())
series("after nvml sampler")
H.timer_start()
for _ in range(20):
    y = H.jprod(d_r1)
print("jprod", H.timer_stop() / 20)
series("after jprod")
for _ in range(3):
    H.iter_solve_two_mixed(0.0, d_r1, d_r2)
series("after mixed")
