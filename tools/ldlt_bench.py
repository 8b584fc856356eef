"""Micro-benchmark for launch lists / ncu: full-size LDLt refactor + 2-RHS solve (dissection ordering)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200, bench
from fpsb200.symbolic import order_dissection
n, m, k, w = 1_000_000, 500_000, 20, 64
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, k, w, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol)
H.set_jac_values(torch.tensor(vals, device="cuda"))
d1 = torch.tensor(r1, device="cuda"); d2 = torch.tensor(r2, device="cuda")
H.ldlt_analyze(order_dissection(n, m, jrow, jcol))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    H.timer_start(); ok = H.ldlt_factorize(1.4901161193847656e-08); tf = H.timer_stop()
    H.timer_start(); out = H.ldlt_solve_two_least_squares(d1, d1); ts = H.timer_stop()
    print("factorize %.2f ms  solve-only %.2f ms  ok=%s" % (tf, ts, ok))
