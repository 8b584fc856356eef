"""Per-level completion times of the LDLt factorisation / solves (needs a -DFPSB_LDLT_TIMERS build:
FPSB200_LIB=variants/libfpsb200_ldt.so python tools/ldlt_levels.py)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200, bench
from fpsb200 import _lib
from fpsb200.symbolic import order_dissection
n, m, k, w = 1_000_000, 500_000, 20, 64
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, k, w, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol)
H.set_jac_values(torch.tensor(vals, device="cuda"))
d1 = torch.tensor(r1, device="cuda")
H.ldlt_analyze(order_dissection(n, m, jrow, jcol))
for _ in range(2):
    H.timer_start(); ok = H.ldlt_factorize(1.4901161193847656e-08); tf = H.timer_stop()
    H.timer_start(); out = H.ldlt_solve_two_least_squares(d1, d1); ts = H.timer_stop()
    print("factorize %.2f ms  solve-only %.2f ms  ok=%s" % (tf, ts, ok))
lib = _lib.lib()
buf = (C.c_double * 256)()
for ph, name in enumerate(("factor", "forward", "backward")):
    nl = lib.fpsb_debug_ldlt_level_times(H.h, ph, buf, 256)
    if nl <= 0:
        print("no timers in this build"); break
    t = np.array(buf[:nl])
    print(name, "level completion (us):", " ".join("%d:%.0f" % (i, t[i]) for i in range(1, nl) if i % 4 == 1 or i >= 56))
