"""Debug: event-timed fixed-iteration fused solve (60 iterations) for A/B experiments."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200, bench
from fpsb200 import _lib
scale = float(os.environ.get("FPSB_SCALE", "1"))
n, m, k, w = int(1_000_000 * scale), int(500_000 * scale), 20, 64
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, k, w, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol)
o = fpsb200.IterOpts()
_lib.lib().fpsb_iter_default_opts(C.c_int64(n), C.c_int64(m), C.byref(o))
o.ls_itmax = 60; o.ln_itmax = 60
H.iter_setup(o)
H.set_jac_values(torch.tensor(vals, device="cuda"))
d1 = torch.tensor(r1, device="cuda"); d2 = torch.tensor(r2, device="cuda")
for _ in range(3):
    out = H.iter_solve_two_mixed(0.0, d1, d2)
    ms, nl = H.iter_last_profile()
    print("loop %.3f ms, %d step launches, %.2f us per launch" % (ms, nl, 1e3 * ms / max(nl, 1)))
