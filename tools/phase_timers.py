"""Debug: per-phase clock64 totals of the step kernel (library built with -DFPSB_PHASE_TIMERS)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200, bench
from fpsb200 import _lib
n, m, k, w = 1_000_000, 500_000, 20, 64
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, k, w, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol)
o = fpsb200.IterOpts()
_lib.lib().fpsb_iter_default_opts(C.c_int64(n), C.c_int64(m), C.byref(o))
o.ls_itmax = 60; o.ln_itmax = 60        # fixed work per solve, also for experiments that break the numerics
H.iter_setup(o)
H.set_jac_values(torch.tensor(vals, device="cuda"))
d1 = torch.tensor(r1, device="cuda"); d2 = torch.tensor(r2, device="cuda")
out = H.iter_solve_two_mixed(0.0, d1, d2)
L = _lib.lib()
buf = (C.c_ulonglong * 16)()
L.fpsb_debug_phase_timers(buf, 1)
l0 = H.launch_count()
opts = fpsb200.IterOpts() if False else None
out = H.iter_solve_two_mixed(0.0, d1, d2)
L.fpsb_debug_phase_timers(buf, 0)
launches = H.launch_count() - l0 - 6
v = np.array(list(buf), dtype=float)
names = ["prodB:meta", "prodB:wait_empty", "prodB:issue", "prodW:meta", "prodW:wait_empty", "prodW:issue",
         "cons:loop_top", "cons:operand+meta issue", "cons:wait_full", "cons:phase1", "cons:group_bar", "cons:phase2"]
print("step launches ~", launches, "iters", [st["niter"] for st in out[4]])
for nm, x in zip(names, v):
    print("%-26s %10.2f us per launch per CTA (at 1.9 GHz)  = %8.0f cycles per tile" % (nm, x / 148 / launches / 1900.0, x / 148 / launches / 13.2))
