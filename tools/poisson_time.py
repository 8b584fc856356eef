"""Single-GPU fused Krylov path on the Poisson-control operator (BASELINE config C3 shape, interleaved
variables): time per iteration of a fixed-length solve_two_mixed."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200
from fpsb200 import models, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
qp = models.poisson_control(N)
A = qp.A.tocsr(); m, n = A.shape
perm = np.empty(n, dtype=np.int64); perm[:m] = 2 * np.arange(m); perm[m:] = 2 * np.arange(m) + 1
coo = A.tocoo()
H = fpsb200.B200Handle(n, m, coo.row.astype(np.int64), perm[coo.col])
o = _lib.IterOpts(); _lib.lib().fpsb_iter_default_opts(C.c_int64(n), C.c_int64(m), C.byref(o)); o.ls_itmax = 100; o.ln_itmax = 100
o.ls_atol = o.ls_rtol = 0.0; o.ln_atol = o.ln_rtol = o.ln_btol = 0.0; o.ln_conlim = 1e300     # both methods run every iteration (as bench.py's partitioned record)
H.iter_setup(o)
H.set_jac_values(coo.data)
rng = np.random.default_rng(1234)
d1 = torch.tensor(rng.standard_normal(n), device="cuda"); d2 = torch.tensor(rng.standard_normal(m), device="cuda")
for _ in range(3):
    out = H.iter_solve_two_mixed(1e-2, d1, d2)
    ms, nl = H.iter_last_profile()
print("tiles", H.tile_stats(), "FPSB_NO_SEGS" in os.environ)
print("grid %d: n=%d m=%d nnz=%d: loop %.2f ms, %d iterations -> %.1f us per iteration (fused single-GPU path)" % (N, n, m, A.nnz, ms, max(s["niter"] for s in out[4]), 1e3 * ms / max(s["niter"] for s in out[4])))
