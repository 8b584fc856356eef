"""Micro-benchmark used for ncu captures: full-size (n=1M, m=500K, nnz=10M) fused 2-RHS Krylov solve."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200, bench
n, m, k, w = 1_000_000, 500_000, 20, 64
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, k, w, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol)
H.iter_setup(None)
H.set_jac_values(torch.tensor(vals, device="cuda"))
d1 = torch.tensor(r1, device="cuda"); d2 = torch.tensor(r2, device="cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    out = H.iter_solve_two_mixed(0.0, d1, d2)
    info = H.profile_info() if hasattr(H, "profile_info") else None
    print("solve iters %s" % [s["niter"] for s in out[4]], info)
