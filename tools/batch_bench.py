"""Throughput mode (BASELINE config C5): 4096 independent small instances (n <= 10, m <= 3) through
fpsb_batch_solve_two; instances/s with host buffers (the ABI call) on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fpsb200
rng = np.random.default_rng(0)
for ninst, n, m in ((4096, 10, 3), (4096, 2, 1), (65536, 10, 3)):
    A = rng.standard_normal((ninst, m, n)); r1 = rng.standard_normal((ninst, n)); r2 = rng.standard_normal((ninst, m))
    fpsb200.batch_solve_two(A, 1e-3, r1, r2)
    t0 = time.perf_counter(); reps = 20
    for _ in range(reps):
        out = fpsb200.batch_solve_two(A, 1e-3, r1, r2)
    dt = (time.perf_counter() - t0) / reps
    res = np.abs(out[0] + np.einsum("imn,im->in", A, out[1]) - r1).max()
    print("ninst %6d n %2d m %d: %.3f ms per batch, %.2f M instances/s (host buffers), max residual %.1e" % (ninst, n, m, 1e3 * dt, ninst / dt / 1e6, res))
