import time, sys, numpy as np
sys.path.insert(0,'.')
import bench
from fpsb200.symbolic import SymbolicAnalysis, order_dissection
from oracle import oracle as O
n, m = 1_000_000, 500_000
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, 20, 64, 1234)
for name in ("amd", "dissection"):
    P = order_dissection(n, m, jrow, jcol) if name == "dissection" else SymbolicAnalysis(n, m, jrow, jcol).get()["P"]
    t = time.time(); lo = O.LDLtOracle(n, m, jrow, jcol, P); ta = time.time() - t
    t = time.time(); out = lo.solve_two_mixed(vals, 1.4901161193847656e-08, r1, r2); tf = time.time() - t
    t = time.time(); out2 = lo.solve_two_least_squares(r1, r1); ts = time.time() - t
    print(name, "oracle analyze %.2fs  refactor+solve %.2fs  solve-only %.3fs ok=%s" % (ta, tf, ts, out[4]), flush=True)
