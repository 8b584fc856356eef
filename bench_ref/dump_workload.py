"""Dump the headline workload of bench.py (same generator, same seed) as raw little-endian files so that the
TRUE reference (Julia) can be timed on the same inputs:  python bench_ref/dump_workload.py out_dir [n m k w seed]
Files: meta.txt (n m nnz), jrow.i64 / jcol.i64 (1-based COO, jac_structure! order), jval.f64, rhs1.f64, rhs2.f64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

out = sys.argv[1]
n, m, k, w, seed = (int(a) for a in sys.argv[2:7]) if len(sys.argv) >= 7 else (1_000_000, 500_000, 20, 64, 1234)
A, jrow, jcol, vals, rhs1, rhs2 = bench.make_workload(n, m, k, w, seed)
os.makedirs(out, exist_ok=True)
open(os.path.join(out, "meta.txt"), "w").write(f"{n} {m} {len(vals)}\n")
(jrow + 1).astype("<i8").tofile(os.path.join(out, "jrow.i64"))
(jcol + 1).astype("<i8").tofile(os.path.join(out, "jcol.i64"))
vals.astype("<f8").tofile(os.path.join(out, "jval.f64"))
rhs1.astype("<f8").tofile(os.path.join(out, "rhs1.f64"))
rhs2.astype("<f8").tofile(os.path.join(out, "rhs2.f64"))
print("wrote", out)
