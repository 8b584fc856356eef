"""Debug: per-phase time line of the persistent Krylov loop kernel (library built with
-DFPSB_LOOP_TIMERS into variants/libfpsb200_lt.so; run with FPSB200_LIB pointing at it).
Stamps (globaltimer, ns) per phase and CTA: 7 phase entered, 0 first tile landed, 1/5/6 groups 0/1/2
finished their tiles, 2 all consumers done, 3 grid barrier passed, 4 recurrences done (phase open)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsb200, bench
from fpsb200 import _lib
n, m, k, w = 1_000_000, 500_000, 20, 64
A, jrow, jcol, vals, r1, r2 = bench.make_workload(n, m, k, w, 1234)
H = fpsb200.B200Handle(n, m, jrow, jcol)
H.iter_setup(None)
H.set_jac_values(torch.tensor(vals, device="cuda"))
d1 = torch.tensor(r1, device="cuda"); d2 = torch.tensor(r2, device="cuda")
L = _lib.lib()
raw = np.zeros(64 * 160 * 16 + 160 * 2 * 3 * 8, dtype=np.uint64)
for _ in range(2):
    out = H.iter_solve_two_mixed(0.0, d1, d2)
L.fpsb_debug_loop_timers(raw.ctypes.data_as(C.c_void_p))      # also resets the segment counters
out = H.iter_solve_two_mixed(0.0, d1, d2)
L.fpsb_debug_loop_timers(raw.ctypes.data_as(C.c_void_p))
buf = raw[:64 * 160 * 16].reshape(64, 160, 16)
seg = raw[64 * 160 * 16:].reshape(160, 2, 3, 8).astype(np.float64)
G = 148
t = buf[:, :G, :].astype(np.int64)
nph = int(os.environ.get("FPSB_LOOP_CHUNK", "24")) * 2
nph = min(nph, 64)
# the stamps belong to the LAST launch that ran (the chunk in which the solve converged): use phases with data
valid = [ph for ph in range(1, nph - 1) if t[ph, :, 3].min() > 0 and t[ph + 1, :, 3].min() > 0 and t[ph, 0, 3] > t[ph - 1, 0, 3]]
print("phases with stamps:", len(valid), "resolution check (distinct last digits):", len(set((t[valid[0], :, 3] % 1000).tolist())))
def stat(x):
    return "mean %7.2f  min %7.2f  max %7.2f" % (x.mean() / 1e3, x.min() / 1e3, x.max() / 1e3)
for par, name in ((0, "even phases"), (1, "odd phases")):
    ph = [p for p in valid if p % 2 == par]
    if not ph:
        continue
    ph = np.array(ph)
    P = t[ph]                      # [p][cta][8]
    Pp = t[ph - 1]
    period = (P[:, :, 3] - Pp[:, :, 3])
    print(f"--- {name} ({len(ph)}): phase period (barrier pass to barrier pass) {stat(period)} us")
    print("  prev pass -> prev open (recurrences)      ", stat(Pp[:, :, 4] - Pp[:, :, 3]))
    if Pp[:, :, 8].min() > 0:
        print("   boundary warp 0: enter -> records in (pass)  ", stat(Pp[:, :, 3] - Pp[:, :, 14]))
        print("   boundary warp 1: enter -> sums ready          ", stat(Pp[:, :, 11] - Pp[:, :, 15]))
        print("   warp 0: pass -> sums ready                    ", stat(Pp[:, :, 8] - Pp[:, :, 3]))
        print("   warp 0: recurrence / coefficients / wait warp 1 + publish", stat(Pp[:, :, 9] - Pp[:, :, 8]), "|", stat(Pp[:, :, 10] - Pp[:, :, 9]), "|", stat(Pp[:, :, 4] - Pp[:, :, 10]))
        print("   warp 1: recurrence / coefficients             ", stat(Pp[:, :, 12] - Pp[:, :, 11]), "|", stat(Pp[:, :, 13] - Pp[:, :, 12]))
        print("   all consumers done -> warp 0 enters boundary  ", stat(P[:, :, 14] - P[:, :, 2]))
    print("  prev pass -> first tile landed            ", stat(P[:, :, 0] - Pp[:, :, 3]))
    print("  first tile landed -> group0 tiles done    ", stat(P[:, :, 1] - P[:, :, 0]))
    print("  prev pass -> group0 / 1 / 2 tiles done    ", stat(P[:, :, 1] - Pp[:, :, 3]), "|", stat(P[:, :, 5] - Pp[:, :, 3]), "|", stat(P[:, :, 6] - Pp[:, :, 3]))
    print("  prev pass -> all consumers done           ", stat(P[:, :, 2] - Pp[:, :, 3]))
    print("  all consumers done -> pass (barrier wait) ", stat(P[:, :, 3] - P[:, :, 2]))
    last = (P[:, :, 2].max(axis=1) - Pp[:, :, 3].min(axis=1))
    print("  slowest CTA done after earliest pass      ", stat(last))
    print("  barrier latency (pass - slowest arrive)   ", stat(P[:, :, 3].min(axis=1) - P[:, :, 2].max(axis=1)), "(first CTA through)")
    print("                                            ", stat(P[:, :, 3].max(axis=1) - P[:, :, 2].max(axis=1)), "(last CTA through)")
    d = (P[:, :, 2] - Pp[:, :, 3]).mean(axis=0)
    order = np.argsort(d)
    print("  per-CTA mean busy time: fastest", [(int(c), round(d[c] / 1e3, 2)) for c in order[:4]], "slowest", [(int(c), round(d[c] / 1e3, 2)) for c in order[-6:]])

names = ["top / operand issue", "wait full", "row sums (phase 1)", "group barrier", "wait coefficients", "epilogue (phase 2)"]
for par in (0, 1):
    S = seg[:G, par]                     # [cta][group][8]
    tiles = S[:, :, 6].sum()
    print(f"--- segments, {'even' if par == 0 else 'odd'} phases: {tiles:.0f} group-tiles; cycles per tile (thread 0 of each group), mean over CTAs and groups")
    tot = 0.0
    for i, nm in enumerate(names):
        c = S[:, :, i].sum() / max(tiles, 1)
        tot += c
        print("   %-22s %8.0f cycles" % (nm, c))
    print("   %-22s %8.0f cycles = %.2f us at 1.9 GHz" % ("total", tot, tot / 1900.0))
