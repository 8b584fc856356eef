"""Round-2 prototype (CPU, numpy): tiles whose gather window is up to three contiguous SEGMENTS of the source vector,
concatenated in shared memory, with 16-bit indices relative to the concatenation — the format the step kernel needs for
stencil operators (BASELINE config C3), where one 1 536-entry window never fits (tools/proto/window_stats.py).

build(A) cuts the rows into tiles of up to 128 rows, finds per tile the <= 3 segments (largest gaps split), shrinks the
tile when the concatenated window exceeds the capacity, and stores (segment starts, segment lengths (even), remapped
u16 column indices).  emulate(tiles, x) performs the product the way the kernel would — copy the segments into a
window buffer, gather with the remapped indices — and is checked against A @ x.  This is the executable specification
of the host-side builder; nothing here is on the product path.
    python tools/proto/multiseg_tiles.py [poisson N]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp

CAP = 1536          # window capacity (entries) — kWinCapMax of the step kernel
MAXSEG = 3
TILE_ROWS = 128


def segments_of(cols):
    """<= MAXSEG (start, length) pairs covering the sorted distinct columns, starts even, lengths even."""
    u = np.unique(cols)
    if len(u) == 0:
        return []
    gaps = np.diff(u)
    cut = np.sort(np.argsort(gaps)[-(MAXSEG - 1):]) if len(u) > 1 else np.array([], dtype=int)
    cut = [c for c in cut if gaps[c] > 8]                      # tiny gaps are not worth a bulk copy
    bounds = [0] + [c + 1 for c in cut] + [len(u)]
    segs = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        start = int(u[a]) & ~1                                  # 16-byte aligned source for vectors of doubles
        length = int(u[b - 1]) - start + 1
        segs.append((start, length + (length & 1)))
    return segs


def build(A):
    A = sp.csr_matrix(A)
    A.sort_indices()
    tiles, r0 = [], 0
    while r0 < A.shape[0]:
        R = min(TILE_ROWS, A.shape[0] - r0)
        while True:
            lo, hi = A.indptr[r0], A.indptr[r0 + R]
            segs = segments_of(A.indices[lo:hi])
            if sum(l for _, l in segs) <= CAP or R <= 32:
                break
            R = max(32, ((R // 2) + 31) & ~31)
        windowed = sum(l for _, l in segs) <= CAP
        idx = A.indices[lo:hi].astype(np.int64)
        if windowed and segs:
            starts = np.array([s for s, _ in segs]); lens = np.array([l for _, l in segs])
            offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
            k = np.searchsorted(starts, idx, side="right") - 1     # segment of every entry
            assert np.all(idx - starts[k] < lens[k])
            rel = (offs[k] + idx - starts[k]).astype(np.uint16)
        else:
            rel = idx                                              # global-gather tile (32-bit indices)
        tiles.append(dict(row0=r0, nrows=R, segs=segs if windowed else [], rel=rel, vals=A.data[lo:hi],
                          rp=A.indptr[r0:r0 + R + 1] - lo))
        r0 += R
    return tiles


def emulate(tiles, x, nrows):
    y = np.zeros(nrows)
    for T in tiles:
        if T["segs"]:
            win = np.concatenate([np.pad(x[s:s + l], (0, max(0, s + l - len(x)))) for s, l in T["segs"]])   # one bulk copy per segment
            g = win[T["rel"]]
        else:
            g = x[T["rel"]]
        prod = T["vals"] * g
        y[T["row0"]:T["row0"] + T["nrows"]] = np.add.reduceat(np.concatenate([prod, [0.0]]), T["rp"][:-1]) * (np.diff(T["rp"]) > 0)
    return y


def main():
    from fpsb200 import models
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    A = models.poisson_control(N).A.tocsr()
    m, n = A.shape
    perm = np.empty(n, dtype=np.int64); perm[:m] = 2 * np.arange(m); perm[m:] = 2 * np.arange(m) + 1
    coo = A.tocoo()
    A = sp.csr_matrix((coo.data, (coo.row, perm[coo.col])), shape=(m, n))
    rng = np.random.default_rng(0)
    for label, M in (("A", A), ("A'", sp.csr_matrix(A.T))):
        tiles = build(M)
        x = rng.standard_normal(M.shape[1])
        err = np.abs(emulate(tiles, x, M.shape[0]) - M @ x).max()
        w = [T for T in tiles if T["segs"]]
        print(f"{label}: {len(tiles)} tiles, {len(w)} windowed ({100 * len(w) / len(tiles):.1f} %), "
              f"segments per tile {np.mean([len(T['segs']) for T in w]):.2f}, window entries {np.mean([sum(l for _, l in T['segs']) for T in w]):.0f}, "
              f"max |y - A x| = {err:.2e}")
        assert err < 1e-9


if __name__ == "__main__":
    main()
