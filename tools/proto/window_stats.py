"""Design data for round 2 (CPU only): how the gather windows of 128-row tiles look on the BASELINE operators.

For every tile of `rows` consecutive rows of A and of A' it computes the sorted distinct columns the tile touches,
cuts them into contiguous segments (gap > `gap` entries starts a new segment) and reports
  * how many tiles fit ONE window of `cap` entries (what the round-1 step kernel needs for its shared-memory gathers),
  * how many would fit with up to 3 segments (one bulk copy per segment), and the window bytes per tile,
  * the share of entries a single best window of `cap` entries would cover (the rest = a short global-gather section).
    python tools/proto/window_stats.py [poisson N | random n m k w]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp
from fpsb200 import models


def tile_stats(A, rows=128, cap=1536, gap=8, max_tiles=4000):
    A = sp.csr_matrix(A)
    nt = (A.shape[0] + rows - 1) // rows
    pick = np.unique(np.linspace(0, nt - 1, min(nt, max_tiles)).astype(int))
    one = three = 0
    seg_counts, win_entries, cover, tile_entries = [], [], [], []
    for t in pick:
        lo, hi = A.indptr[t * rows], A.indptr[min((t + 1) * rows, A.shape[0])]
        cols = A.indices[lo:hi]
        if len(cols) == 0:
            continue
        u = np.unique(cols)
        cuts = np.flatnonzero(np.diff(u) > gap)
        starts = np.concatenate([[0], cuts + 1]); ends = np.concatenate([cuts, [len(u) - 1]])
        seg_len = u[ends] - u[starts] + 1
        span = u[-1] - u[0] + 1
        one += span <= cap
        # up to 3 segments: merge the smallest gaps until 3 remain
        if len(seg_len) > 3:
            gaps = u[starts[1:]] - u[ends[:-1]] - 1
            keep = np.sort(np.argsort(gaps)[-2:])
            bounds = np.concatenate([[0], keep + 1, [len(seg_len)]])
            seg3 = np.array([u[ends[bounds[i + 1] - 1]] - u[starts[bounds[i]]] + 1 for i in range(3)])
        else:
            seg3 = seg_len
        three += seg3.sum() <= cap
        seg_counts.append(len(seg_len)); win_entries.append(int(seg3.sum())); tile_entries.append(len(cols))
        # best single window of `cap` entries (sliding over the sorted columns)
        cs = np.sort(cols)
        j = np.searchsorted(cs, cs + cap, side="left")
        cover.append((j - np.arange(len(cs))).max() / len(cs))
    n = len(seg_counts)
    return {"tiles_sampled": n, "entries_per_tile": float(np.mean(tile_entries)), "one_window_fits": one / n,
            "three_segments_fit": three / n, "segments_median": float(np.median(seg_counts)),
            "window_entries_3seg_mean": float(np.mean(win_entries)), "window_bytes_3seg_mean": 16 * float(np.mean(win_entries)),
            "block_bytes_mean": 10 * float(np.mean(tile_entries)), "best_single_window_coverage": float(np.mean(cover))}


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "poisson"
    if kind == "poisson":
        N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
        A = models.poisson_control(N).A.tocsr()
        m, n = A.shape
        perm = np.empty(n, dtype=np.int64); perm[:m] = 2 * np.arange(m); perm[m:] = 2 * np.arange(m) + 1
        coo = A.tocoo()
        A = sp.csr_matrix((coo.data, (coo.row, perm[coo.col])), shape=(m, n))   # interleaved y_i, u_i (tools/dist_bench.py)
        name = f"poisson-control {N}x{N} (interleaved variables)"
    else:
        n, m, k, w = (int(a) for a in sys.argv[2:6]) if len(sys.argv) > 5 else (1_000_000, 500_000, 20, 64)
        A = models.window_random_jacobian(m, n, k, w=w, seed=1234)
        name = f"window-random n={n} m={m} {k}/row w={w}"
    for label, M in (("A  (m-space rows)", A), ("A' (n-space rows)", sp.csr_matrix(A.T))):
        s = tile_stats(M)
        print(name, "|", label, "|", " ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in s.items()))


if __name__ == "__main__":
    main()
