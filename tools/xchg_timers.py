"""Where the in-kernel peer exchange of the row-partitioned Krylov loop spends its time (debug build: csrc compiled with
-DFPSB_XCHG_TIMERS into variants/libfpsb200_xt.so; run with FPSB200_LIB pointing at it).  C3 operator, fixed 100 iterations.

    FPSB200_LIB=variants/libfpsb200_xt.so python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/xchg_timers.py --grid 2048
"""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

SEG = ["wait_local_grid", "scatter_puts+fence", "signal_round_trip_1", "boundary_rows", "local_sums+tot_puts+fence",
       "signal_round_trip_2", "gather_copy+totals+fences"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=100)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import fpsb200
    from fpsb200 import models, _lib
    from fpsb200.partition import RowPartition, DistHandle
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    N = args.grid
    A = models.poisson_control(N).A.tocsr()
    m, n = A.shape
    perm = np.empty(n, dtype=np.int64)
    perm[:m] = 2 * np.arange(m); perm[m:] = 2 * np.arange(m) + 1
    coo = A.tocoo()
    jr, jc, vals = coo.row.astype(np.int64), perm[coo.col], coo.data
    part = RowPartition(n, m, jr, jc, world)
    o = _lib.IterOpts()
    L = _lib.lib()
    L.fpsb_iter_default_opts(C.c_int64(n), C.c_int64(m), C.byref(o))
    o.ls_itmax = args.iters; o.ln_itmax = args.iters
    o.ls_atol = o.ls_rtol = 0.0; o.ln_atol = o.ln_rtol = o.ln_btol = 0.0; o.ln_conlim = 1e300
    rng = np.random.default_rng(1234)
    g1, g2 = rng.standard_normal(n), rng.standard_normal(m)
    D = DistHandle(part, rank, device=lr, dist=dist if world > 1 else None, opts=o, peer=True)
    D.set_jac_values(vals)
    loc = D.loc
    r1 = g1[loc.col0:loc.col0 + loc.n_own]; r2 = g2[loc.row0:loc.row0 + loc.m_loc]
    raw = np.zeros(32, dtype=np.uint64)
    for _ in range(2):
        D.solve_two_mixed(1e-2, r1, r2)
    torch.cuda.synchronize()
    L.fpsb_debug_xchg_timers(raw.ctypes.data_as(C.c_void_p))          # reset
    if world > 1:
        dist.barrier()
    D.solve_two_mixed(1e-2, r1, r2)
    ms, _ = D.H.iter_last_profile()
    torch.cuda.synchronize()
    L.fpsb_debug_xchg_timers(raw.ctypes.data_as(C.c_void_p))
    t = raw.reshape(2, 16).astype(np.float64)
    out = {"rank": rank, "world": world, "us_per_iteration": 1e3 * ms / args.iters}
    for o_, name in ((0, "after_n_space_phase"), (1, "after_m_space_phase")):
        cnt = max(t[o_, 15], 1.0)
        out[name] = {"exchanges": int(t[o_, 15]), **{SEG[i]: round(t[o_, i] / cnt / 1e3, 2) for i in range(7)},
                     "total_us": round(t[o_, :7].sum() / cnt / 1e3, 2), "cta1_wait_for_release_us": round(t[o_, 12] / cnt / 1e3, 2)}
    for r in range(world):
        if r == rank:
            print(json.dumps(out), flush=True)
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
