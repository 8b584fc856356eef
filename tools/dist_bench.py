"""Row-partitioned Krylov path (BASELINE config C3 shape): Poisson-constrained optimal control,
A = [L  -I] on an N x N grid with the state / control variables interleaved (banded Jacobian), rows
split in strips over the ranks.  Reports the time per Krylov iteration of a fixed-length
solve_two_mixed (LSQR + CRAIG in lock step) as the max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/dist_bench.py --grid 1024
"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--delta", type=float, default=1e-2)
    ap.add_argument("--both", action="store_true", help="time the NCCL transport next to the peer-memory one")
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
    qp = models.poisson_control(N)
    A = qp.A.tocsr()
    m, n = A.shape
    perm = np.empty(n, dtype=np.int64)            # interleave y_i, u_i -> banded Jacobian
    perm[:m] = 2 * np.arange(m); perm[m:] = 2 * np.arange(m) + 1
    coo = A.tocoo()
    jr, jc, vals = coo.row.astype(np.int64), perm[coo.col], coo.data
    part = RowPartition(n, m, jr, jc, world)
    o = _lib.IterOpts()
    _lib.lib().fpsb_iter_default_opts(C.c_int64(n), C.c_int64(m), C.byref(o))
    o.ls_itmax = args.iters; o.ln_itmax = args.iters
    rng = np.random.default_rng(1234)
    g1, g2 = rng.standard_normal(n), rng.standard_normal(m)
    for peer in ([True, False] if (args.both and world > 1) else [True]):
        D = DistHandle(part, rank, device=lr, dist=dist if world > 1 else None, opts=o, peer=peer)
        D.set_jac_values(vals)
        L = D.loc
        r1 = g1[L.col0:L.col0 + L.n_own]; r2 = g2[L.row0:L.row0 + L.m_loc]
        D.solve_two_mixed(args.delta, r1, r2)
        ts = []
        for _ in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            out = D.solve_two_mixed(args.delta, r1, r2)
            ms, nl = D.H.iter_last_profile()
            ts.append(ms)
        t = torch.tensor([min(ts)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            it = max(s["niter"] for s in out[4])
            print(json.dumps({"workload": f"poisson-control grid {N}x{N}: n={n} m={m} nnz={A.nnz}, row strips over {world} GPU(s)",
                              "n_gpus": world, "transport": "peer-memory" if D.peer else "nccl", "krylov_loop_ms": float(t.item()),
                              "iterations": it, "us_per_iteration": 1e3 * float(t.item()) / it,
                              "halo_entries_rank0": int(L.n_ext - L.n_own)}), flush=True)
        del D
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
