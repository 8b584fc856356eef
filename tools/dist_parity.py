"""Parity of the row-partitioned solve_two_mixed against the single-GPU solve on the headline (window-random) operator.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_parity.py [--size N] [--fixed K]
"""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=1_000_000)
    ap.add_argument("--fixed", type=int, default=0, help="fixed iteration count (tolerances 0)")
    ap.add_argument("--delta", type=float, default=0.0)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import fpsb200, bench
    from fpsb200 import _lib
    from fpsb200.partition import RowPartition, DistHandle
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    n, m = args.n, args.n // 2
    A, jr, jc, vals, _, _ = bench.make_workload(n, m, 20, 64, 1234)
    o = _lib.IterOpts()
    _lib.lib().fpsb_iter_default_opts(C.c_int64(n), C.c_int64(m), C.byref(o))
    if args.fixed > 0:
        o.ls_itmax = o.ln_itmax = args.fixed
        o.ls_atol = o.ls_rtol = 0.0
        o.ln_atol = o.ln_rtol = o.ln_btol = 0.0
        o.ln_conlim = 1e300
    rng = np.random.default_rng(1234)
    g1, g2 = rng.standard_normal(n), rng.standard_normal(m)
    part = RowPartition(n, m, jr, jc, world)
    D = DistHandle(part, rank, device=lr, dist=dist, opts=o, peer=True)
    D.set_jac_values(vals)
    L = D.loc
    r1, r2 = g1[L.col0:L.col0 + L.n_own], g2[L.row0:L.row0 + L.m_loc]
    outs = [D.solve_two_mixed(args.delta, r1, r2) for _ in range(3)]
    out = outs[-1]
    repeat = max(float(np.abs(outs[0][k] - outs[2][k]).max()) for k in range(4))
    full = [torch.empty(k, dtype=torch.float64, device=dev) for k in (n, m, n, m)]
    it1 = torch.zeros(2, dtype=torch.float64, device=dev)
    if rank == 0:
        H1 = fpsb200.B200Handle(n, m, jr, jc, device=lr)
        H1.iter_setup(o)
        H1.set_jac_values(vals)
        o1 = H1.iter_solve_two_mixed(args.delta, torch.tensor(g1, device=dev), torch.tensor(g2, device=dev))
        for k in range(4):
            full[k].copy_(o1[k])
        it1[0], it1[1] = o1[4][0]["niter"], o1[4][1]["niter"]
    dist.broadcast(it1, src=0)
    err = torch.zeros(8, dtype=torch.float64, device=dev)
    sl = [(L.col0, L.n_own), (L.row0, L.m_loc), (L.col0, L.n_own), (L.row0, L.m_loc)]
    for k in range(4):
        dist.broadcast(full[k], src=0)
        ref = full[k][sl[k][0]:sl[k][0] + sl[k][1]]
        mine = torch.tensor(out[k], device=dev)
        err[2 * k] = torch.sum((mine - ref) ** 2); err[2 * k + 1] = torch.sum(ref ** 2)
    dist.all_reduce(err, op=dist.ReduceOp.SUM)
    e = err.tolist()
    rel = [float(np.sqrt(e[2 * k] / max(e[2 * k + 1], 1e-300))) for k in range(4)]
    rp = torch.tensor([repeat], dtype=torch.float64, device=dev)
    dist.all_reduce(rp, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"tag": args.tag, "n": n, "world": world, "fixed": args.fixed, "iters": [out[4][0]["niter"], out[4][1]["niter"]],
                          "iters_single": [int(it1[0].item()), int(it1[1].item())], "rel_err": rel,
                          "max_abs_diff_between_repeated_solves": float(rp.item()),
                          "env": {k: v for k, v in os.environ.items() if k.startswith("FPSB_")}}), flush=True)
    del D
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
