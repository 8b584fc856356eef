"""BASELINE configs beside the headline one, one JSON line each (1 GPU):
  C2  synthetic sparse equality QP n=1e5, m=5e4, 10 nnz/row: LDLt path (analyze, refactor + 2-RHS solve, solve-only),
      the same operations by the CPU oracle, and a whole `fps_solve` (host buffers and device-resident);
  C5  4096 small dense instances through fpsb_batch_solve_two (host buffers).
    python tools/configs_bench.py [--skip-cpu]
"""
import argparse, json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    import fpsb200
    from fpsb200 import models
    warnings.simplefilter("ignore")
    n, m = 100_000, 50_000
    qp = models.sparse_qp(n, m, nnz_per_row=10, w=64, seed=1234)
    rng = np.random.default_rng(1234)
    r1, r2 = rng.standard_normal(n), rng.standard_normal(m)
    coo = qp.A.tocoo()
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    for ordering in ("amd", "dissection"):
        t = time.time(); S = fpsb200.LDLtSolver(qp, 0.0, ordering=ordering); t_an = time.time() - t
        H = S.handle
        H.set_jac_values(qp.jac_coord(None))
        d1, d2 = torch.tensor(r1, device="cuda"), torch.tensor(r2, device="cuda")
        delta = float(np.sqrt(np.finfo(float).eps))
        best_f = best_s = 1e30
        for _ in range(5):
            H.timer_start(); out = H.ldlt_solve_two_mixed(delta, d1, d2); best_f = min(best_f, H.timer_stop())
            H.timer_start(); out2 = H.ldlt_solve_two_least_squares(d1, d1); best_s = min(best_s, H.timer_stop())
        p1, q1, p2, q2 = (o.cpu().numpy() for o in out[:4])
        A = qp.A
        res = max(rel(p1 + A.T @ q1, r1), float(np.linalg.norm(A @ p1 - delta * q1) / np.linalg.norm(r1)),
                  float(np.linalg.norm(p2 + A.T @ q2) / np.linalg.norm(r2)), rel(A @ p2 - delta * q2, r2))
        info = H.ldlt_plan_info()
        line = {"config": "C2 LDLt", "n": n, "m": m, "nnz": int(A.nnz), "ordering": ordering, "analyze_host_s": round(t_an, 3),
                "refactor_plus_solve_ms": round(best_f, 3), "solve_only_ms": round(best_s, 3), "factorized": bool(out[4]),
                "k_residual_rel": res, "panel_nnz": int(info.get("panel_nnz", 0)), "nsuper": int(info.get("nsuper", 0)), "flops": float(info.get("flops", 0))}
        if not args.skip_cpu:
            from oracle import oracle as O
            P = H.ldlt_symbolic()["P"]
            lo = O.LDLtOracle(n, m, coo.row, coo.col, P)
            t = time.time(); ref = lo.solve_two_mixed(coo.data, delta, r1, r2); line["cpu_oracle_refactor_plus_solve_ms"] = round(1e3 * (time.time() - t), 1)
            t = time.time(); lo.solve_two_least_squares(r1, r1); line["cpu_oracle_solve_only_ms"] = round(1e3 * (time.time() - t), 1)
            line["max_rel_diff_vs_oracle"] = max(rel(a, b) for a, b in zip((p1, q1, p2, q2), ref[:4]))
        print(json.dumps(line), flush=True)
        del S, H
    # whole fps_solve on the C2 problem
    F = __import__("importlib").import_module("fpsb200.fps_solve")
    tight = dict(ls_atol=1e-11, ls_rtol=1e-11, ln_atol=1e-11, ln_rtol=1e-11, ln_btol=1e-11)
    for solver, kw in (("ldlt", dict(ordering="dissection")), ("iterative", {}), ("iterative", tight)):
        t = time.time(); st = F.fps_solve(qp, qds_solver=solver, **kw); t_host = time.time() - t
        dqp = fpsb200.DeviceSparseQP(qp)
        t = time.time()
        sd = F.fps_solve(dqp, torch.zeros(n, dtype=torch.float64, device="cuda"), qds_solver=solver,
                         model_factory=fpsb200.DeviceFletcherPenaltyNLP, **kw)
        t_dev = time.time() - t
        sub = sd.solver_specific["solver"].sub_stats
        print(json.dumps({"config": "C2 fps_solve", "qds_solver": solver, "krylov_tolerances": "1e-11" if kw is tight else "reference defaults (sqrt eps)", "status": st.status, "iter": st.iter,
                          "host_buffers_s": round(t_host, 3), "device_resident_s": round(t_dev, 3), "status_device": sd.status,
                          "iter_device": sd.iter, "primal_feas": st.primal_feas, "dual_feas": st.dual_feas,
                          "objective": st.objective, "last_subproblem": {"tr_iterations": sub.iter, "hprods": sub.cg_iter},
                          "note": "wall time including the host-side analysis / tile building of the constructor"}), flush=True)
    # C5: 4096 small dense instances
    ninst, nn, mm = 4096, 10, 3
    Ab = rng.standard_normal((ninst, mm, nn)); b1 = rng.standard_normal((ninst, nn)); b2 = rng.standard_normal((ninst, mm))
    fpsb200.batch_solve_two(Ab, 0.0, b1, b2)
    ts = []
    for _ in range(10):
        t = time.perf_counter(); out = fpsb200.batch_solve_two(Ab, 0.0, b1, b2); ts.append(time.perf_counter() - t)
    print(json.dumps({"config": "C5 batch", "instances": ninst, "n": nn, "m": mm, "ms_host_buffers": round(1e3 * min(ts), 3),
                      "instances_per_s": round(ninst / min(ts)), "all_factorized": bool(out[4].all())}), flush=True)


if __name__ == "__main__":
    main()
