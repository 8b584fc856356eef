"""CPU: the host logic of the row-partitioned Krylov path (fpsb200.partition) — the halo gather and
scatter-add patterns reproduce A v and A' u exactly, both emulated in-process for several world
sizes and across two real processes over gloo (world_size 2), which is how the N>1 path is wired."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from fpsb200 import models
from fpsb200.partition import RowPartition, balanced_bounds


def _problem(m=300, n=600, k=7, w=24, seed=3, banded=True):
    if banded:
        A = models.window_random_jacobian(m, n, k, w=w, seed=seed).tocsr()
    else:
        A = sp.random(m, n, density=0.02, random_state=seed, format="csr")
        A.data[:] = np.random.default_rng(seed).standard_normal(A.nnz)
    coo = A.tocoo()
    return A, coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data


def _emulate(part, vals, x, u):
    """Run both exchanges with plain numpy using only what each rank would hold."""
    W = part.world
    L = [part.local(r) for r in range(W)]
    Aloc = [sp.coo_matrix((vals[l.coo_sel], (l.jrow_loc, l.jcol_ext)), shape=(l.m_loc, l.n_ext)).tocsr() for l in L]
    # gather: every rank fills its halo slots from the owners' send lists
    xe = []
    for r, l in enumerate(L):
        v = np.zeros(l.n_ext)
        v[l.own_off:l.own_off + l.n_own] = x[l.col0:l.col0 + l.n_own]
        xe.append(v)
    for r, l in enumerate(L):
        for p in range(W):
            if p == r or l.recv_cnt[p] == 0:
                continue
            lp = L[p]
            sent = xe[p][lp.send_idx[lp.send_ptr[r]:lp.send_ptr[r + 1]]]
            assert len(sent) == l.recv_cnt[p]
            xe[r][l.recv_start[p]:l.recv_start[p] + l.recv_cnt[p]] = sent
    y = np.concatenate([Aloc[r] @ xe[r] for r in range(W)])
    # scatter-add: halo partial sums go back to the owners
    S = [Aloc[r].T @ u[l.row0:l.row0 + l.m_loc] for r, l in enumerate(L)]
    z = []
    for r, l in enumerate(L):
        own = S[r][l.own_off:l.own_off + l.n_own].copy()
        for p in range(W):
            if p == r:
                continue
            lp = L[p]
            got = S[p][lp.recv_start[r]:lp.recv_start[r] + lp.recv_cnt[r]]
            idx = l.send_idx[l.send_ptr[p]:l.send_ptr[p + 1]]
            assert len(got) == len(idx)
            own[idx - l.own_off] += got
        z.append(own)
    return y, np.concatenate(z)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("banded", [True, False])
def test_halo_patterns_reproduce_products(world, banded):
    A, jr, jc, vals = _problem(banded=banded)
    m, n = A.shape
    part = RowPartition(n, m, jr, jc, world)
    rng = np.random.default_rng(0)
    x, u = rng.standard_normal(n), rng.standard_normal(m)
    y, z = _emulate(part, vals, x, u)
    assert np.allclose(y, A @ x, rtol=1e-13, atol=1e-13)
    assert np.allclose(z, A.T @ u, rtol=1e-13, atol=1e-13)
    for r in range(world):
        l = part.local(r)
        assert np.array_equal(l.ext[l.own_off:l.own_off + l.n_own], part.owned_cols(r))
        assert np.all(np.diff(l.ext) > 0)
        if banded and world > 1:
            assert l.n_ext - l.n_own < 200          # halos stay thin for banded Jacobians


def test_partition_edge_cases():
    # empty Jacobian, more ranks than rows, explicit bounds
    part = RowPartition(5, 3, [], [], 2)
    for r in range(2):
        l = part.local(r)
        assert l.n_ext == l.n_own and l.send_ptr[-1] == 0
    part = RowPartition(4, 2, [0, 1], [0, 3], 4)
    assert sum(part.local(r).m_loc for r in range(4)) == 2
    assert sum(part.local(r).n_own for r in range(4)) == 4
    b = balanced_bounds(10, 3)
    assert b[0] == 0 and b[-1] == 10 and np.all(np.diff(b) >= 3)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, jr, jc, vals = _problem()
    m, n = A.shape
    part = RowPartition(n, m, jr, jc, world)
    l = part.local(rank)
    Aloc = sp.coo_matrix((vals[l.coo_sel], (l.jrow_loc, l.jcol_ext)), shape=(l.m_loc, l.n_ext)).tocsr()
    rng = np.random.default_rng(0)
    x, u = rng.standard_normal(n), rng.standard_normal(m)
    # gather (the ncclSend / ncclRecv pattern of fpsb_dist.inl, over gloo)
    xe = np.zeros(l.n_ext)
    xe[l.own_off:l.own_off + l.n_own] = x[l.col0:l.col0 + l.n_own]
    reqs, bufs = [], {}
    for p in range(world):
        if p == rank:
            continue
        ns = l.send_ptr[p + 1] - l.send_ptr[p]
        if ns:
            reqs.append(dist.isend(torch.from_numpy(xe[l.send_idx[l.send_ptr[p]:l.send_ptr[p + 1]]].copy()), p))
        if l.recv_cnt[p]:
            bufs[p] = torch.empty(int(l.recv_cnt[p]), dtype=torch.float64)
            reqs.append(dist.irecv(bufs[p], p))
    for rq in reqs:
        rq.wait()
    for p, b in bufs.items():
        xe[l.recv_start[p]:l.recv_start[p] + l.recv_cnt[p]] = b.numpy()
    y = Aloc @ xe
    # scatter-add
    S = Aloc.T @ u[l.row0:l.row0 + l.m_loc]
    reqs, bufs = [], {}
    for p in range(world):
        if p == rank:
            continue
        if l.recv_cnt[p]:
            reqs.append(dist.isend(torch.from_numpy(S[l.recv_start[p]:l.recv_start[p] + l.recv_cnt[p]].copy()), p))
        ns = l.send_ptr[p + 1] - l.send_ptr[p]
        if ns:
            bufs[p] = torch.empty(int(ns), dtype=torch.float64)
            reqs.append(dist.irecv(bufs[p], p))
    for rq in reqs:
        rq.wait()
    z = S[l.own_off:l.own_off + l.n_own].copy()
    for p in sorted(bufs):
        z[l.send_idx[l.send_ptr[p]:l.send_ptr[p + 1]] - l.own_off] += bufs[p].numpy()
    # inner product across ranks = the all-reduce of the Krylov norms
    t = torch.tensor([float(y @ y), float(z @ z)], dtype=torch.float64)
    dist.all_reduce(t)
    ok_y = np.allclose(y, (A @ x)[l.row0:l.row0 + l.m_loc], rtol=1e-13, atol=1e-13)
    ok_z = np.allclose(z, (A.T @ u)[l.col0:l.col0 + l.n_own], rtol=1e-13, atol=1e-13)
    q.put((rank, bool(ok_y), bool(ok_z), t.tolist(), [float((A @ x) @ (A @ x)), float((A.T @ u) @ (A.T @ u))]))
    dist.barrier()
    dist.destroy_process_group()


def test_row_partition_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] and r[2]
        assert np.allclose(r[3], r[4], rtol=1e-12)


def test_multisegment_window_prototype():
    """tools/proto/multiseg_tiles.py (round-2 design prototype, CPU only): tiles of the Poisson-control operator with up to
    three window segments and 16-bit concatenation-relative indices reproduce A x and A' x."""
    import importlib.util
    import scipy.sparse as sp
    from fpsb200 import models
    spec = importlib.util.spec_from_file_location("multiseg_tiles", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "proto", "multiseg_tiles.py"))
    ms = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ms)
    A = models.poisson_control(48).A.tocsr()
    rng = np.random.default_rng(1)
    for M in (A, sp.csr_matrix(A.T)):
        tiles = ms.build(M)
        assert all(len(T["segs"]) <= 3 and sum(l for _, l in T["segs"]) <= ms.CAP for T in tiles)
        assert sum(T["nrows"] for T in tiles) == M.shape[0]
        x = rng.standard_normal(M.shape[1])
        assert np.abs(ms.emulate(tiles, x, M.shape[0]) - M @ x).max() < 1e-9 * np.abs(M.data).max()
