"""Row-block partition of the Jacobian for the multi-GPU Krylov path (BASELINE config C3:
"Krylov (MINRES/LSQR) path row-partitioned over 8xB200").

Rank r owns the constraint rows [row_bounds[r], row_bounds[r+1]) and the variables
[col_bounds[r], col_bounds[r+1]).  Its local operator A_loc has the owned rows and the "extended"
columns  ext_r = owned variables ∪ halo (variables of other ranks its rows touch), in GLOBAL column
order — so owned columns form one contiguous block and the halo columns of every peer another one.

  jprod   needs the halo values     : owner p sends x[halo_r ∩ owned_p] to r   (gather)
  jtprod  produces halo partial sums: r sends them to the owner p, which adds  (scatter-add)

Both exchanges use the same index sets in opposite directions.  This module is pure numpy (host
logic, tested on CPU with gloo); the device side is fpsb_dist_* in include/fpsb.h.
"""
import ctypes as C

import numpy as np


def balanced_bounds(n, world):
    """world + 1 boundaries splitting range(n) into nearly equal contiguous blocks."""
    return np.array([(n * r) // world for r in range(world + 1)], dtype=np.int64)


class LocalPart:
    """What one rank needs: local COO in extended numbering + the exchange pattern."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class RowPartition:
    def __init__(self, nvar, ncon, jrow, jcol, world, row_bounds=None, col_bounds=None):
        """jrow / jcol: 0-based global COO structure (jac_structure!)."""
        self.nvar, self.ncon, self.world = int(nvar), int(ncon), int(world)
        self.jrow = np.asarray(jrow, dtype=np.int64)
        self.jcol = np.asarray(jcol, dtype=np.int64)
        self.row_bounds = balanced_bounds(ncon, world) if row_bounds is None else np.asarray(row_bounds, dtype=np.int64)
        if col_bounds is None:
            col_bounds = self._col_bounds_from_rows()
        self.col_bounds = np.asarray(col_bounds, dtype=np.int64)
        assert len(self.row_bounds) == world + 1 and len(self.col_bounds) == world + 1
        assert self.row_bounds[0] == 0 and self.row_bounds[-1] == ncon
        assert self.col_bounds[0] == 0 and self.col_bounds[-1] == nvar
        self.row_owner = np.searchsorted(self.row_bounds, self.jrow, side="right") - 1
        # halo[r]: sorted global columns touched by the rows of r but owned elsewhere
        self._ext = []
        for r in range(world):
            cols = np.unique(self.jcol[self.row_owner == r])
            own = np.arange(self.col_bounds[r], self.col_bounds[r + 1], dtype=np.int64)
            self._ext.append(np.union1d(cols, own))

    def _col_bounds_from_rows(self):
        """Variable ownership follows the rows: the boundary between rank r-1 and r is the median first
        column of the rows around the row boundary (keeps the halos small for banded Jacobians)."""
        b = np.zeros(self.world + 1, dtype=np.int64)
        b[-1] = self.nvar
        if len(self.jrow) == 0:
            return balanced_bounds(self.nvar, self.world)
        first = np.full(self.ncon, self.nvar, dtype=np.int64)
        np.minimum.at(first, self.jrow, self.jcol)
        for r in range(1, self.world):
            lo, hi = max(self.row_bounds[r] - 8, 0), min(self.row_bounds[r] + 8, self.ncon)
            f = first[lo:hi]
            f = f[f < self.nvar]
            b[r] = int(np.median(f)) if len(f) else (self.nvar * r) // self.world
        b = np.maximum.accumulate(b)
        return b

    def ext_cols(self, r):
        return self._ext[r]

    def owned_cols(self, r):
        return np.arange(self.col_bounds[r], self.col_bounds[r + 1], dtype=np.int64)

    def local(self, r):
        W = self.world
        ext = self._ext[r]
        c0, c1 = self.col_bounds[r], self.col_bounds[r + 1]
        own_off = int(np.searchsorted(ext, c0))
        n_own = int(c1 - c0)
        sel = np.nonzero(self.row_owner == r)[0]
        jrow_loc = self.jrow[sel] - self.row_bounds[r]
        jcol_ext = np.searchsorted(ext, self.jcol[sel])
        owner_of_ext = np.searchsorted(self.col_bounds, ext, side="right") - 1
        recv_start = np.zeros(W, dtype=np.int64)
        recv_cnt = np.zeros(W, dtype=np.int64)
        for p in range(W):
            if p == r:
                continue
            idx = np.nonzero(owner_of_ext == p)[0]
            if len(idx):
                assert idx[-1] - idx[0] + 1 == len(idx)      # contiguous: ext is sorted by global column
                recv_start[p], recv_cnt[p] = idx[0], len(idx)
        send_ptr = np.zeros(W + 1, dtype=np.int64)
        send = []
        for p in range(W):
            if p != r:
                ep = self._ext[p]
                mine = ep[(ep >= c0) & (ep < c1)]             # columns I own that p keeps as halo
                send.append(own_off + (mine - c0))
            else:
                send.append(np.zeros(0, dtype=np.int64))
            send_ptr[p + 1] = send_ptr[p] + len(send[-1])
        send_idx = np.concatenate(send).astype(np.int64) if send else np.zeros(0, dtype=np.int64)
        return LocalPart(rank=r, n_ext=len(ext), m_loc=int(self.row_bounds[r + 1] - self.row_bounds[r]), ext=ext,
                         own_off=own_off, n_own=n_own, coo_sel=sel, jrow_loc=jrow_loc.astype(np.int64),
                         jcol_ext=jcol_ext.astype(np.int64), recv_start=recv_start, recv_cnt=recv_cnt,
                         send_ptr=send_ptr, send_idx=send_idx, row0=int(self.row_bounds[r]), col0=int(c0))


class DistHandle:
    """One rank of the row-partitioned Krylov solver (needs a GPU; NCCL is bound by libfpsb200)."""

    def __init__(self, part, rank, device=0, dist=None, opts=None, peer=True):
        from . import _lib
        from .qdsolver import B200Handle
        try:
            # libfpsb200 binds NCCL with dlopen("libnccl.so.2"): when PyTorch lives in the same process
            # its bundled NCCL must be the one that gets loaded (same SONAME, newer symbols)
            import torch  # noqa: F401
        except ImportError:
            pass
        self.part, self.rank, self.world = part, rank, part.world
        self.loc = L = part.local(rank)
        self.H = B200Handle(L.n_ext, L.m_loc, L.jrow_loc, L.jcol_ext, device=device)
        lib = _lib.lib()
        o = _lib.IterOpts()
        _lib.check(lib.fpsb_iter_default_opts(C.c_int64(part.nvar), C.c_int64(part.ncon), C.byref(o)), "fpsb_iter_default_opts")
        if opts is not None:
            o = opts
        self.H.iter_setup(o)
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(lib.fpsb_dist_unique_id(ident), "fpsb_dist_unique_id")
        if self.world > 1:
            box = [bytes(ident)]
            dist.broadcast_object_list(box, src=0)
            ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        p64 = lambda a: np.ascontiguousarray(a, dtype=np.int64).ctypes.data_as(C.POINTER(C.c_int64))
        self._keep = [np.ascontiguousarray(a, dtype=np.int64) for a in (L.recv_start, L.recv_cnt, L.send_ptr, L.send_idx)]
        _lib.check(lib.fpsb_dist_attach(self.H.h, C.c_int(self.world), C.c_int(rank), ident, C.c_int64(L.own_off),
                                        C.c_int64(L.n_own), p64(self._keep[0]), p64(self._keep[1]), p64(self._keep[2]),
                                        p64(self._keep[3]) if len(L.send_idx) else None), "fpsb_dist_attach")
        # peer-memory transport: all-gather the mailbox descriptors, map the peers (NCCL stays the fallback)
        self.peer = False
        if peer and self.world <= 8:
            lib.fpsb_dist_peer_blob_bytes.restype = C.c_int64
            nb = int(lib.fpsb_dist_peer_blob_bytes())
            blob = (C.c_ubyte * nb)()
            _lib.check(lib.fpsb_dist_peer_export(self.H.h, blob), "fpsb_dist_peer_export")
            blobs = [bytes(blob)]
            if self.world > 1:
                blobs = [None] * self.world
                dist.all_gather_object(blobs, bytes(blob))
            allb = (C.c_ubyte * (nb * self.world)).from_buffer_copy(b"".join(blobs))
            _lib.check(lib.fpsb_dist_peer_attach(self.H.h, allb), "fpsb_dist_peer_attach")
            self.peer = bool(lib.fpsb_dist_peer_active(self.H.h))

    def set_jac_values(self, vals_global):
        """vals_global: jac_coord values in the GLOBAL COO order (numpy); this rank keeps its rows."""
        self.H.set_jac_values(np.ascontiguousarray(np.asarray(vals_global, dtype=np.float64)[self.loc.coo_sel]))

    def _call2(self, fn, name, x, nout):
        from . import _lib
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(nout)
        _lib.check(fn(self.H.h, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int(_lib.FPSB_HOST)), name)
        return y

    def jprod(self, x_own):
        from . import _lib
        return self._call2(_lib.lib().fpsb_dist_jprod, "fpsb_dist_jprod", x_own, self.loc.m_loc)

    def jtprod(self, u_loc):
        from . import _lib
        return self._call2(_lib.lib().fpsb_dist_jtprod, "fpsb_dist_jtprod", u_loc, self.loc.n_own)

    def _solve(self, fn, name, delta, rhs1, rhs2):
        from . import _lib
        n, m = self.loc.n_own, self.loc.m_loc
        rhs1 = np.ascontiguousarray(rhs1, dtype=np.float64)
        rhs2 = np.ascontiguousarray(rhs2, dtype=np.float64)
        outs = [np.empty(n), np.empty(m), np.empty(n), np.empty(m)]
        st = (_lib.KrylovStats * 2)()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(fn(self.H.h, C.c_double(delta), C.c_int64(self.part.nvar), C.c_int64(self.part.ncon), vp(rhs1), vp(rhs2),
                      vp(outs[0]), vp(outs[1]), vp(outs[2]), vp(outs[3]), C.c_int(_lib.FPSB_HOST), st), name)
        return outs[0], outs[1], outs[2], outs[3], [st[0].as_dict(), st[1].as_dict()]

    def solve_two_mixed(self, delta, rhs1_own, rhs2_loc):
        from . import _lib
        return self._solve(_lib.lib().fpsb_dist_solve_two_mixed, "fpsb_dist_solve_two_mixed", delta, rhs1_own, rhs2_loc)

    def solve_two_extras(self, delta, rhs1_own, rhs2_loc):
        """(u1, u2, stats): u1 = LSQR(A', rhs1, sqrt(tau)), u2 = MINRES(A A' + tau I, rhs2) on the local m-space rows
        (src/solve_linear_system.jl:45-77)."""
        from . import _lib
        m = self.loc.m_loc
        rhs1 = np.ascontiguousarray(rhs1_own, dtype=np.float64)
        rhs2 = np.ascontiguousarray(rhs2_loc, dtype=np.float64)
        u1, u2 = np.empty(m), np.empty(m)
        st = (_lib.KrylovStats * 2)()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(_lib.lib().fpsb_dist_solve_two_extras(self.H.h, C.c_double(delta), C.c_int64(self.part.nvar), C.c_int64(self.part.ncon),
                                                         vp(rhs1), vp(rhs2), vp(u1), vp(u2), C.c_int(_lib.FPSB_HOST), st),
                   "fpsb_dist_solve_two_extras")
        return u1, u2, [st[0].as_dict(), st[1].as_dict()]

    def solve_two_least_squares(self, delta, rhs1_own, rhs2_own):
        from . import _lib
        return self._solve(_lib.lib().fpsb_dist_solve_two_least_squares, "fpsb_dist_solve_two_least_squares", delta,
                           rhs1_own, rhs2_own)
