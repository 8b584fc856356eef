// fpsb_loop.inl — the persistent Krylov loop kernel (included by fpsb_krylov.cu after gk_step_kernel).
//
// Reference surface: the iteration loops of Krylov.jl's lsqr / craig! behind solve_least_square /
// solve_least_norm (src/solve_two_systems_struct.jl:167-185, 210-244), as driven two at a time by
// solve_two_mixed / solve_two_least_squares (src/solve_linear_system.jl:79-140).
//
// gk_step_kernel pays ~11 us per half iteration that do not scale with the operator (launch, prologue,
// pipeline fill and drain) — a quarter of every launch at the headline size.  Here ONE launch runs a
// whole chunk of half iterations ("phases"): phase ph streams the tiles of operator (first + ph) & 1
// exactly like gk_step_kernel does, then the CTAs meet at a grid barrier (release/acquire on a
// monotonic counter in HBM), EVERY CTA reduces the per-CTA norm partials in the same fixed order and
// runs the two slots' scalar recurrences on its own shared-memory copy of the slot states (so the
// states stay bitwise identical across CTAs and every CTA takes the same decisions), and the next
// phase starts — without leaving the SM:
//   * the TMA ring (mbarriers, stage parities) runs straight through the phase boundary: tile
//     numbers keep counting, nothing is re-initialised;
//   * the tile BLOCKS of the next phase do not depend on the current one (matrix values), so the block
//     producers request the first `nspec` tiles of phase ph + 1 while phase ph still drains and the
//     grid barrier is pending; only the gather windows (the vector the previous phase wrote) wait;
//   * with `early`, the consumers already compute the row sums of their first tile while two of their
//     warps run the recurrences; only the row epilogue waits for the new coefficients;
//   * convergence is seen by all CTAs at the same phase; the speculatively requested tiles of the
//     phase that does not run any more are drained and the kernel ends.  CTA 0 publishes the states.
// Requirements (checked by the host): both operators have no long rows, one ring geometry fits both,
// grid <= tiles of either operator, all CTAs co-resident (cooperative launch, 1 CTA per SM).
// Row-partitioned runs (fpsb_dist.inl): what CTA 0 needs to do the exchange of a phase boundary itself, through the
// peers' mailboxes (CUDA IPC over NVLink), while the other CTAs wait for ITS release instead of the grid barrier:
//   n-space phase:  halo partial sums -> owners ; owners add them (rank order), run the row epilogue of their
//                   boundary rows (long boundaries: shared out over the consumer threads of a few helper CTAs, which
//                   are idle until the halo is back) and put the fresh pair values into the halo slots of the peers ;
//                   norm sums
//   m-space phase:  norm sums only
// The four all-reduced sums go to gtot; every CTA runs the scalar recurrences on those, exactly as on one GPU.
// One launch = a chunk of iterations of every rank: an iteration costs two NVLink signal round trips (n) + one (m),
// no kernel launch and no host involvement.
struct DistLoop {
    int nranks, rank;
    int nbound;                               // owned rows that peers contribute to (and keep as halo)
    const int4 *brow;                         // per boundary row {row, first entry, end entry, inbox position of the first entry}
    const int *bidx, *bptr, *bsrc, *bpeer;    // per boundary row: its send-list entries (rank order) and their peers
    const long long *meta;                    // recv_start[R] | recv_cnt[R] | send_ptr[R+1] | ga_off[R]
    double2 *S;                               // raw row sums of the halo / boundary rows (StepParams::raw_out of op[0])
    double2 *pair;                            // gathered pair of the extended n-space (Gn)
    long long nsend, nrecv;                   // my inbox sizes (entries)
    unsigned char *mine;
    unsigned char *peer[kMboxMaxRanks];
    long long sc_at_peer[kMboxMaxRanks], ga_at_peer[kMboxMaxRanks];
    long long peer_nsend[kMboxMaxRanks], peer_nrecv[kMboxMaxRanks];
    unsigned long long *seq;                  // device copy of [0] signals issued, [1] scatters, [2] gathers, [3] tot exchanges
    double *gtot;                             // [2][4] all-reduced sums, ping-pong by phase parity
    unsigned long long *xbar;                 // exchanges completed in this launch (zero at launch)
    // boundary rows shared out over the consumer threads of CTAs 0 .. nhelp-1 (one SM's L2 bandwidth, ~100 GB/s, was the
    // bound: 2.5 MB of row operands per exchange at an interior C3 strip = 36 of 67 us, tools/xchg_timers.py)
    int nhelp;
    unsigned long long *sbar;                 // n-space exchanges whose scatter inbox is complete (zero at launch)
    unsigned long long *hbar;                 // arrivals of the helper CTAs that finished their slice (zero at launch)
    double *bparts;                           // [nhelp][4] norm partials of the helpers' slices
    double2 *gastage;                         // fresh values of the boundary rows in send-list order (local): CTA 0 puts them to the peers
    int *err;
};
// mailbox layout (bytes):  flags u64[8] | tot double[2][8][4] | scatter inbox double2[2][nsend] | gather inbox double2[2][nrecv]
__host__ __device__ static inline size_t mbox_off_tot() { return kMboxOffTot; }
__host__ __device__ static inline size_t mbox_off_sc() { return mbox_off_tot() + 2 * kMboxMaxRanks * 4 * sizeof(double); }
__host__ __device__ static inline size_t mbox_off_ga(int64_t nsend) { return mbox_off_sc() + 2 * (size_t)nsend * sizeof(double2); }

struct LoopParams {
    StepParams op[2];            // [0] n-space step (rows of A'), [1] m-space step (rows of A); io modes fixed for the loop
    const DistLoop *dx;          // non-null: row-partitioned run, CTA 0 exchanges with the peers at every phase boundary
    int first;                   // operator of phase 0
    int nphase;                  // phases this launch may run at most
    int nspec;                   // tiles of a phase the block producers may request before the phase opens
    int early;                   // 1: row sums of the first tiles overlap the scalar recurrences
    SlotState *st;               // slot states in / out
    double *parts;               // [2][grid][4] per-CTA norm partials, ping-pong by phase parity
    unsigned long long *gbar;    // arrival counter of the grid barrier (zero at launch)
    int *done_flag;              // [0] done, [1] += phases run, [2] error
};

constexpr int kLoopMaxGrid = 192;      // CTAs of the persistent loop (records are polled 3 per lane of the two recurrence warps)

struct LoopCtl {
    int pass;        // grid barriers passed: phase ph may read what phase ph - 1 wrote once pass >= ph
    int open1;       // recurrence warp 1: coefficients of phases < open1 are ready
    int open;        // phases < open are decided (coefficients of both slots ready, stop_at valid)
    int stop_at;     // first phase that does not run
    int abort;
};

__device__ __forceinline__ void st_release_cta(int *p, int v) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta(const int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Wait until *p >= want (shared-memory counter published with st_release_cta).  ONE thread polls, rarely:
// a dozen warps spinning on shared memory flood the SM's memory-instruction queue, which is exactly what the
// recurrence warps' own shared-memory accesses and shuffles have to get through (measured: 5 us per boundary)
__device__ __forceinline__ void loop_wait_thread(const int *p, int want, const LoopCtl *ctl) {
    unsigned it = 0;
    while (*reinterpret_cast<const volatile int *>(p) < want) {
        __nanosleep(100);
        if ((++it & 1023) == 0 && *reinterpret_cast<const volatile int *>(&ctl->abort)) __trap();    // error path: end the launch, never hang
    }
    __threadfence_block();                                   // acquire side of st_release_cta
}
__device__ __forceinline__ void loop_wait(const int *p, int want, const LoopCtl *ctl) {     // whole warp
#ifdef FPSB_LOOP_TIGHTPOLL
    for (unsigned it = 0;; ++it) {
        if (ld_acquire_cta(p) >= want) return;
        if ((it & 15) == 15) { if (*reinterpret_cast<const volatile int *>(&ctl->abort)) __trap(); __nanosleep(32); }
    }
#else
    if ((threadIdx.x & 31) == 0) loop_wait_thread(p, want, ctl);
    __syncwarp();
#endif
}

// debug builds (-DFPSB_XCHG_TIMERS, tools/xchg_timers.py): globaltimer ns CTA 0's thread 0 spent in the segments of
// loop_exchange, summed per phase kind [o][segment]; [o][15] counts the exchanges; [o][12] is CTA 1's wait for the release
__device__ unsigned long long g_xchg_t[2][16];
#ifdef FPSB_XCHG_TIMERS
#define XT_DECL unsigned long long xt0 = global_ns()
#define XT_MARK(i) do { if (ct == 0) { const unsigned long long xt1 = global_ns(); g_xchg_t[o][(i)] += xt1 - xt0; xt0 = xt1; } } while (0)
#else
#define XT_DECL
#define XT_MARK(i)
#endif
#ifndef FPSB_BOUNDARY_UNROLL
#define FPSB_BOUNDARY_UNROLL 2
#endif
#ifdef FPSB_LOOP_TIMERS
// debug builds: per phase (first 64 of a launch) and CTA, globaltimer stamps of 8 points of the phase
__device__ unsigned long long g_loop_t[64][160][16];
#define LT_STAMP(ph, i) do { if ((ph) < 64 && cta < 160) g_loop_t[(ph)][cta][(i)] = global_ns(); } while (0)
// per CTA, phase parity and consumer group: clock64 cycles thread 0 of the group spent in 6 segments of its tiles
// (0 top / operand issue, 1 wait full, 2 row sums, 3 group barrier, 4 wait coefficients, 5 epilogue), [6] = tiles
__device__ unsigned long long g_loop_seg[160][2][kGroups][8];
// (accumulated in shared memory by thread 0 of each group, flushed once per phase: a global RMW per mark distorts)
#define LS_DECL long long ls_t0 = clock64()
#define LS_MARK(i) do { if (t == 0) { const long long ls_t1 = clock64(); s_seg[g][(i)] += (unsigned long long)(ls_t1 - ls_t0); ls_t0 = ls_t1; } } while (0)
#define LS_RESET do { ls_t0 = clock64(); } while (0)
#define LS_COUNT do { if (t == 0) s_seg[g][6] += 1; } while (0)
#define LS_FLUSH do { if (t == 0 && cta < 160) { for (int q_ = 0; q_ < 8; ++q_) { g_loop_seg[cta][ph & 1][g][q_] += s_seg[g][q_]; s_seg[g][q_] = 0; } } } while (0)
#else
#define LT_STAMP(ph, i)
#define LS_DECL
#define LS_MARK(i)
#define LS_RESET
#define LS_COUNT
#define LS_FLUSH
#endif

// (Tried and dropped: per-CTA records stamped in the sign bit so that arrival and partials are one L2 round trip:
//  the 148 x 64 polling lanes slowed the CTAs that were still streaming by more than the round trip saved.)

// dst[i] = src[i] (16-byte entries, dst possibly remote) by the CTA's consumer threads, 4 loads in flight per thread
// (a plain `dst[i] = src[i]` loop serialises an L2 round trip per entry: the remote store may alias the next load)
__device__ __forceinline__ void copy_batched(double2 *dst, const double2 *src, long long cnt, int ct) {
    constexpr int kCons = kGroups * kGroupThreads, kU = 4;
    for (long long i0 = ct; i0 < cnt; i0 += (long long)kCons * kU) {
        double2 v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) { const long long i = i0 + (long long)u * kCons; v[u] = i < cnt ? __ldcg(src + i) : make_double2(0.0, 0.0); }
#pragma unroll
        for (int u = 0; u < kU; ++u) { const long long i = i0 + (long long)u * kCons; if (i < cnt) dst[i] = v[u]; }
    }
}

// helper CTAs of this rank: none for a short boundary (the hand-shake costs more than 4 rows per thread of CTA 0:
// measured on the headline operator, 108 boundary rows, 79 -> 89 us per iteration with 16 helpers)
__device__ __forceinline__ int loop_helpers(const DistLoop &X, int gsz) {
    return X.nbound >= 4 * kGroups * kGroupThreads ? max(1, min(X.nhelp, gsz)) : 1;
}

// Boundary rows [lo, hi) by the CTA's 384 consumer threads: add the peers' partial sums (rank order), Krylov row epilogue,
// fresh pair value -> the row itself and the send-list slots of the peers that keep it as halo (gastage, local).
// kU rows per thread at a time; the packed record {row, first entry, end entry, inbox position of the first entry} of the
// NEXT batch is requested before the operands of the current one, the operands and the first inbox entry of all kU rows
// before the first use.
__device__ __noinline__ void boundary_rows_slice(const LoopParams &L, const DistLoop &X, const Coef *sC, int lo, int hi, int ct,
                                                 double (&acc)[4]) {
    constexpr int kCons = kGroups * kGroupThreads;
    constexpr int kU = FPSB_BOUNDARY_UNROLL;
    const CoefR C0 = to_regs(sC[0]), C1 = to_regs(sC[1]);
    const bool act0 = C0.mode != MD_NONE, act1 = C1.mode != MD_NONE;
    const StepParams &Pn = L.op[0];
    const bool peers = X.nranks > 1;
    const int4 *brow = X.brow;
    const double2 *S = X.S;
    double2 *pair = X.pair;
    double2 *stage = X.gastage;
    const double2 *inbox = reinterpret_cast<const double2 *>(X.mine + mbox_off_sc()) + (size_t)(__ldcg(X.seq + 1) & 1) * X.nsend;
    int4 ra[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        const int b = lo + ct + u * kCons;
        ra[u] = make_int4(-1, 0, 0, -1);
        if (b < hi) ra[u] = __ldg(brow + b);
    }
    for (int b0 = lo + ct; b0 < hi; b0 += kCons * kU) {
        int4 na[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int b = b0 + (kU + u) * kCons;
            na[u] = make_int4(-1, 0, 0, -1);
            if (b < hi) na[u] = __ldg(brow + b);
        }
        double2 sm[kU], old2[kU], in0[kU];
        double a00[kU], a01[kU], a10[kU], a11[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            sm[u] = old2[u] = in0[u] = make_double2(0.0, 0.0);
            a00[u] = a01[u] = a10[u] = a11[u] = 0.0;
            const int row = ra[u].x;
            if (peers && ra[u].w >= 0) in0[u] = __ldcg(inbox + ra[u].w);
            if (row >= 0) {
                sm[u] = __ldcg(S + row);
                old2[u] = __ldcg(pair + row);
                if (C0.rd0()) a00[u] = __ldcg(Pn.io[0].a0 + row);
                if (C0.rd1()) a01[u] = __ldcg(Pn.io[0].a1 + row);
                if (C1.rd0()) a10[u] = __ldcg(Pn.io[1].a0 + row);
                if (C1.rd1()) a11[u] = __ldcg(Pn.io[1].a1 + row);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int row = ra[u].x, kb = ra[u].y, ke = ra[u].z;
            if (row < 0) continue;
            const bool contrib = peers && ra[u].w >= 0;
            if (contrib) {
                // contributions in rank order: the first was requested above, more than one is rare (a row shared by 3 ranks)
                sm[u].x += in0[u].x; sm[u].y += in0[u].y;
                for (int k = kb + 1; k < ke; ++k) { const double2 a = __ldcg(inbox + X.bsrc[k]); sm[u].x += a.x; sm[u].y += a.y; }
            }
            double n0 = old2[u].x, n1 = old2[u].y;
            if (act0) n0 = row_epilogue(C0, sm[u].x, old2[u].x, a00[u], a01[u], acc[0], acc[1]);
            if (act1) n1 = row_epilogue(C1, sm[u].y, old2[u].y, a10[u], a11[u], acc[2], acc[3]);
            const double2 val = make_double2(n0, n1);
            pair[row] = val;
            if (C0.wr0()) Pn.io[0].a0[row] = a00[u];
            if (C0.wr1()) Pn.io[0].a1[row] = a01[u];
            if (C1.wr0()) Pn.io[1].a0[row] = a10[u];
            if (C1.wr1()) Pn.io[1].a1[row] = a11[u];
            if (contrib) {
                stage[ra[u].w] = val;
                for (int k = kb + 1; k < ke; ++k) stage[X.bsrc[k]] = val;
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) ra[u] = na[u];
    }
}

// CTAs 1 .. nhelp-1 after an n-space phase: their consumer threads (idle until the exchange is over: the next phase
// gathers the halo) take a slice of the boundary rows once CTA 0 says the scatter inbox is complete, leave the slice's norm
// partials in bparts and arrive at hbar.  No remote traffic here: the fresh values go to gastage, CTA 0 puts them to the peers.
__device__ __noinline__ void loop_help(const LoopParams &L, const DistLoop &X, int ph, int ct, int cw, int lane, int cta, int gsz,
                                       const Coef *sC, double *s_red /* 4*32 */, LoopCtl &ctl) {
    if (((L.first + ph) & 1) != 0 || X.nranks <= 1) return;
    const int K = loop_helpers(X, gsz);
    if (cta >= K) return;
    const unsigned long long nn = (unsigned long long)((ph + 2 - (L.first & 1)) >> 1);
    if (ct == 0) {
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while (ld_acquire_gpu(X.sbar) < nn) {
            __nanosleep(40);
            if ((++spins & 4095) == 0 && global_ns() - t0 > 6000000000ull) {
                ctl.abort = 1; atomicExch(L.done_flag + 2, 2); __threadfence_system(); __trap();
            }
        }
    }
    consumers_bar();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int chunk = (X.nbound + K - 1) / K;
    boundary_rows_slice(L, X, sC, min(cta * chunk, X.nbound), min((cta + 1) * chunk, X.nbound), ct, acc);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double x = warp_sum(acc[q]);
        if (lane == 0) s_red[q * 32 + cw] = x;
    }
    fence_proxy_async();                                     // this slice's generic writes of the pair vs the next phase's TMA reads
    consumers_bar();                                         // every thread's stores of the slice precede the arrival below
    if (cw == 0) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = warp_sum(lane < kGroups * kGroupWarps ? s_red[q * 32 + lane] : 0.0);
        if (lane == 0) {
            double *pp = X.bparts + (size_t)cta * 4;
            pp[0] = v[0]; pp[1] = v[1]; pp[2] = v[2]; pp[3] = v[3];
            __threadfence();
            red_release_gpu(X.hbar, 1ull);
        }
    }
    consumers_bar();                                         // s_red is free again
}

// What loop_exchange needs of a peer for ONE exchange, gathered once into shared memory (every field of DistLoop lives in
// global memory: a dependent L2 round trip per use — measured at 8 GPUs: 36 of 67 us of the n-space exchange went into the
// per-row chains bpeer -> peer -> inbox sizes -> store address, tools/xchg_timers.py)
struct XPeer {
    double2 *sc_dst;                  // my slice of the peer's scatter inbox (this exchange's half)
    const double2 *sc_src;            // the raw sums of the peer's columns: S + recv_start[p]
    double2 *ga_dst;                  // the peer's gather inbox (this exchange's half); my entries start at ga_at_peer[p]
    const double2 *ga_src;            // the peer's slice of my gather inbox (this exchange's half)
    double2 *ga_copy_dst;             // my halo slots of the peer's columns: pair + recv_start[p]
    double *tot_dst;                  // my slot of the peer's tot inbox (this exchange's half)
    long long cnt;                    // halo entries shared with the peer (0: not a neighbour)
    long long send_cnt;               // my boundary values the peer keeps as halo
    double2 *ga_put_dst;              // my slice of the peer's gather inbox (this exchange's half)
    const double2 *ga_put_src;        // gastage + send_ptr[p]
    unsigned long long *flag_dst;     // my flag in the peer's mailbox
    const unsigned long long *flag_src;   // the peer's flag in mine
};

// The exchange of the boundary after phase `ph`, by the 384 consumer threads of CTA 0 (see DistLoop).
__device__ __noinline__ void loop_exchange(const LoopParams &L, const DistLoop &X, int ph, int ct, int cw, int lane, int gsz,
                                           const Coef *sC, double *s_red /* 4*32 */, double *s_x /* >= 8 */, LoopCtl &ctl) {
    constexpr int kCons = kGroups * kGroupThreads;
    __shared__ XPeer s_peer[kMboxMaxRanks];
    __shared__ const double *s_tot_in;
    const int R = X.nranks;
    const int o = (L.first + ph) & 1;                      // 0: n-space phase (rows of A_loc'), 1: m-space phase
    auto die = [&]() { ctl.abort = 1; atomicExch(X.err, 1); atomicExch(L.done_flag + 2, 2); __threadfence_system(); __trap(); };
    XT_DECL;
    // every CTA of this GPU has left its record and its raw sums
    if (ct == 0) {
        const unsigned long long want = (unsigned long long)(ph + 1) * (unsigned long long)gsz;
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while (ld_acquire_gpu(L.gbar) < want) {
            __nanosleep(20);
            if ((++spins & 4095) == 0 && global_ns() - t0 > 4000000000ull) die();
        }
    }
    // meanwhile consumer warp 1 gathers the peers' descriptions of this exchange (the sequence numbers were left by the
    // previous exchange, everything else is constant)
    if (ct >= 32 && ct < 32 + R) {
        const int p = ct - 32;
        const long long *recv_start = X.meta, *recv_cnt = X.meta + R, *ga_off = X.meta + 3 * R + 1;
        const int sc_par = (int)(__ldcg(X.seq + 1) & 1), ga_par = (int)(__ldcg(X.seq + 2) & 1), tot_par = (int)(__ldcg(X.seq + 3) & 1);
        XPeer P;
        unsigned char *pm = X.peer[p];
        const long long pns = X.peer_nsend[p], pnr = X.peer_nrecv[p];
        P.cnt = p == X.rank ? 0 : recv_cnt[p];
        P.sc_dst = reinterpret_cast<double2 *>(pm + mbox_off_sc()) + (size_t)sc_par * pns + X.sc_at_peer[p];
        P.sc_src = X.S + recv_start[p];
        P.ga_dst = reinterpret_cast<double2 *>(pm + mbox_off_ga(pns)) + (size_t)ga_par * pnr;
        {
            const long long *send_ptr = X.meta + 2 * R;
            P.send_cnt = p == X.rank ? 0 : send_ptr[p + 1] - send_ptr[p];
            P.ga_put_dst = P.ga_dst + X.ga_at_peer[p];
            P.ga_put_src = X.gastage + send_ptr[p];
        }
        P.ga_src = reinterpret_cast<const double2 *>(X.mine + mbox_off_ga(X.nsend)) + (size_t)ga_par * X.nrecv + ga_off[p];
        P.ga_copy_dst = X.pair + recv_start[p];
        P.tot_dst = reinterpret_cast<double *>(pm + mbox_off_tot()) + ((size_t)tot_par * kMboxMaxRanks + X.rank) * 4;
        P.flag_dst = reinterpret_cast<unsigned long long *>(pm) + X.rank;
        P.flag_src = reinterpret_cast<const unsigned long long *>(X.mine) + p;
        s_peer[p] = P;
        if (p == 0) s_tot_in = reinterpret_cast<const double *>(X.mine + mbox_off_tot()) + (size_t)tot_par * kMboxMaxRanks * 4;
    }
    consumers_bar();
    XT_MARK(0);                                              // waited for the local grid
    unsigned long long sig = __ldcg(X.seq + 0);
    const int rank = X.rank;
    auto signal_wait = [&](unsigned long long seq) {
        // (every thread has fenced its own remote puts system-wide before the barrier)
        consumers_bar();
        if (ct < R && ct != rank) {
            st_release_sys(s_peer[ct].flag_dst, seq);
            const unsigned long long *f = s_peer[ct].flag_src;
            const unsigned long long t0 = global_ns();
            unsigned spins = 0;
            while (ld_acquire_sys(f) < seq) {
                if ((++spins & 1023) == 0 && global_ns() - t0 > 4000000000ull) die();     // 4 s: a peer died
            }
        }
        consumers_bar();
    };
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (o == 0) {
        if (R > 1) {
            // 1. halo partial sums -> the owners' scatter inboxes
            for (int p = 0; p < R; ++p) {
                const long long cnt = s_peer[p].cnt;
                if (cnt > 0) copy_batched(s_peer[p].sc_dst, s_peer[p].sc_src, cnt, ct);
            }
            __threadfence_system();
            XT_MARK(1);                                      // scatter puts + fence
            signal_wait(++sig);
            XT_MARK(2);                                      // signal round trip 1
        }
        // 2. boundary rows (add the peers' partial sums in rank order, Krylov row epilogue, fresh value -> gastage), shared
        //    with the helper CTAs; then the fresh values go to the peers' gather inboxes in whole slices
        const int K = loop_helpers(X, gsz);
        const unsigned long long nn = (unsigned long long)((ph + 2 - (L.first & 1)) >> 1);     // n-space exchanges of this launch so far
        if (K > 1 && ct == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(X.sbar), "l"(nn) : "memory");
        {
            const int chunk = (X.nbound + K - 1) / K;
            boundary_rows_slice(L, X, sC, 0, min(chunk, X.nbound), ct, acc);
        }
        if (R > 1) {
            consumers_bar();
            if (K > 1 && ct == 0) {
                const unsigned long long want = nn * (unsigned long long)(K - 1);
                const unsigned long long t0 = global_ns();
                unsigned spins = 0;
                while (ld_acquire_gpu(X.hbar) < want) {
                    __nanosleep(20);
                    if ((++spins & 4095) == 0 && global_ns() - t0 > 4000000000ull) die();
                }
            }
            if (K > 1) consumers_bar();
            for (int p = 0; p < R; ++p) {
                const long long cnt = s_peer[p].send_cnt;
                if (cnt > 0) copy_batched(s_peer[p].ga_put_dst, s_peer[p].ga_put_src, cnt, ct);
            }
        }
    }
    XT_MARK(3);                                              // boundary rows
    // 3. local sums: the CTAs' records in CTA order, then the boundary rows
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double x = warp_sum(acc[q]);
        if (lane == 0) s_red[q * 32 + cw] = x;
    }
    consumers_bar();
    if (cw == 0) {
        const double *base = L.parts + (size_t)(ph & 1) * gsz * 4;
        double tot[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = lane; i < gsz; i += 32) {
            const double2 a = __ldcg(reinterpret_cast<const double2 *>(base + (size_t)i * 4));
            const double2 b = __ldcg(reinterpret_cast<const double2 *>(base + (size_t)i * 4 + 2));
            tot[0] += a.x; tot[1] += a.y; tot[2] += b.x; tot[3] += b.y;
        }
        double bt[4] = {0.0, 0.0, 0.0, 0.0};                 // the helpers' slices of the boundary rows, in CTA order
        if (o == 0 && R > 1) {
            const int K = loop_helpers(X, gsz);
            for (int i = 1 + lane; i < K; i += 32) {
                const double2 a = __ldcg(reinterpret_cast<const double2 *>(X.bparts + (size_t)i * 4));
                const double2 b = __ldcg(reinterpret_cast<const double2 *>(X.bparts + (size_t)i * 4 + 2));
                bt[0] += a.x; bt[1] += a.y; bt[2] += b.x; bt[3] += b.y;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            tot[q] = warp_sum(tot[q]) + (warp_sum(lane < kGroups * kGroupWarps ? s_red[q * 32 + lane] : 0.0) + warp_sum(bt[q]));
        if (lane < 4) {
            const double v = lane == 0 ? tot[0] : (lane == 1 ? tot[1] : (lane == 2 ? tot[2] : tot[3]));
            s_x[lane] = v;
            for (int p = 0; p < R; ++p) {
                if (p == rank) continue;
                s_peer[p].tot_dst[lane] = v;
            }
        }
    }
    if (o == 0 || cw == 0) __threadfence_system();           // (after an m-space phase only warp 0 has put anything to the peers)
    XT_MARK(4);                                              // local sums + puts of the totals + fence
    if (R > 1) signal_wait(++sig); else consumers_bar();
    XT_MARK(5);                                              // signal round trip 2
    // 4. the peers' fresh values -> my halo slots ; the ranks' sums in rank order
    if (o == 0 && R > 1) {
        for (int p = 0; p < R; ++p) {
            const long long cnt = s_peer[p].cnt;
            if (cnt > 0) copy_batched(s_peer[p].ga_copy_dst, s_peer[p].ga_src, cnt, ct);
        }
    }
    if (ct < 4) {
        const double *in = s_tot_in;
        double v[kMboxMaxRanks];                             // all ranks' sums requested at once, added in rank order
#pragma unroll
        for (int r = 0; r < kMboxMaxRanks; ++r) v[r] = (r < R && r != rank) ? __ldcg(in + r * 4 + ct) : 0.0;
        double g = 0.0;
#pragma unroll
        for (int r = 0; r < kMboxMaxRanks; ++r) if (r < R) g += (r == rank) ? s_x[ct] : v[r];
        X.gtot[(size_t)(ph & 1) * 4 + ct] = g;
    }
    if (ct == 0) {
        X.seq[0] = sig;
        if (o == 0) { X.seq[1] += 1; X.seq[2] += 1; }
        X.seq[3] += 1;
    }
    fence_proxy_async();                                     // the halo / boundary entries of the pair vs the next phase's TMA reads
    __threadfence();
    consumers_bar();
    XT_MARK(6);                                              // gather copy-in, totals, fences
#ifdef FPSB_XCHG_TIMERS
    if (ct == 0) g_xchg_t[o][15] += 1;
#endif
    if (ct == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(X.xbar), "l"((unsigned long long)(ph + 1)) : "memory");
}

// second half of the boundary: slot r's recurrence on its two sums, the coefficients of phase ph, the stop decision
__device__ __forceinline__ void loop_boundary_finish(const LoopParams &L, int ph, int r, int lane, int cta, SlotState *sS, Coef *sC,
                                                     LoopCtl &ctl, double tot0, double tot1) {
    if (lane == 0) {
        const StepParams &Pp = L.op[(L.first + ph - 1) & 1];
        const int mode = Pp.io[r].mode;
        const double t0 = tot0, t1 = tot1;
        LT_STAMP(ph - 1, r ? 11 : 8);
        if (mode != MD_NONE && sS[r].active) finish_step(sS[r], mode, t0, t1);
        LT_STAMP(ph - 1, r ? 12 : 9);
        const StepParams &Pn = L.op[(L.first + ph) & 1];
        load_coef(sC[r], Pn.io[r], &sS[r], true);
        if (!(Pn.io[r].mode != MD_NONE && sS[r].active)) {
            sC[r].mode = MD_NONE; sC[r].rd0 = sC[r].rd1 = sC[r].wr0 = sC[r].wr1 = sC[r].rdself = 0;
        }
        LT_STAMP(ph - 1, r ? 13 : 10);
        if (r == 1) st_release_cta(&ctl.open1, ph + 1);
        else {
            loop_wait_thread(&ctl.open1, ph + 1, &ctl);
            if (!sS[0].active && !sS[1].active) ctl.stop_at = ph;
            st_release_cta(&ctl.open, ph + 1);
            LT_STAMP(ph - 1, 4);
        }
    }
    if (r == 1) loop_wait(&ctl.open, ph + 1, &ctl);          // warp 1 learns the stop decision from warp 0
    __syncwarp();
    if (r == 0 && cta == 0 && (ph >= L.nphase || ctl.stop_at <= ph)) {
        const double *src = reinterpret_cast<const double *>(sS);
        double *dst = reinterpret_cast<double *>(L.st);
        for (int i = lane; i < (int)(2 * sizeof(SlotState) / sizeof(double)); i += 32) dst[i] = src[i];
        if (lane == 0) {
            if (!sS[0].active && !sS[1].active) L.done_flag[0] = 1;
            L.done_flag[1] += ph;
        }
    }
}


// Boundary work of recurrence warp r (consumer warp 0 or 1) before phase `ph` (1 <= ph <= nphase): wait for
// every CTA's record of phase ph - 1, add them in a fixed order, run slot r's scalar recurrence, prepare the
// coefficients of phase ph.  Warp 0 also publishes pass / open / stop_at and, in CTA 0, the final states.
// (Inlined: a call inside the phase loop makes ptxas spill the tile loop. state around it.)
template <bool DIST>
__device__ __forceinline__ void loop_boundary(const LoopParams &L, int ph, int r, int lane, int cta, int gsz, SlotState *sS, Coef *sC,
                                              LoopCtl &ctl, double *rsum /* 64 * 4 */) {
    const double *base = L.parts + (size_t)((ph - 1) & 1) * gsz * 4;
    if (lane == 0) LT_STAMP(ph - 1, r ? 15 : 14);
    if (DIST) {
        // row-partitioned run: CTA 0 publishes the all-reduced sums after the exchange with the peers
        double tot[4] = {0.0, 0.0, 0.0, 0.0};
        if (r == 0 && lane == 0) {
            const unsigned long long t0 = global_ns();
            unsigned spins = 0;
            // After an m-space phase the next phase gathers the m-space pair, which has no halo: it is complete as soon as
            // this GPU's own CTAs arrived, so `pass` (early row sums, gather windows) does not wait for the peers; only the
            // coefficients do.  After an n-space phase the halo slots must have been refreshed first.
            const bool local_pass = ((L.first + ph - 1) & 1) == 1;
            if (local_pass) {
                const unsigned long long want = (unsigned long long)ph * (unsigned long long)gsz;
                while (ld_acquire_gpu(L.gbar) < want) {
                    __nanosleep(40);
                    if ((++spins & 4095) == 0 && global_ns() - t0 > 6000000000ull) {
                        ctl.abort = 1; atomicExch(L.done_flag + 2, 2); __threadfence_system(); __trap();
                    }
                }
                st_release_cta(&ctl.pass, ph);
            }
#ifdef FPSB_XCHG_TIMERS
            const unsigned long long xw0 = global_ns();
#endif
            while (ld_acquire_gpu(L.dx->xbar) < (unsigned long long)ph) {
                __nanosleep(40);
                if ((++spins & 4095) == 0 && global_ns() - t0 > 6000000000ull) {     // 6 s: the exchange died
                    ctl.abort = 1; atomicExch(L.done_flag + 2, 2); __threadfence_system(); __trap();
                }
            }
#ifdef FPSB_XCHG_TIMERS
            if (cta == 1) g_xchg_t[(L.first + ph - 1) & 1][12] += global_ns() - xw0;
#endif
            if (!local_pass) st_release_cta(&ctl.pass, ph);
            LT_STAMP(ph - 1, 3);
        }
        asm volatile("bar.sync %0, %1;" ::"n"(2 + kGroups), "n"(64) : "memory");      // the two recurrence warps
        if (lane == 0) {
            const double2 v = __ldcg(reinterpret_cast<const double2 *>(L.dx->gtot + (size_t)((ph - 1) & 1) * 4) + r);
            tot[0] = v.x; tot[1] = v.y;
        }
        loop_boundary_finish(L, ph, r, lane, cta, sS, sC, ctl, tot[0], tot[1]);
        return;
    }
#ifdef FPSB_LOOP_V1B
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    {
        if (lane == 0) {
            const unsigned long long want = (unsigned long long)ph * (unsigned long long)gsz;
            while (ld_acquire_gpu(L.gbar) < want) { }
        }
        __syncwarp();
        if (r == 0 && lane == 0) { st_release_cta(&ctl.pass, ph); LT_STAMP(ph - 1, 3); }
        for (int i = lane; i < gsz; i += 32) {
            const double2 a = __ldcg(reinterpret_cast<const double2 *>(base + (size_t)i * 4));
            const double2 b = __ldcg(reinterpret_cast<const double2 *>(base + (size_t)i * 4 + 2));
            tot[0] += a.x; tot[1] += a.y; tot[2] += b.x; tot[3] += b.y;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) tot[q] = warp_sum(tot[q]);
        if (r == 1) { tot[0] = tot[2]; tot[1] = tot[3]; }
    }
#else
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    {
        // ONE thread per CTA polls the arrival counter (148 x 64 lanes polling the records themselves took a
        // measurable share of the L2 request rate away from the CTAs that were still streaming)
        if (r == 0 && lane == 0) {
            const unsigned long long want = (unsigned long long)ph * (unsigned long long)gsz;
            const unsigned long long t0 = global_ns();
            unsigned spins = 0;
            while (ld_acquire_gpu(L.gbar) < want) {
                __nanosleep(40);
                if ((++spins & 4095) == 0 && global_ns() - t0 > 2000000000ull) {     // 2 s: a CTA died
                    ctl.abort = 1; atomicExch(L.done_flag + 2, 2); __threadfence_system(); __trap();
                }
            }
        }
        asm volatile("bar.sync %0, %1;" ::"n"(2 + kGroups), "n"(64) : "memory");      // the two recurrence warps
        // the 64 lanes of the two recurrence warps share the records: lane gl adds those of CTAs gl, gl + 64, gl + 128
        constexpr int kRecPerLane = kLoopMaxGrid / 64;
        const int gl = r * 32 + lane;
        double2 w[kRecPerLane][2];
#pragma unroll
        for (int j = 0; j < kRecPerLane; ++j) {
            const int i = gl + 64 * j < gsz ? gl + 64 * j : gl;          // (out of range: a valid record, not added)
            const double2 *rp = reinterpret_cast<const double2 *>(base + (size_t)i * 4);
            w[j][0] = __ldcg(rp); w[j][1] = __ldcg(rp + 1);
        }
#pragma unroll
        for (int j = 0; j < kRecPerLane; ++j) {
            if (gl + 64 * j < gsz) { tot[0] += w[j][0].x; tot[1] += w[j][0].y; tot[2] += w[j][1].x; tot[3] += w[j][1].y; }
        }
        // lane partials -> shared memory -> one lane per warp adds them in lane order.  (A shuffle tree is 20
        // DEPENDENT trips through the SM's memory-instruction queue, which the other consumers' row sums keep full
        // at this point: measured 5.6 us; this is two trips.)
        reinterpret_cast<double4 *>(rsum)[gl] = make_double4(tot[0], tot[1], tot[2], tot[3]);
        asm volatile("bar.sync %0, %1;" ::"n"(2 + kGroups), "n"(64) : "memory");
    }
    if (r == 0 && lane == 0) { st_release_cta(&ctl.pass, ph); LT_STAMP(ph - 1, 3); }
    if (lane == 0) {
        // this warp's slot needs two of the four sums
        const double2 *all = reinterpret_cast<const double2 *>(rsum) + r;
        tot[0] = tot[1] = 0.0;
#pragma unroll 8
        for (int l = 0; l < 64; ++l) { const double2 v = all[2 * l]; tot[0] += v.x; tot[1] += v.y; }
    }
#endif
    loop_boundary_finish(L, ph, r, lane, cta, sS, sC, ctl, tot[0], tot[1]);
}

template <bool DIST>
__global__ void __launch_bounds__(kStepThreads, 1) gk_loop_kernel(const __grid_constant__ LoopParams L) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ double2 s_sum[kGroups][2][kTileRows];        // [group][double buffer][row]
    __shared__ double s_red[4 * 32];
    __shared__ alignas(8) uint64_t full_bar[kMaxStages];
    __shared__ alignas(8) uint64_t empty_bar[kMaxStages];
    __shared__ int s_rel[kMaxStages];                       // rounds seen filled, per stage (see stage_turn_wait)
    __shared__ SlotState sS[2];
    __shared__ Coef sC[2];
    __shared__ LoopCtl ctl;
    __shared__ __align__(32) double s_rsum[64 * 4];        // lane partials of the two recurrence warps
#ifdef FPSB_LOOP_TIMERS
    __shared__ unsigned long long s_seg[kGroups][8];
    if (threadIdx.x < kGroups * 8) s_seg[threadIdx.x / 8][threadIdx.x % 8] = 0;
#endif

    const int tid = threadIdx.x;
    const int cta = blockIdx.x, gsz = (int)gridDim.x;
    const int nstage = L.op[0].nstage;                       // one ring geometry for both operators (host)
    const size_t stage_bytes = (size_t)L.op[0].stage_bytes;
    const size_t blk_cap = (size_t)L.op[0].blk_cap;

    // ---- prologue: slot states, ring barriers, coefficients of phase 0 ----
    {
        const double *g = reinterpret_cast<const double *>(L.st);
        double *d = reinterpret_cast<double *>(sS);
        for (int i = tid; i < (int)(2 * sizeof(SlotState) / sizeof(double)); i += kStepThreads) d[i] = __ldcg(g + i);
    }
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], kGroupWarps); s_rel[s] = 0; }
        mbar_fence_init();
        ctl.pass = 0; ctl.open1 = 1; ctl.open = 1; ctl.stop_at = 0x7fffffff; ctl.abort = 0;
    }
    __syncthreads();
    if (!sS[0].active && !sS[1].active) {                    // nothing left to do (chunk queued after convergence)
        if (cta == 0 && tid == 0) L.done_flag[0] = 1;
        return;
    }
    if (tid < 2) {
        const StepParams &P0 = L.op[L.first & 1];
        load_coef(sC[tid], P0.io[tid], &sS[tid], true);
        if (!(P0.io[tid].mode != MD_NONE && sS[tid].active)) {
            sC[tid].mode = MD_NONE; sC[tid].rd0 = sC[tid].rd1 = sC[tid].wr0 = sC[tid].wr1 = sC[tid].rdself = 0;
        }
    }
    __syncthreads();
    bool ok = true;

    if (tid < kProducerThreads) {
        reg_dec<kProducerRegs>();
        const int role = tid >> 5;
        const int par = role & 1;
        // multi-segment windows (stencil operators): lanes 0..3 of a window producer warp own one segment each and issue
        // their bulk copies side by side; everything else is lane 0's
        const int seg_lane = tid & 31;
        const bool segd = role >= 2 && (L.op[0].wsegs != nullptr || L.op[1].wsegs != nullptr);
        if (seg_lane != 0 && !(segd && seg_lane < 4)) return;
        const unsigned seg_mask = segd ? 0xfu : 0x1u;
        if (role < 2) {
            // ---- warps 0 / 1: tile blocks of the CTA's even / odd tiles.  The blocks do not depend on the previous
            //      phase: the first nspec tiles of a phase are requested before the phase is decided ----
            int kb = 0;                                      // tiles of this CTA in the phases before ph
            for (int ph = 0; ph < L.nphase && ok; ++ph) {
                const StepParams &P = L.op[(L.first + ph) & 1];
                const TileMeta *tiles = P.tiles;
                const int nt = P.ntiles;
                const int cnt = cta < nt ? (nt - cta + gsz - 1) / gsz : 0;
                bool decided = false, stop = false;
                // this warp's tiles are those with an even (odd) GLOBAL ring index k = kb + kl: with an even number of
                // stages every ring stage is then always filled by the same block / window producer (see the throttle)
#ifdef FPSB_LOOP_KLPAR
                const int kl0 = par;
#else
                const int kl0 = (par + kb) & 1;
#endif
                int tile = cta + kl0 * gsz;
                PTile T{}, T1{};
                if (kl0 < cnt) T = load_ptile(tiles, tile);
#ifdef FPSB_LOOP_TWOAHEAD
                PTile T2{};
                if (kl0 + 2 < cnt) T1 = load_ptile(tiles, tile + 2 * gsz);
#endif
                for (int kl = kl0; kl < cnt; kl += 2) {
                    if (!decided && kl >= L.nspec) {
                        loop_wait_thread(&ctl.open, ph + 1, &ctl);
                        decided = true;
                        if (ctl.stop_at <= ph) { stop = true; break; }
                    }
                    const int t1 = tile + 2 * gsz;
#ifdef FPSB_LOOP_TWOAHEAD
                    if (kl + 4 < cnt) T2 = load_ptile(tiles, t1 + 2 * gsz);
#else
                    if (kl + 2 < cnt) T1 = load_ptile(tiles, t1);
#endif
                    const int k = kb + kl, s = k % nstage;
                    if (k >= nstage) ok = mbar_wait(&empty_bar[s], (uint32_t)((k / nstage - 1) & 1)) && ok;
                    // bounded run-ahead (see gk_step_kernel): this warp's previous tile must have landed.  The wait only
                    // OBSERVES a full barrier, which is safe only while the observer cannot fall a whole barrier phase
                    // behind (the parity test would then wait for the wrong phase and never return): tile k - 2 sits in
                    // a stage that only this warp's own later tiles reuse (nstage is even), so it cannot.
                    if (k >= 2) ok = mbar_wait(&full_bar[(k - 2) % nstage], (uint32_t)(((k - 2) / nstage) & 1)) && ok;
                    const uint32_t bytes = tile_block_bytes(T.elems, T.ns, T.ccnt);
                    mbar_expect_tx(&full_bar[s], bytes);
                    if (bytes) tma_bulk_g2s(s_dyn + (size_t)s * stage_bytes, P.tbuf + T.boff, bytes, &full_bar[s]);
#ifdef FPSB_LOOP_TWOAHEAD
                    T = T1; T1 = T2; tile = t1;
#else
                    T = T1; tile = t1;
#endif
                }
                if (!ok || stop) break;
                if (!decided) {
                    loop_wait_thread(&ctl.open, ph + 1, &ctl);
                    if (ctl.stop_at <= ph) break;
                }
                kb += cnt;
            }
            if (!ok) { ctl.abort = 1; atomicExch(L.done_flag + 2, 1); __threadfence_system(); __trap(); }
            return;
        }
        // ---- warps 2 / 3: gather windows of the even / odd tiles ----
        int kb = 0;
        for (int ph = 0; ph < L.nphase && ok; ++ph) {
            const StepParams &P = L.op[(L.first + ph) & 1];
            const TileMeta *tiles = P.tiles;
            const int nt = P.ntiles;
            const int cnt = cta < nt ? (nt - cta + gsz - 1) / gsz : 0;
            bool opened = false;
            // the gather windows hold what the previous phase wrote: wait for the grid
            loop_wait_thread(L.early ? &ctl.pass : &ctl.open, L.early ? ph : ph + 1, &ctl);
            fence_proxy_async();                             // generic-proxy writes of the grid -> this thread's TMA reads
#ifdef FPSB_LOOP_KLPAR
            const int kl0 = par;
#else
            const int kl0 = (par + kb) & 1;
#endif
            int kl = kl0, tile = cta + kl0 * gsz;
            PTile T{}, T1{};
            const bool psegs = P.wsegs != nullptr;           // this phase's operator has multi-segment tiles
            int2 G = make_int2(0, 0), G1 = make_int2(0, 0);  // this lane's segment of tiles T / T1
            if (tile < nt) { T = load_ptile(tiles, tile); if (psegs) G = __ldg(P.wsegs + (size_t)tile * 4 + seg_lane); }
            bool stop = false;
            for (; kl < cnt; kl += 2) {
                if (!opened && kl >= L.nspec) {
                    loop_wait_thread(&ctl.open, ph + 1, &ctl);
                    opened = true;
                    if (ctl.stop_at <= ph) { stop = true; break; }
                }
                const int t1 = tile + 2 * gsz;
                if (t1 < nt) { T1 = load_ptile(tiles, t1); if (psegs) G1 = __ldg(P.wsegs + (size_t)t1 * 4 + seg_lane); }
                const int k = kb + kl, s = k % nstage;
                if (k >= nstage) ok = mbar_wait(&empty_bar[s], (uint32_t)((k / nstage - 1) & 1)) && ok;
                if (k >= 2) ok = mbar_wait(&full_bar[(k - 2) % nstage], (uint32_t)(((k - 2) / nstage) & 1)) && ok;
                const uint32_t wbytes = T.ccnt > 0 ? (uint32_t)T.ccnt * 16u : 0u;
                if (seg_lane == 0) mbar_expect_tx(&full_bar[s], wbytes);
                if (segd) __syncwarp(seg_mask);              // the expectation is posted before any lane's copy can complete
                if (psegs) {
                    const int glen = G.y >> 16, goff = G.y & 0xffff;
                    if (wbytes && glen > 0)
                        tma_bulk_g2s(s_dyn + (size_t)s * stage_bytes + blk_cap + (size_t)goff * 16, P.gin2 + G.x, (uint32_t)glen * 16u, &full_bar[s]);
                } else if (wbytes && seg_lane == 0)
                    tma_bulk_g2s(s_dyn + (size_t)s * stage_bytes + blk_cap, P.gin2 + T.cmin, wbytes, &full_bar[s]);
                T = T1; G = G1; tile = t1;
            }
            if (!ok || stop) break;
            if (!opened) {
                loop_wait_thread(&ctl.open, ph + 1, &ctl);
                if (ctl.stop_at <= ph) break;
            }
            kb += cnt;
        }
        if (!ok) { ctl.abort = 1; atomicExch(L.done_flag + 2, 1); __threadfence_system(); __trap(); }
        return;
    }

    // ------------------------------- consumers -------------------------------
    reg_inc<kConsumerRegs>();
    const int ct = tid - kProducerThreads;
    const int g = ct / kGroupThreads, t = ct % kGroupThreads;
    const int lane = t & 31, wid = t >> 5;
    const int cw = ct >> 5;                                  // consumer warp 0..11
    int kb = 0;
    for (int ph = 0; ph < L.nphase; ++ph) {
        const StepParams &P = L.op[(L.first + ph) & 1];
        const TileMeta *tiles = P.tiles;
        const int nt = P.ntiles;
        const int cnt = cta < nt ? (nt - cta + gsz - 1) / gsz : 0;
        bool have_coef = false;
        CoefR C0{}, C1{};
        bool act0 = false, act1 = false;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        // global gathers read what the previous phase wrote
        loop_wait(L.early ? &ctl.pass : &ctl.open, L.early ? ph : ph + 1, &ctl);
        if (!L.early) {
            if (ctl.stop_at <= ph) {
                // drain the tiles the block producers requested before the phase was decided
                for (int kl = g; kl < cnt && kl < L.nspec; kl += kGroups) {
                    const int k = kb + kl, s = k % nstage;
                    stage_turn_wait(&s_rel[s], k / nstage);
                    ok = mbar_wait(&full_bar[s], (uint32_t)((k / nstage) & 1)) && ok;
                    if (t == 0) stage_seen(&s_rel[s], k / nstage);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                }
                return;
            }
            C0 = to_regs(sC[0]); C1 = to_regs(sC[1]);
            act0 = C0.mode != MD_NONE; act1 = C1.mode != MD_NONE;
            have_coef = true;
        }
        if (ct == 0) LT_STAMP(ph, 7);
        int kl = g, tile = cta + kl * gsz;
        CTile T{}, Tn{}, Tnn{};
        if (tile < nt) T = load_ctile(tiles, tile);
        if (tile + kGroups * gsz < nt) Tn = load_ctile(tiles, tile + kGroups * gsz);
        LS_DECL;
        for (; kl < cnt; kl += kGroups) {
            LS_RESET; LS_COUNT;
            const int k = kb + kl, s = k % nstage;
            const int ntile = tile + kGroups * gsz, nntile = tile + 2 * kGroups * gsz;
            if (nntile < nt) Tnn = load_ctile(tiles, nntile);
            if (!have_coef && kl >= L.nspec) {
                // this tile is only requested once the phase is decided
                loop_wait(&ctl.open, ph + 1, &ctl);
                if (ctl.stop_at <= ph) return;
                C0 = to_regs(sC[0]); C1 = to_regs(sC[1]);
                act0 = C0.mode != MD_NONE; act1 = C1.mode != MD_NONE;
                have_coef = true;
            }
            const int rowA = T.row0 + t, rowB = rowA + kGroupThreads;
            const bool inA = t < T.nrows(), inB = t + kGroupThreads < T.nrows();
            const int flagA = (P.rowflag != nullptr && inA) ? (int)P.rowflag[rowA] : 0;
            const int flagB = (P.rowflag != nullptr && inB) ? (int)P.rowflag[rowB] : 0;
            RowOps RA{};
            bool loaded = false;
            if (have_coef) {
                if (inA) load_row_ops<true>(P, C0, C1, rowA, RA);
                loaded = true;
                if (ntile < nt) {
                    if (t < Tn.nrows()) prefetch_row_ops<true>(P, C0, C1, Tn.row0 + t, t);
                    if (t + kGroupThreads < Tn.nrows()) prefetch_row_ops<true>(P, C0, C1, Tn.row0 + kGroupThreads + t, t);
                }
            }
            const unsigned char *st = s_dyn + (size_t)s * stage_bytes;
            LS_MARK(0);
            stage_turn_wait(&s_rel[s], k / nstage);
            ok = mbar_wait(&full_bar[s], (uint32_t)((k / nstage) & 1)) && ok;
            if (t == 0) stage_seen(&s_rel[s], k / nstage);
            LS_MARK(1);
            if (ct == 0 && kl == 0) LT_STAMP(ph, 0);
            // ---------------- phase 1: row sums ----------------
            double2 *sum = s_sum[g][(kl / kGroups) & 1];
            tile_row_sums<true, true>(P, T, st, st + blk_cap, true, true, wid, lane, sum);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            LS_MARK(2);
            group_bar(g);
            LS_MARK(3);
            if (!have_coef) {
                // first tile of the phase: its row sums overlapped the recurrences
                loop_wait(&ctl.open, ph + 1, &ctl);
                if (ctl.stop_at <= ph) return;
                C0 = to_regs(sC[0]); C1 = to_regs(sC[1]);
                act0 = C0.mode != MD_NONE; act1 = C1.mode != MD_NONE;
                have_coef = true;
            }
            if (!loaded && inA) load_row_ops<true>(P, C0, C1, rowA, RA);
            LS_MARK(4);
            // ---------------- phase 2: row epilogue in natural row order ----------------
            RowOps RB{};
            if (inB) load_row_ops<true>(P, C0, C1, rowB, RB);
            if (inA) finish_row<true>(P, C0, C1, act0, act1, rowA, flagA, sum[t], RA, acc);
            if (inB) finish_row<true>(P, C0, C1, act0, act1, rowB, flagB, sum[t + kGroupThreads], RB, acc);
            T = Tn; Tn = Tnn;
            tile = ntile;
            LS_MARK(5);
        }
        if (!have_coef) {
            // no tile of this group in the phase: still learn whether the phase runs
            loop_wait(&ctl.open, ph + 1, &ctl);
            if (ctl.stop_at <= ph) return;
        }
        kb += cnt;
        if (t == 0) LT_STAMP(ph, g == 0 ? 1 : (g == 1 ? 5 : 6));
        LS_FLUSH;

        // ---- end of the phase: the CTA's record (arrival at the grid barrier + its four norm partials) ----
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double x = warp_sum(acc[q]);
            if (lane == 0) s_red[q * 32 + cw] = x;
        }
        fence_proxy_async();                                  // this phase's generic writes vs the next phase's TMA reads
        consumers_bar();
        if (cw == 0) {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = warp_sum(lane < kGroups * kGroupWarps ? s_red[q * 32 + lane] : 0.0);
            if (lane == 0) {
                LT_STAMP(ph, 2);
                if (!ok) atomicExch(L.done_flag + 2, 1);
                double *pp = L.parts + ((size_t)(ph & 1) * gsz + cta) * 4;
                pp[0] = v[0]; pp[1] = v[1]; pp[2] = v[2]; pp[3] = v[3];
                red_release_gpu(L.gbar, 1ull);               // release: everything this CTA wrote in the phase + its record
            }
        }
        // consumer warps 0 / 1: the boundary before phase ph + 1 (the other warps go on and wait for `pass` / `open`)
        if (DIST && cta == 0) loop_exchange(L, *L.dx, ph, ct, cw, lane, gsz, sC, s_red, s_rsum, ctl);
        if (DIST && cta != 0) loop_help(L, *L.dx, ph, ct, cw, lane, cta, gsz, sC, s_red, ctl);
        if (cw < 2) loop_boundary<DIST>(L, ph + 1, cw, lane, cta, gsz, sS, sC, ctl, s_rsum);
    }
}
