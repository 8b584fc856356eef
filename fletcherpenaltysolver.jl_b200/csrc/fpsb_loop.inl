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
struct LoopParams {
    StepParams op[2];            // [0] n-space step (rows of A'), [1] m-space step (rows of A); io modes fixed for the loop
    int first;                   // operator of phase 0
    int nphase;                  // phases this launch may run at most
    int nspec;                   // tiles of a phase the block producers may request before the phase opens
    int early;                   // 1: row sums of the first tiles overlap the scalar recurrences
    SlotState *st;               // slot states in / out
    double *parts;               // [2][grid][4] per-CTA norm partials, ping-pong by phase parity
    unsigned long long *gbar;    // arrival counter of the grid barrier (zero at launch)
    int *done_flag;              // [0] done, [1] += phases run, [2] error
};

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
__device__ __forceinline__ void consumers_bar() {
    asm volatile("bar.sync %0, %1;" ::"n"(1 + kGroups), "n"(kGroups * kGroupThreads) : "memory");
}
// spin until *p >= want (shared-memory counter published with st_release_cta); false: the CTA aborted
__device__ __forceinline__ bool loop_wait(const int *p, int want, const LoopCtl *ctl) {
    for (unsigned it = 0;; ++it) {
        if (ld_acquire_cta(p) >= want) return true;
        if ((it & 15) == 15) {
            if (*reinterpret_cast<const volatile int *>(&ctl->abort)) __trap();    // error path: end the launch, never hang
            __nanosleep(32);
        }
    }
}

__global__ void __launch_bounds__(kStepThreads, 1) gk_loop_kernel(const __grid_constant__ LoopParams L) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ double2 s_sum[kGroups][2][kTileRows];
    __shared__ double s_red[4 * 32];
    __shared__ alignas(8) uint64_t full_bar[kMaxStages];
    __shared__ alignas(8) uint64_t empty_bar[kMaxStages];
    __shared__ SlotState sS[2];
    __shared__ Coef sC[2];
    __shared__ LoopCtl ctl;

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
        for (int s = 0; s < nstage; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], kGroupWarps); }
        mbar_fence_init();
        ctl.pass = 0; ctl.open1 = 1; ctl.open = 1; ctl.stop_at = 0x7fffffff; ctl.abort = 0;
    }
    __syncthreads();
    if (!sS[0].active && !sS[1].active) {                    // nothing left to do (chunk queued after convergence)
        if (cta == 0 && tid == 0) L.done_flag[0] = 1;
        return;
    }
    if (tid < 2) {
        const StepParams *P0 = &L.op[L.first & 1];
        load_coef(sC[tid], P0->io[tid], &sS[tid], true);
        if (!(P0->io[tid].mode != MD_NONE && sS[tid].active)) {
            sC[tid].mode = MD_NONE; sC[tid].rd0 = sC[tid].rd1 = sC[tid].wr0 = sC[tid].wr1 = sC[tid].rdself = 0;
        }
    }
    __syncthreads();
    bool ok = true;

    if (tid < kProducerThreads) {
        if ((tid & 31) != 0) return;
        const int role = tid >> 5;
        const bool blocks = role < 2;
        const int par = role & 1;
        int kb = 0;                                          // tiles of this CTA in the phases before ph
        for (int ph = 0; ph < L.nphase && ok; ++ph) {
            const StepParams *P = &L.op[(L.first + ph) & 1];
            const TileMeta *tiles = P->tiles;
            const int nt = P->ntiles;
            const int cnt = cta < nt ? (nt - cta + gsz - 1) / gsz : 0;
            bool opened = false;
            if (!blocks) {
                // the gather windows hold what the previous phase wrote: wait for the grid
                if (!loop_wait(L.early ? &ctl.pass : &ctl.open, L.early ? ph : ph + 1, &ctl)) { ok = false; break; }
                fence_proxy_async();                         // generic-proxy writes of the grid -> this thread's TMA reads
            }
            int kl = par, tile = cta + par * gsz, t1 = tile + 2 * gsz;
            TileMeta T{}, T1{};
            if (tile < nt) T = load_tile(tiles, tile);
            if (t1 < nt) T1 = load_tile(tiles, t1);
            bool stop = false;
            for (; kl < cnt; kl += 2) {
                if (!opened && kl >= L.nspec) {
                    if (!loop_wait(&ctl.open, ph + 1, &ctl)) { ok = false; break; }
                    opened = true;
                    if (ctl.stop_at <= ph) { stop = true; break; }
                }
                const int t2 = t1 + 2 * gsz;
                TileMeta T2{};
                if (t2 < nt) T2 = load_tile(tiles, t2);
                const int k = kb + kl, s = k % nstage;
                if (k >= nstage) ok = mbar_wait(&empty_bar[s], (uint32_t)((k / nstage - 1) & 1)) && ok;
                if (k >= P->inflight) {
                    const int kq = k - P->inflight;
                    ok = mbar_wait(&full_bar[kq % nstage], (uint32_t)((kq / nstage) & 1)) && ok;
                }
                unsigned char *st = s_dyn + (size_t)s * stage_bytes;
                if (blocks) {
                    const uint32_t bytes = tile_block_bytes(T.elems, T.ns, T.ccnt);
                    mbar_expect_tx(&full_bar[s], bytes);
                    if (bytes) tma_bulk_g2s(st, P->tbuf + T.boff, bytes, &full_bar[s]);
                } else {
                    const uint32_t wbytes = T.ccnt > 0 ? (uint32_t)T.ccnt * 16u : 0u;
                    mbar_expect_tx(&full_bar[s], wbytes);
                    if (wbytes) tma_bulk_g2s(st + blk_cap, P->gin2 + T.cmin, wbytes, &full_bar[s]);
                }
                T = T1; T1 = T2; tile = t1; t1 = t2;
            }
            if (!ok || stop) break;
            if (!opened) {
                if (!loop_wait(&ctl.open, ph + 1, &ctl)) { ok = false; break; }
                if (ctl.stop_at <= ph) break;
            }
            kb += cnt;
        }
        if (!ok) { ctl.abort = 1; atomicExch(L.done_flag + 2, 1); __threadfence_system(); __trap(); }
        return;
    }

    // ------------------------------- consumers -------------------------------
    const int ct = tid - kProducerThreads;
    const int g = ct / kGroupThreads, t = ct % kGroupThreads;
    const int lane = t & 31, wid = t >> 5;
    const int cw = ct >> 5;                                  // consumer warp 0..11 ; warps 0 and 1 run the recurrences
    int kb = 0;
    for (int ph = 0; ph < L.nphase; ++ph) {
        const StepParams *P = &L.op[(L.first + ph) & 1];
        const TileMeta *tiles = P->tiles;
        const int nt = P->ntiles;
        const int cnt = cta < nt ? (nt - cta + gsz - 1) / gsz : 0;
        bool have_coef = false;
        CoefR C0{}, C1{};
        bool act0 = false, act1 = false;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        // global gathers / the epilogue read what the previous phase wrote
        if (!loop_wait(L.early ? &ctl.pass : &ctl.open, L.early ? ph : ph + 1, &ctl)) return;
        if (!L.early) {
            if (ctl.stop_at <= ph) {
                // drain the tiles the block producers requested before the phase was decided
                for (int kl = g; kl < cnt && kl < L.nspec; kl += kGroups) {
                    const int k = kb + kl, s = k % nstage;
                    ok = mbar_wait(&full_bar[s], (uint32_t)((k / nstage) & 1)) && ok;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                }
                return;
            }
            C0 = to_regs(sC[0]); C1 = to_regs(sC[1]);
            act0 = C0.mode != MD_NONE; act1 = C1.mode != MD_NONE;
            have_coef = true;
        }
        int kl = g, tile = cta + kl * gsz;
        CTile T{}, Tn{}, Tnn{};
        if (tile < nt) T = load_ctile(tiles, tile);
        if (tile + kGroups * gsz < nt) Tn = load_ctile(tiles, tile + kGroups * gsz);
        for (; kl < cnt; kl += kGroups) {
            const int k = kb + kl, s = k % nstage;
            const int ntile = tile + kGroups * gsz, nntile = tile + 2 * kGroups * gsz;
            if (nntile < nt) Tnn = load_ctile(tiles, nntile);
            if (!have_coef && kl >= L.nspec) {
                // this tile is only requested once the phase is decided
                if (!loop_wait(&ctl.open, ph + 1, &ctl)) return;
                if (ctl.stop_at <= ph) return;
                C0 = to_regs(sC[0]); C1 = to_regs(sC[1]);
                act0 = C0.mode != MD_NONE; act1 = C1.mode != MD_NONE;
                have_coef = true;
            }
            const int row = T.row0 + t;
            const bool in_row = t < T.nrows();
            int rflag = (P->rowflag != nullptr && in_row) ? (int)P->rowflag[row] : 0;
            double2 old2 = make_double2(0.0, 0.0);
            double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
            bool loaded = false;
            if (have_coef && in_row) {
                old2 = P->self2[row];
                if (C0.rd0()) a00 = __ldcs(P->io[0].a0 + row);
                if (C0.rd1()) a01 = __ldcs(P->io[0].a1 + row);
                if (C1.rd0()) a10 = __ldcs(P->io[1].a0 + row);
                if (C1.rd1()) a11 = __ldcs(P->io[1].a1 + row);
                loaded = true;
            }
            if (have_coef && ntile < nt && t < Tn.nrows()) {
                const int r2 = Tn.row0 + t;
                if ((t & 7) == 0) prefetch_l2(P->self2 + r2);
                if ((t & 15) == 0) {
                    if (C0.rd0()) prefetch_l2(P->io[0].a0 + r2);
                    if (C0.rd1()) prefetch_l2(P->io[0].a1 + r2);
                    if (C1.rd0()) prefetch_l2(P->io[1].a0 + r2);
                    if (C1.rd1()) prefetch_l2(P->io[1].a1 + r2);
                }
            }
            const unsigned char *st = s_dyn + (size_t)s * stage_bytes;
            const double *s_val = reinterpret_cast<const double *>(st);
            const bool windowed = T.ccnt > 0;
            const int *s_col = reinterpret_cast<const int *>(st + (size_t)T.elems * 8);
            const unsigned short *s_c16 = reinterpret_cast<const unsigned short *>(st + (size_t)T.elems * 8);
            const unsigned char *s_map = st + (size_t)T.elems * (windowed ? 10 : 12);
            const double2 *win2 = reinterpret_cast<const double2 *>(st + blk_cap);

            ok = mbar_wait(&full_bar[s], (uint32_t)((k / nstage) & 1)) && ok;
            // ---------------- phase 1: row sums, a warp per slice ----------------
            double2 *sum = s_sum[g][(kl / kGroups) & 1];
            {
                int off = 0, width = T.width(0);
#pragma unroll
                for (int i = 1; i < kTileSlices; ++i) if (i <= wid) { off += T.width(i - 1) * 32; width = T.width(i); }
                const int npair = width >> 1;
                double s0 = 0.0, s1 = 0.0, u0 = 0.0, u1 = 0.0;
                int lrow = -1;
                if (wid < T.ns()) {
                    lrow = s_map[wid * 32 + lane];
                    if (lrow == 255) lrow = -1;
                    const double2 *sv = reinterpret_cast<const double2 *>(s_val + off) + lane;
                    const bool tail = (width & 1) != 0;
                    if (windowed) {
                        const ushort2 *sc16 = reinterpret_cast<const ushort2 *>(s_c16 + off) + lane;
#pragma unroll 5
                        for (int p = 0; p < npair; ++p) {
                            const double2 v = sv[p * 32];
                            const ushort2 c = sc16[p * 32];
                            const double2 x0 = win2[c.x], x1 = win2[c.y];
                            s0 = fma(v.x, x0.x, s0); s1 = fma(v.x, x0.y, s1);
                            u0 = fma(v.y, x1.x, u0); u1 = fma(v.y, x1.y, u1);
                        }
                        if (tail) {
                            const double v = s_val[off + npair * 64 + lane];
                            const double2 x = win2[s_c16[off + npair * 64 + lane]];
                            s0 = fma(v, x.x, s0); s1 = fma(v, x.y, s1);
                        }
                    } else {
                        // (L2 loads: the gathered pair changes from phase to phase inside this kernel)
                        const double2 *gin2 = P->gin2;
                        const int2 *sc = reinterpret_cast<const int2 *>(s_col + off) + lane;
#pragma unroll 5
                        for (int p = 0; p < npair; ++p) {
                            const double2 v = sv[p * 32];
                            const int2 c = sc[p * 32];
                            const double2 x0 = __ldcg(gin2 + c.x), x1 = __ldcg(gin2 + c.y);
                            s0 = fma(v.x, x0.x, s0); s1 = fma(v.x, x0.y, s1);
                            u0 = fma(v.y, x1.x, u0); u1 = fma(v.y, x1.y, u1);
                        }
                        if (tail) {
                            const double v = s_val[off + npair * 64 + lane];
                            const double2 x = __ldcg(gin2 + s_col[off + npair * 64 + lane]);
                            s0 = fma(v, x.x, s0); s1 = fma(v, x.y, s1);
                        }
                    }
                }
                if (lrow >= 0) sum[lrow] = make_double2(s0 + u0, s1 + u1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            group_bar(g);
            if (!have_coef) {
                // first tile of the phase: its row sums overlapped the recurrences
                if (!loop_wait(&ctl.open, ph + 1, &ctl)) return;
                if (ctl.stop_at <= ph) return;
                C0 = to_regs(sC[0]); C1 = to_regs(sC[1]);
                act0 = C0.mode != MD_NONE; act1 = C1.mode != MD_NONE;
                have_coef = true;
            }
            if (!loaded && in_row) {
                old2 = P->self2[row];
                if (C0.rd0()) a00 = __ldcs(P->io[0].a0 + row);
                if (C0.rd1()) a01 = __ldcs(P->io[0].a1 + row);
                if (C1.rd0()) a10 = __ldcs(P->io[1].a0 + row);
                if (C1.rd1()) a11 = __ldcs(P->io[1].a1 + row);
            }
            // ---------------- phase 2: row epilogue in natural row order, a thread per row ----------------
            if (rflag == 2 && P->raw_out == nullptr) rflag = 0;
            if (rflag == 2) P->raw_out[row] = sum[t];
            if (in_row && rflag == 0) {
                const double2 sm = sum[t];
                double n0 = old2.x, n1 = old2.y;
                if (act0) n0 = row_epilogue(C0, sm.x, old2.x, a00, a01, acc[0], acc[1]);
                if (act1) n1 = row_epilogue(C1, sm.y, old2.y, a10, a11, acc[2], acc[3]);
                P->self2[row] = make_double2(n0, n1);
                if (C0.wr0()) __stcs(P->io[0].a0 + row, a00);
                if (C0.wr1()) __stcs(P->io[0].a1 + row, a01);
                if (C1.wr0()) __stcs(P->io[1].a0 + row, a10);
                if (C1.wr1()) __stcs(P->io[1].a1 + row, a11);
            }
            T = Tn; Tn = Tnn;
            tile = ntile;
        }
        if (!have_coef) {
            // no tile of this group in the phase: still learn whether the phase runs
            if (!loop_wait(&ctl.open, ph + 1, &ctl)) return;
            if (ctl.stop_at <= ph) return;
        }
        kb += cnt;

        // ---- end of the phase: CTA partial -> grid barrier -> totals -> recurrences ----
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double x = warp_sum(acc[q]);
            if (lane == 0) s_red[q * 32 + cw] = x;
        }
        fence_proxy_async();                                  // this phase's generic writes vs the next phase's TMA reads
        consumers_bar();
        if (cw < 2) {
            if (cw == 0) {
                double v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = warp_sum(lane < kGroups * kGroupWarps ? s_red[q * 32 + lane] : 0.0);
                if (lane == 0) {
                    double *pp = L.parts + ((size_t)(ph & 1) * gsz + cta) * 4;
                    pp[0] = v[0]; pp[1] = v[1]; pp[2] = v[2]; pp[3] = v[3];
                    if (!ok) atomicExch(L.done_flag + 2, 1);
                    red_release_gpu(L.gbar, 1ull);
                }
            }
            bool fine = true;
            if (lane == 0) {
                const unsigned long long want = (unsigned long long)(ph + 1) * (unsigned long long)gsz;
                const unsigned long long t0 = global_ns();
                unsigned it = 0;
                while (ld_acquire_gpu(L.gbar) < want) {
                    if ((++it & 255) == 0 && global_ns() - t0 > 2000000000ull) { fine = false; break; }   // 2 s: a CTA died
                }
            }
            fine = __shfl_sync(0xffffffffu, fine ? 1 : 0, 0) != 0;
            if (!fine) {
                if (lane == 0) { ctl.abort = 1; atomicExch(L.done_flag + 2, 2); __threadfence_system(); }
                __trap();
            }
            if (cw == 0 && lane == 0) st_release_cta(&ctl.pass, ph + 1);
            // totals: every CTA adds the per-CTA partials in the same order
            double tot[4] = {0.0, 0.0, 0.0, 0.0};
            {
                const double *base = L.parts + (size_t)(ph & 1) * gsz * 4;
                for (int i = lane; i < gsz; i += 32) {
                    const double2 a = __ldcg(reinterpret_cast<const double2 *>(base + (size_t)i * 4));
                    const double2 b = __ldcg(reinterpret_cast<const double2 *>(base + (size_t)i * 4 + 2));
                    tot[0] += a.x; tot[1] += a.y; tot[2] += b.x; tot[3] += b.y;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) tot[q] = warp_sum(tot[q]);
            }
            const bool last = ph + 1 >= L.nphase;
            if (lane == 0) {
                const int mode = P->io[cw].mode;
                const double t0 = cw ? tot[2] : tot[0], t1 = cw ? tot[3] : tot[1];
                if (mode != MD_NONE && sS[cw].active) finish_step(sS[cw], mode, t0, t1);
                const StepParams *Pn = &L.op[(L.first + ph + 1) & 1];
                load_coef(sC[cw], Pn->io[cw], &sS[cw], true);
                if (!(Pn->io[cw].mode != MD_NONE && sS[cw].active)) {
                    sC[cw].mode = MD_NONE; sC[cw].rd0 = sC[cw].rd1 = sC[cw].wr0 = sC[cw].wr1 = sC[cw].rdself = 0;
                }
                if (cw == 1) st_release_cta(&ctl.open1, ph + 2);
                else {
                    loop_wait(&ctl.open1, ph + 2, &ctl);
                    if (!sS[0].active && !sS[1].active) ctl.stop_at = ph + 1;
                    st_release_cta(&ctl.open, ph + 2);
                }
            }
            if (cw == 0 && cta == 0) {
                __syncwarp();
                const bool fin = last || *reinterpret_cast<volatile int *>(&ctl.stop_at) <= ph + 1;
                if (fin) {
                    const double *src = reinterpret_cast<const double *>(sS);
                    double *dst = reinterpret_cast<double *>(L.st);
                    for (int i = lane; i < (int)(2 * sizeof(SlotState) / sizeof(double)); i += 32) dst[i] = src[i];
                    if (lane == 0) {
                        if (!sS[0].active && !sS[1].active) L.done_flag[0] = 1;
                        L.done_flag[1] += ph + 1;
                    }
                }
            }
        }
    }
}
