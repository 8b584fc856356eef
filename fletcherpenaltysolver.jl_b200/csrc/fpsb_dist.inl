// fpsb_dist.inl — row-partitioned Krylov path over several GPUs (BASELINE config C3).
// Included at the end of fpsb_krylov.cu (same translation unit: it drives the same kernels).
//
// One process per GPU.  Rank r owns a block of constraint rows of A (m-space) and a contiguous block
// of variables (n-space).  Its local operator is A_loc (m_loc x n_ext) where the n_ext "extended"
// columns are the owned block plus the halo columns its rows touch, kept in global column order so
// the tile windows stay narrow:   [ halo of lower ranks | owned | halo of higher ranks ].
//
//   jprod  (m-space step)  needs the halo entries of the gathered n-space pair:  owners pack the
//                          requested entries, ncclSend/ncclRecv lands them in the halo slots, then
//                          the ordinary fused step kernel runs on A_loc.
//   jtprod (n-space step)  A_loc' u produces partial sums for every extended row: the halo partials
//                          travel to their owners (ncclSend/ncclRecv), are added there in rank
//                          order, and the Krylov row epilogue runs on the owned rows only.
//   norms                  the kernels leave their local sums in a 4-double buffer (tot_out);
//                          ncclAllReduce(sum) + finish_kernel replace the fused last-CTA recurrences.
//
// Every rank runs the same scalar recurrences on the same all-reduced sums, so the slot states stay
// bitwise identical across ranks and all ranks take the same number of iterations.
//
// Two transports for the three exchanges:
//   * peer memory (default once fpsb_dist_peer_attach has run): every rank owns a "mailbox" in HBM that
//     its peers map through CUDA IPC and write over NVLink.  ONE small kernel per half iteration
//     (xchg_kernel) does, in order: put the halo partial sums into the owners' mailboxes -> signal ->
//     wait -> add them in rank order -> epilogue of the boundary rows -> put the fresh halo values of
//     the gathered pair and the 4 local norm sums into the peers' mailboxes -> signal -> wait -> sum the
//     norm partials in rank order -> scalar recurrences.  No NCCL call, no host involvement; an
//     iteration is 4 launches (step, xchg, step, xchg) instead of 8 kernels + 4 NCCL operations.
//   * NCCL (FPSB_DIST_NCCL=1, or when peer attach was not called): ncclSend/ncclRecv + ncclAllReduce.
// NCCL is bound at run time (dlopen) so that libfpsb200.so loads on machines without it.
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // an NCCL already in the process (e.g. PyTorch's) wins; FPSB_NCCL_LIB overrides the search
        const char *names[] = {getenv("FPSB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            api.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (api.lib) break;
        }
        if (api.lib) {
#define FPSB_SYM(field, name) *(void **)(&api.field) = dlsym(api.lib, name)
            FPSB_SYM(GetUniqueId, "ncclGetUniqueId");
            FPSB_SYM(CommInitRank, "ncclCommInitRank");
            FPSB_SYM(CommDestroy, "ncclCommDestroy");
            FPSB_SYM(AllReduce, "ncclAllReduce");
            FPSB_SYM(Send, "ncclSend");
            FPSB_SYM(Recv, "ncclRecv");
            FPSB_SYM(GroupStart, "ncclGroupStart");
            FPSB_SYM(GroupEnd, "ncclGroupEnd");
            FPSB_SYM(GetErrorString, "ncclGetErrorString");
#undef FPSB_SYM
            if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd)
                api.lib = nullptr;
        }
    }
    return api.lib ? &api : nullptr;
}

#define FPSB_NCCL(call)                                                                   \
    do {                                                                                  \
        ncclResult_t r__ = (call);                                                        \
        if (r__ != ncclSuccess) {                                                         \
            NcclApi *a__ = nccl_api();                                                    \
            set_error("%s:%d: %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, (int)r__, \
                      (a__ && a__->GetErrorString) ? a__->GetErrorString(r__) : "?");     \
            throw CudaFail{FPSB_ECUDA};                                                   \
        }                                                                                 \
    } while (0)

struct DistCtx {
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    int64_t own_off = 0, n_own = 0;
    std::vector<int64_t> recv_start, recv_cnt;      // per peer: my halo slots [start, start + cnt) of the extended space
    std::vector<int64_t> send_ptr;                  // per peer: range of send_idx
    DevBuf<int> send_idx;                           // owned extended indices other ranks keep as halo
    DevBuf<double2> sendbuf, recvbuf;               // packed entries (sum of send counts)
    DevBuf<double2> S;                              // raw A_loc' partial sums of both columns (n_ext)
    DevBuf<double> tot;                             // 4 doubles: local sums -> all-reduced sums
    DevBuf<int> bidx;                               // owned rows that receive halo contributions (sorted, unique)
    int64_t nbound = 0;
    bool fused_n = false;                           // n-space step fused for the interior rows
    int64_t nsend = 0;
    // ---- peer-memory transport ----
    static constexpr int kMaxRanks = 8;
    bool peer = false;                              // mailboxes mapped, xchg_kernel replaces NCCL
    unsigned char *mbox = nullptr;                  // my mailbox (cudaMalloc, exported through CUDA IPC)
    size_t mbox_bytes = 0;
    int64_t nrecv = 0;                              // sum of recv_cnt
    std::vector<int64_t> ga_off;                    // per source peer: start (in double2) of its slice of my gather inbox
    unsigned char *peer_mbox[kMaxRanks] = {};       // peers' mailboxes mapped into this process (self = mbox)
    int64_t sc_at_peer[kMaxRanks] = {}, ga_at_peer[kMaxRanks] = {};   // where my slices start in the peers' inboxes
    int64_t peer_nsend[kMaxRanks] = {}, peer_nrecv[kMaxRanks] = {};   // the peers' inbox sizes (mailbox layout)
    DevBuf<int64_t> d_meta;                         // device copy of recv_start | recv_cnt | send_ptr | ga_off
    DevBuf<int> d_err;
    DevBuf<unsigned long long> d_bar;               // grid-barrier arrivals of xchg_kernel (monotonic)
    unsigned long long bar_base = 0;
    static constexpr int kXchgMaxGrid = 64;
    DevBuf<double> xparts;                          // per-CTA norm partials of the boundary rows
    DevBuf<int> brow;                               // per boundary row {row, first entry, end entry, inbox position of the first entry}: one 16-byte load
    DevBuf<int> bptr, bsrc, bpeer;                  // per boundary row: its entries of the send list (= scatter-inbox positions) and their peers
    DevBuf<PeerTail> d_ptail;                       // for the m-space step kernel's own exchange (see gk_step_kernel)
    uint64_t sig = 0, n_sc = 0, n_ga = 0, n_tot = 0;   // signals / exchanges issued so far (same on every rank)
    bool halo_fresh = false;                        // the halo slots of Gn hold the current values
    bool tail = false;                              // m-space steps all-reduce inside the step kernel's last CTA (experiment)
    // ---- persistent loop with the exchange inside the kernel (gk_loop_kernel<true>) ----
    DevBuf<DistLoop> d_dloop;
    DevBuf<unsigned long long> d_seq;               // device copy of sig | n_sc | n_ga | n_tot while the loop kernel runs
    DevBuf<double> gtot, bparts;
    DevBuf<double2> gastage;                        // fresh boundary values in send-list order (see DistLoop)
    bool loop_ok = false;
};

// mailbox layout: mbox_off_tot / mbox_off_sc / mbox_off_ga (fpsb_loop.inl)
static inline size_t mbox_size(int64_t nsend, int64_t nrecv) { return mbox_off_ga(nsend) + 2 * (size_t)nrecv * sizeof(double2) + 64; }

// what a rank publishes so that its peers can map and address its mailbox
struct PeerBlob {
    cudaIpcMemHandle_t handle;
    int64_t nsend, nrecv;
    int64_t sc_off[DistCtx::kMaxRanks];             // per source peer p: where p's partial sums land (double2 units) = send_ptr[p]
    int64_t ga_off[DistCtx::kMaxRanks];             // per source peer p: where p's halo values land
};

void dist_free(Handle *h) {
    if (!h->dist) return;
    NcclApi *api = nccl_api();
    if (h->dist->comm && api && api->CommDestroy) api->CommDestroy(h->dist->comm);
    for (int p = 0; p < h->dist->nranks && p < DistCtx::kMaxRanks; ++p)
        if (p != h->dist->rank && h->dist->peer_mbox[p]) cudaIpcCloseMemHandle(h->dist->peer_mbox[p]);
    if (h->dist->mbox) cudaFree(h->dist->mbox);
    delete h->dist;
    h->dist = nullptr;
}

int64_t dist_n_own(Handle *h) { return h->dist ? h->dist->n_own : 0; }

void dist_unique_id(void *out128) {
    NcclApi *api = nccl_api();
    if (!api) { set_error("NCCL (libnccl.so.2) is not available on this machine"); throw CudaFail{FPSB_ECUDA}; }
    ncclUniqueId id;
    FPSB_NCCL(api->GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
}

void dist_attach(Handle *h, int nranks, int rank, const void *id128, int64_t own_off, int64_t n_own,
                 const int64_t *recv_start, const int64_t *recv_cnt, const int64_t *send_ptr, const int64_t *send_idx) {
    NcclApi *api = nccl_api();
    if (!api) { set_error("NCCL (libnccl.so.2) is not available on this machine"); throw CudaFail{FPSB_ECUDA}; }
    dist_free(h);
    iter_setup(h);
    DistCtx *D = new DistCtx();
    h->dist = D;
    D->nranks = nranks; D->rank = rank; D->own_off = own_off; D->n_own = n_own;
    D->recv_start.assign(recv_start, recv_start + nranks);
    D->recv_cnt.assign(recv_cnt, recv_cnt + nranks);
    D->send_ptr.assign(send_ptr, send_ptr + nranks + 1);
    D->nsend = send_ptr[nranks];
    std::vector<int> idx((size_t)D->nsend);
    for (int64_t i = 0; i < D->nsend; ++i) idx[(size_t)i] = (int)send_idx[i];
    D->send_idx.from(idx, h->stream);
    D->sendbuf.alloc((size_t)D->nsend + 8);
    D->recvbuf.alloc((size_t)D->nsend + 8);
    D->S.alloc((size_t)h->nvar + 8);
    D->tot.alloc(8);
    D->tot.zero(h->stream);
    // Fused n-space step: interior owned rows get the ordinary fused epilogue; halo rows and the owned
    // rows other ranks contribute to ("boundary") only leave their raw sums (row flag 2) and are finished
    // by boundary_epilogue_kernel after the exchange.  Needs the lane-per-row tiles (no long rows in A').
    {
        std::vector<unsigned char> flag((size_t)h->nvar + 8, 0);
        for (int64_t i = 0; i < h->nvar; ++i) if (i < own_off || i >= own_off + n_own) flag[(size_t)i] = 2;
        std::vector<int> b(idx.begin(), idx.end());
        std::sort(b.begin(), b.end());
        b.erase(std::unique(b.begin(), b.end()), b.end());
        for (int v : b) flag[(size_t)v] = 2;
        D->nbound = (int64_t)b.size();
        D->fused_n = h->At.nlong == 0 && D->nbound <= 65536;
        {
            // per boundary row the positions of its contributions in the scatter inbox, peers in rank order
            std::vector<int> bp(b.size() + 1, 0), bs(idx.size());
            for (size_t i = 0; i < idx.size(); ++i) bp[(size_t)(std::lower_bound(b.begin(), b.end(), idx[i]) - b.begin()) + 1]++;
            for (size_t r = 0; r < b.size(); ++r) bp[r + 1] += bp[r];
            std::vector<int> fill(bp.begin(), bp.end() - 1);
            for (size_t i = 0; i < idx.size(); ++i) bs[(size_t)fill[(size_t)(std::lower_bound(b.begin(), b.end(), idx[i]) - b.begin())]++] = (int)i;
            std::vector<int> bpr(bs.size());
            for (size_t k = 0; k < bs.size(); ++k) {
                int pp = 0;
                while (pp + 1 < nranks && (int64_t)bs[k] >= send_ptr[pp + 1]) ++pp;
                bpr[k] = pp;
            }
            std::vector<int> br(4 * b.size() + 8, 0);
            for (size_t r = 0; r < b.size(); ++r) {
                br[4 * r] = b[r]; br[4 * r + 1] = bp[r]; br[4 * r + 2] = bp[r + 1];
                br[4 * r + 3] = bp[r] < bp[r + 1] ? bs[(size_t)bp[r]] : -1;
            }
            D->brow.from(br, h->stream);
            D->bptr.from(bp, h->stream);
            D->bsrc.from(bs, h->stream);
            D->bpeer.from(bpr, h->stream);
            D->bidx.from(b, h->stream);
        }
        if (D->fused_n) {
            FPSB_CUDA(cudaMemcpyAsync(h->At.rowflag.p, flag.data(), (size_t)h->nvar, cudaMemcpyHostToDevice, h->stream));
            h->At.has_raw_rows = true;
        }
    }
    FPSB_CUDA(cudaStreamSynchronize(h->stream));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    FPSB_NCCL(api->CommInitRank(&D->comm, nranks, id, rank));
}

// ---- peer-memory transport: mailbox export / mapping ---------------------------------------------
int64_t dist_peer_blob_bytes() { return (int64_t)sizeof(PeerBlob); }

// allocates this rank's mailbox and fills the blob its peers need (the host program all-gathers the blobs)
void dist_peer_export(Handle *h, void *blob_out) {
    DistCtx *D = h->dist;
    if (D->nranks > DistCtx::kMaxRanks) { set_error("peer-memory transport supports up to %d ranks", DistCtx::kMaxRanks); throw CudaFail{FPSB_EINVAL}; }
    if (!D->mbox) {
        D->nrecv = 0;
        D->ga_off.assign((size_t)D->nranks, 0);
        for (int p = 0; p < D->nranks; ++p) { D->ga_off[(size_t)p] = D->nrecv; D->nrecv += D->recv_cnt[(size_t)p]; }
        D->mbox_bytes = mbox_size(D->nsend, D->nrecv);
        FPSB_CUDA(cudaMalloc((void **)&D->mbox, D->mbox_bytes));
        FPSB_CUDA(cudaMemset(D->mbox, 0, D->mbox_bytes));
        std::vector<int64_t> meta;
        meta.insert(meta.end(), D->recv_start.begin(), D->recv_start.end());
        meta.insert(meta.end(), D->recv_cnt.begin(), D->recv_cnt.end());
        meta.insert(meta.end(), D->send_ptr.begin(), D->send_ptr.end());
        meta.insert(meta.end(), D->ga_off.begin(), D->ga_off.end());
        D->d_meta.from(meta, h->stream);
        D->d_err.alloc(8);
        D->d_err.zero(h->stream);
        D->d_bar.alloc(8);
        D->d_bar.zero(h->stream);
        D->xparts.alloc((size_t)DistCtx::kXchgMaxGrid * 4 + 8);
        D->xparts.zero(h->stream);
        FPSB_CUDA(cudaStreamSynchronize(h->stream));
    }
    PeerBlob B;
    memset(&B, 0, sizeof(B));
    FPSB_CUDA(cudaIpcGetMemHandle(&B.handle, D->mbox));
    B.nsend = D->nsend; B.nrecv = D->nrecv;
    for (int p = 0; p < D->nranks; ++p) { B.sc_off[p] = D->send_ptr[(size_t)p]; B.ga_off[p] = D->ga_off[(size_t)p]; }
    memcpy(blob_out, &B, sizeof(B));
}

// blobs: nranks PeerBlobs in rank order (every rank passes the same array)
void dist_peer_attach(Handle *h, const void *blobs) {
    DistCtx *D = h->dist;
    if (!D->mbox) { set_error("fpsb_dist_peer_attach: call fpsb_dist_peer_export first"); throw CudaFail{FPSB_ESTATE}; }
    const PeerBlob *B = reinterpret_cast<const PeerBlob *>(blobs);
    for (int p = 0; p < D->nranks; ++p) {
        D->peer_nsend[p] = B[p].nsend; D->peer_nrecv[p] = B[p].nrecv;
        D->sc_at_peer[p] = B[p].sc_off[D->rank];
        D->ga_at_peer[p] = B[p].ga_off[D->rank];
        if (p == D->rank) { D->peer_mbox[p] = D->mbox; continue; }
        // what I send to p in the scatter is what p counts as "sent to me" in the gather, and vice versa
        if (B[p].sc_off[D->rank] + D->recv_cnt[(size_t)p] > B[p].nsend || B[p].ga_off[D->rank] + (D->send_ptr[(size_t)p + 1] - D->send_ptr[(size_t)p]) > B[p].nrecv) {
            set_error("fpsb_dist_peer_attach: the halo lists of ranks %d and %d do not match", D->rank, p);
            throw CudaFail{FPSB_EINVAL};
        }
        void *ptr = nullptr;
        FPSB_CUDA(cudaIpcOpenMemHandle(&ptr, B[p].handle, cudaIpcMemLazyEnablePeerAccess));
        D->peer_mbox[p] = reinterpret_cast<unsigned char *>(ptr);
    }
    {
        PeerTail T{};
        T.nranks = D->nranks; T.rank = D->rank; T.mine = D->mbox; T.err = D->d_err.p;
        for (int p = 0; p < D->nranks; ++p) T.peer[p] = D->peer_mbox[p];
        std::vector<PeerTail> v(1, T);
        D->d_ptail.from(v, h->stream);
        FPSB_CUDA(cudaStreamSynchronize(h->stream));
    }
    const char *force = getenv("FPSB_DIST_NCCL");
    D->peer = !(force && *force && *force != '0');
    {
        // the description the persistent loop kernel's CTA 0 works from (see DistLoop)
        D->d_seq.alloc(4); D->d_seq.zero(h->stream);
        D->gtot.alloc(8); D->gtot.zero(h->stream);
        D->bparts.alloc((size_t)kLoopMaxGrid * 4 + 8); D->bparts.zero(h->stream);
        D->gastage.alloc((size_t)D->nsend + 8); D->gastage.zero(h->stream);
        DistLoop X{};
        X.nranks = D->nranks; X.rank = D->rank;
        X.nbound = (int)D->nbound; X.brow = reinterpret_cast<const int4 *>(D->brow.p); X.bidx = D->bidx.p; X.bptr = D->bptr.p; X.bsrc = D->bsrc.p; X.bpeer = D->bpeer.p;
        X.meta = reinterpret_cast<const long long *>(D->d_meta.p);
        X.S = D->S.p; X.pair = h->iter->Gn.p;
        X.nsend = D->nsend; X.nrecv = D->nrecv;
        X.mine = D->mbox;
        for (int p = 0; p < D->nranks; ++p) {
            X.peer[p] = D->peer_mbox[p];
            X.sc_at_peer[p] = D->sc_at_peer[p]; X.ga_at_peer[p] = D->ga_at_peer[p];
            X.peer_nsend[p] = D->peer_nsend[p]; X.peer_nrecv[p] = D->peer_nrecv[p];
        }
        X.seq = D->d_seq.p; X.gtot = D->gtot.p; X.xbar = h->iter->gbar.p + 1; X.err = D->d_err.p;
        // helper CTAs for the boundary rows (FPSB_DIST_HELPERS=1: CTA 0 alone)
        { const char *e = getenv("FPSB_DIST_HELPERS"); X.nhelp = std::max(1, std::min(kLoopMaxGrid, (e && *e) ? atoi(e) : 16)); }
        X.sbar = h->iter->gbar.p + 2; X.hbar = h->iter->gbar.p + 3; X.bparts = D->bparts.p; X.gastage = D->gastage.p;
        std::vector<DistLoop> v(1, X);
        D->d_dloop.from(v, h->stream);
        FPSB_CUDA(cudaStreamSynchronize(h->stream));
        const char *lp = getenv("FPSB_DIST_LOOP");
        D->loop_ok = D->peer && D->fused_n && !(lp && *lp == '0');
    }
    // experiment (off by default: measured 130.6 vs 128.5 us per iteration on 2 GPUs, no gain over the
    // separate single-CTA exchange launch): FPSB_DIST_TAIL=1 lets the m-space step kernel's last CTA all-reduce
    const char *tl = getenv("FPSB_DIST_TAIL");
    D->tail = (tl && *tl && *tl != '0');
}
bool dist_peer_active(Handle *h) { return h->dist && h->dist->peer; }

// ---- kernels of the exchange ------------------------------------------------------------------------
__global__ void pack_pairs_kernel(int n, const int *idx, const double2 *src, double2 *dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void unpack_cols_kernel(int n, const double2 *src, double *dst0, double *dst1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { if (dst0) dst0[i] = src[i].x; if (dst1) dst1[i] = src[i].y; }
}
// owner side of the scatter-add: one launch per peer, peers in rank order (deterministic sums)
__global__ void add_pairs_kernel(int n, const int *idx, const double2 *src, double2 *dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double2 a = src[i]; double2 d = dst[idx[i]]; d.x += a.x; d.y += a.y; dst[idx[i]] = d; }
}

// Krylov row epilogue over the OWNED rows of the n-space after the scatter-add (the step kernel's
// phase 2, un-fused): S holds the complete row sums, self2 the interleaved pair of this row space
struct DistEpiParams {
    int n_own, own_off;
    const double2 *S;
    double2 *self2;
    SlotIO io[2];
    SlotState *st;
    double *partials;
    unsigned *counter;
    double *tot_out;
};
__global__ void __launch_bounds__(kBlock) dist_epilogue_kernel(DistEpiParams P) {
    __shared__ double s_red[4 * 32];
    __shared__ Coef sC[2];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const bool act0 = P.io[0].mode != MD_NONE && P.st[0].active;
    const bool act1 = P.io[1].mode != MD_NONE && P.st[1].active;
    if (!act0 && !act1) return;
    if (tid == 0) {
        load_coef(sC[0], P.io[0], &P.st[0], true);
        load_coef(sC[1], P.io[1], &P.st[1], true);
        if (!act0) { sC[0].mode = MD_NONE; sC[0].rd0 = sC[0].rd1 = sC[0].wr0 = sC[0].wr1 = sC[0].rdself = 0; }
        if (!act1) { sC[1].mode = MD_NONE; sC[1].rd0 = sC[1].rd1 = sC[1].wr0 = sC[1].wr1 = sC[1].rdself = 0; }
    }
    __syncthreads();
    const CoefR C0 = to_regs(sC[0]), C1 = to_regs(sC[1]);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + tid; i < P.n_own; i += gridDim.x * blockDim.x) {
        const int row = P.own_off + i;
        const double2 sm = P.S[row];
        const double2 old2 = P.self2[row];
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
        if (C0.rd0()) a00 = P.io[0].a0[row];
        if (C0.rd1()) a01 = P.io[0].a1[row];
        if (C1.rd0()) a10 = P.io[1].a0[row];
        if (C1.rd1()) a11 = P.io[1].a1[row];
        double n0 = old2.x, n1 = old2.y;
        if (act0) n0 = row_epilogue(C0, sm.x, old2.x, a00, a01, acc[0], acc[1]);
        if (act1) n1 = row_epilogue(C1, sm.y, old2.y, a10, a11, acc[2], acc[3]);
        P.self2[row] = make_double2(n0, n1);
        if (C0.wr0()) P.io[0].a0[row] = a00;
        if (C0.wr1()) P.io[0].a1[row] = a01;
        if (C1.wr0()) P.io[1].a0[row] = a10;
        if (C1.wr1()) P.io[1].a1[row] = a11;
    }
    block_sum<4>(acc, s_red);
    if (tid == 0) {
        double *pp = P.partials + (size_t)blockIdx.x * 4;
        pp[0] = acc[0]; pp[1] = acc[1]; pp[2] = acc[2]; pp[3] = acc[3];
        __threadfence();
        const unsigned t = atomicAdd(P.counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < (int)gridDim.x; i += kBlock) {
        const double *pp = P.partials + (size_t)i * 4;
        tot[0] += __ldcg(pp + 0); tot[1] += __ldcg(pp + 1); tot[2] += __ldcg(pp + 2); tot[3] += __ldcg(pp + 3);
    }
    block_sum<4>(tot, s_red);
    if (tid == 0) {
        P.tot_out[0] = tot[0]; P.tot_out[1] = tot[1]; P.tot_out[2] = tot[2]; P.tot_out[3] = tot[3];
        *P.counter = 0;
    }
}

// epilogue of the boundary rows after the exchange (one CTA: the list is short), its norm sums are
// added to the ones the fused step kernel left in tot
__device__ __forceinline__ void boundary_epilogue(int nb, const int *bidx, const DistEpiParams &P, double *s_red /* 4*32 */, Coef *sC /* 2 */) {
    const int tid = threadIdx.x;
    const bool act0 = P.io[0].mode != MD_NONE && P.st[0].active;
    const bool act1 = P.io[1].mode != MD_NONE && P.st[1].active;
    if (!act0 && !act1) return;                              // uniform over the CTA
    if (tid == 0) {
        load_coef(sC[0], P.io[0], &P.st[0], true);
        load_coef(sC[1], P.io[1], &P.st[1], true);
        if (!act0) { sC[0].mode = MD_NONE; sC[0].rd0 = sC[0].rd1 = sC[0].wr0 = sC[0].wr1 = sC[0].rdself = 0; }
        if (!act1) { sC[1].mode = MD_NONE; sC[1].rd0 = sC[1].rd1 = sC[1].wr0 = sC[1].wr1 = sC[1].rdself = 0; }
    }
    __syncthreads();
    const CoefR C0 = to_regs(sC[0]), C1 = to_regs(sC[1]);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < nb; i += (int)blockDim.x) {
        const int row = bidx[i];
        const double2 sm = P.S[row];
        const double2 old2 = P.self2[row];
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
        if (C0.rd0()) a00 = P.io[0].a0[row];
        if (C0.rd1()) a01 = P.io[0].a1[row];
        if (C1.rd0()) a10 = P.io[1].a0[row];
        if (C1.rd1()) a11 = P.io[1].a1[row];
        double n0 = old2.x, n1 = old2.y;
        if (act0) n0 = row_epilogue(C0, sm.x, old2.x, a00, a01, acc[0], acc[1]);
        if (act1) n1 = row_epilogue(C1, sm.y, old2.y, a10, a11, acc[2], acc[3]);
        P.self2[row] = make_double2(n0, n1);
        if (C0.wr0()) P.io[0].a0[row] = a00;
        if (C0.wr1()) P.io[0].a1[row] = a01;
        if (C1.wr0()) P.io[1].a0[row] = a10;
        if (C1.wr1()) P.io[1].a1[row] = a11;
    }
    block_sum<4>(acc, s_red);
    if (tid == 0) { P.tot_out[0] += acc[0]; P.tot_out[1] += acc[1]; P.tot_out[2] += acc[2]; P.tot_out[3] += acc[3]; }
}
__global__ void __launch_bounds__(kBlock) boundary_epilogue_kernel(int nb, const int *bidx, DistEpiParams P) {
    __shared__ double s_red[4 * 32];
    __shared__ Coef sC[2];
    boundary_epilogue(nb, bidx, P, s_red, sC);
}

// the scalar recurrences on the all-reduced sums (what the last CTA does on a single GPU)
__global__ void finish_kernel(SlotState *st, int kind, int m0, int m1, const double *tot, int *done_flag) {
    __shared__ SlotState sS[2];
    const int tid = threadIdx.x;
    for (int i = tid; i < (int)(2 * sizeof(SlotState) / sizeof(double)); i += blockDim.x)
        reinterpret_cast<double *>(sS)[i] = reinterpret_cast<const double *>(st)[i];
    __syncthreads();
    if (tid == 0) {
        if (kind == 0) {                      // step: m0 / m1 are the modes of the two slots
            if (m0 != MD_NONE && sS[0].active) finish_step(sS[0], m0, tot[0], tot[1]);
            if (m1 != MD_NONE && sS[1].active) finish_step(sS[1], m1, tot[2], tot[3]);
        } else {                              // element-wise: m0 = op, m1 = slot
            const bool is_init = (m0 == EW_INIT_LSQR || m0 == EW_INIT_CRAIG || m0 == EW_MINRES_INIT || m0 == EW_CGLS_INIT);
            if (is_init || sS[m1].active) finish_ew(sS[m1], m0, tot[0]);
        }
        if (!sS[0].active && !sS[1].active) *done_flag = 1;
    }
    __syncthreads();
    for (int i = tid; i < (int)(2 * sizeof(SlotState) / sizeof(double)); i += blockDim.x)
        reinterpret_cast<double *>(st)[i] = reinterpret_cast<const double *>(sS)[i];
}

// ---- peer-memory exchange -------------------------------------------------------------------------
struct XchgParams {
    int nranks, rank;
    int do_scatter, do_epi, do_gather, kind, m0, m1;   // kind: -1 no recurrence, 0 step (m0/m1 = modes), 1 element-wise (m0 = op, m1 = slot)
    unsigned long long sig;                            // signals issued before this exchange
    unsigned long long bar_base;                       // arrivals at the grid barrier before this launch
    int sc_par, ga_par, tot_par;                       // inbox halves this exchange uses
    long long nsend, nrecv;                            // my inbox sizes (entries)
    const long long *meta;                             // recv_start[R] | recv_cnt[R] | send_ptr[R+1] | ga_off[R]
    const int *send_idx;
    double2 *S;                                        // raw / partial row sums of the extended n-space
    double2 *pair;                                     // gathered pair of the extended n-space
    int nbound;
    const int *bidx, *bptr, *bsrc, *bpeer;             // boundary rows; per row its send-list entries (rank order) and their peers
    DistEpiParams Q;
    double *tot;
    double *xparts;                                    // per-CTA norm partials of the boundary rows
    SlotState *st;
    int *done_flag;
    int *err;
    unsigned long long *bar;
    unsigned char *mine;
    unsigned char *peer[DistCtx::kMaxRanks];
    long long sc_at_peer[DistCtx::kMaxRanks], ga_at_peer[DistCtx::kMaxRanks];
    long long peer_nsend[DistCtx::kMaxRanks], peer_nrecv[DistCtx::kMaxRanks];
};

// "Last CTA" ticket on a monotonic counter: every CTA arrives after its own writes (remote puts included)
// are visible system-wide; the one that completes the round learns that everybody's are.
__device__ __forceinline__ bool xchg_ticket(const XchgParams &P, unsigned long long last_value, int *s_flag) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long old = atomicAdd(P.bar, 1ull);
        *s_flag = (old == last_value);
        __threadfence();
    }
    __syncthreads();
    return *s_flag != 0;
}
__device__ __forceinline__ void xchg_signal(const XchgParams &P, unsigned long long seq) {
    const int t = threadIdx.x;
    if (t < P.nranks && t != P.rank) st_release_sys(reinterpret_cast<unsigned long long *>(P.peer[t]) + P.rank, seq);
}
__device__ __forceinline__ void xchg_wait(const XchgParams &P, unsigned long long seq) {
    const int t = threadIdx.x;
    if (t < P.nranks && t != P.rank) {
        const unsigned long long *f = reinterpret_cast<const unsigned long long *>(P.mine) + t;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(f) < seq) {
            if (global_ns() - t0 > 4000000000ull) { atomicExch(P.err, 1); break; }     // 4 s: a peer died
        }
    }
    __syncthreads();
}

// <= 64 CTAs, all resident (the kernel is stream-ordered behind a step kernel).  Sequence:
//   1. halo partial sums -> owners' scatter inboxes ; ticket ; the last CTA signals ; everybody waits for the peers
//   2. a thread per boundary row: add the peers' partials (rank order), Krylov row epilogue, put the row's
//      fresh pair value into the gather inboxes of the peers that keep it as halo ; per-CTA norm partials
//   3. ticket ; the last CTA sums the partials, puts the four local sums to every peer and signals ;
//      everybody waits for the peers, unpacks its share of the gather inbox ; the last CTA sums the ranks'
//      norm sums in rank order and runs the scalar recurrences
__global__ void __launch_bounds__(kBlock) xchg_kernel(XchgParams P) {
    __shared__ double s_red[4 * 32];
    __shared__ Coef sC[2];
    __shared__ SlotState sS[2];
    __shared__ int s_flag;
    const int tid = threadIdx.x, nt = (int)blockDim.x, R = P.nranks;
    const long long gtid = (long long)blockIdx.x * nt + tid, gsz = (long long)gridDim.x * nt;
    const long long *recv_start = P.meta, *recv_cnt = P.meta + R, *send_ptr = P.meta + 2 * R, *ga_off = P.meta + 3 * R + 1;
    // static index data of this thread's first boundary row: requested before anything else
    int row0 = -1, k0 = 0, k1 = 0;
    if (gtid < P.nbound) { row0 = P.bidx[gtid]; k0 = P.bptr[gtid]; k1 = P.bptr[gtid + 1]; }
    if (*reinterpret_cast<volatile int *>(P.err) != 0) return;
    unsigned long long seq = P.sig, bar = P.bar_base;
    const unsigned long long G = gridDim.x;

    // ---- 1
    if (P.do_scatter) {
        for (int p = 0; p < R; ++p) {
            if (p == P.rank) continue;
            const long long cnt = recv_cnt[p];
            if (cnt == 0) continue;
            double2 *dst = reinterpret_cast<double2 *>(P.peer[p] + mbox_off_sc()) + (size_t)P.sc_par * P.peer_nsend[p] + P.sc_at_peer[p];
            const double2 *src = P.S + recv_start[p];
            for (long long i = gtid; i < cnt; i += gsz) dst[i] = src[i];
        }
        bar += G;
        if (xchg_ticket(P, bar - 1, &s_flag)) xchg_signal(P, seq + 1);
        xchg_wait(P, ++seq);
    }

    // ---- 2
    const bool act0 = P.do_epi && P.Q.io[0].mode != MD_NONE && P.Q.st[0].active;
    const bool act1 = P.do_epi && P.Q.io[1].mode != MD_NONE && P.Q.st[1].active;
    const bool epi = act0 || act1;
    if (epi) {
        if (tid == 0) {
            load_coef(sC[0], P.Q.io[0], &P.Q.st[0], true);
            load_coef(sC[1], P.Q.io[1], &P.Q.st[1], true);
            if (!act0) { sC[0].mode = MD_NONE; sC[0].rd0 = sC[0].rd1 = sC[0].wr0 = sC[0].wr1 = sC[0].rdself = 0; }
            if (!act1) { sC[1].mode = MD_NONE; sC[1].rd0 = sC[1].rd1 = sC[1].wr0 = sC[1].wr1 = sC[1].rdself = 0; }
        }
        __syncthreads();
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (P.do_scatter || epi || P.do_gather) {
        CoefR C0{}, C1{};
        if (epi) { C0 = to_regs(sC[0]); C1 = to_regs(sC[1]); }
        const double2 *inbox = reinterpret_cast<const double2 *>(P.mine + mbox_off_sc()) + (size_t)P.sc_par * P.nsend;
        for (long long b = gtid; b < P.nbound; b += gsz) {
            int row = row0, kb = k0, ke = k1;
            if (b != gtid) { row = P.bidx[b]; kb = P.bptr[b]; ke = P.bptr[b + 1]; }
            double2 val = P.pair[row];                      // the row's entry of the gathered pair (== Q.self2 when epi)
            if (P.do_scatter || epi) {
                double2 sm = P.S[row];
                if (P.do_scatter) {
                    for (int k = kb; k < ke; ++k) {
                        const double2 a = __ldcg(inbox + P.bsrc[k]);
                        sm.x += a.x; sm.y += a.y;
                    }
                    P.S[row] = sm;
                }
                if (epi) {
                    const double2 old2 = val;
                    double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
                    if (C0.rd0()) a00 = P.Q.io[0].a0[row];
                    if (C0.rd1()) a01 = P.Q.io[0].a1[row];
                    if (C1.rd0()) a10 = P.Q.io[1].a0[row];
                    if (C1.rd1()) a11 = P.Q.io[1].a1[row];
                    double n0 = old2.x, n1 = old2.y;
                    if (act0) n0 = row_epilogue(C0, sm.x, old2.x, a00, a01, acc[0], acc[1]);
                    if (act1) n1 = row_epilogue(C1, sm.y, old2.y, a10, a11, acc[2], acc[3]);
                    val = make_double2(n0, n1);
                    P.pair[row] = val;
                    if (C0.wr0()) P.Q.io[0].a0[row] = a00;
                    if (C0.wr1()) P.Q.io[0].a1[row] = a01;
                    if (C1.wr0()) P.Q.io[1].a0[row] = a10;
                    if (C1.wr1()) P.Q.io[1].a1[row] = a11;
                }
            }
            if (P.do_gather) {
                // the peers that keep this row as halo: entry i of my send list belongs to peer bpeer[k]
                for (int k = kb; k < ke; ++k) {
                    const int i = P.bsrc[k], p = P.bpeer[k];
                    double2 *dst = reinterpret_cast<double2 *>(P.peer[p] + mbox_off_ga(P.peer_nsend[p])) + (size_t)P.ga_par * P.peer_nrecv[p] + P.ga_at_peer[p];
                    dst[i - send_ptr[p]] = val;
                }
            }
        }
    }
    if (epi) {
        block_sum<4>(acc, s_red);
        if (tid == 0) { double *pp = P.xparts + (size_t)blockIdx.x * 4; pp[0] = acc[0]; pp[1] = acc[1]; pp[2] = acc[2]; pp[3] = acc[3]; }
    }

    // ---- 3
    bar += G;
    const bool last = xchg_ticket(P, bar - 1, &s_flag);
    const bool talk = P.do_gather || P.kind >= 0;
    if (last) {
        if (epi && tid < 4) {
            double v = P.tot[tid];
            for (unsigned c = 0; c < gridDim.x; ++c) v += __ldcg(P.xparts + (size_t)c * 4 + tid);
            P.tot[tid] = v;
        }
        __syncthreads();
        if (P.kind >= 0 && tid < 4 && R > 1) {
            const double v = P.tot[tid];
            for (int p = 0; p < R; ++p) {
                if (p == P.rank) continue;
                reinterpret_cast<double *>(P.peer[p] + mbox_off_tot())[((size_t)P.tot_par * DistCtx::kMaxRanks + P.rank) * 4 + tid] = v;
            }
            __threadfence_system();
        }
        __syncthreads();
        if (talk && R > 1) xchg_signal(P, seq + 1);
    }
    if (talk && R > 1) xchg_wait(P, ++seq);
    if (P.do_gather) {
        const double2 *inbox = reinterpret_cast<const double2 *>(P.mine + mbox_off_ga(P.nsend)) + (size_t)P.ga_par * P.nrecv;
        for (int p = 0; p < R; ++p) {
            if (p == P.rank) continue;
            const long long cnt = recv_cnt[p];
            double2 *dst = P.pair + recv_start[p];
            const double2 *src = inbox + ga_off[p];
            for (long long i = gtid; i < cnt; i += gsz) dst[i] = __ldcg(src + i);
        }
    }
    if (P.kind >= 0 && last) {
        // norm sums of all ranks in rank order (identical on every rank), then the scalar recurrences
        for (int i = tid; i < (int)(2 * sizeof(SlotState) / sizeof(double)); i += nt)
            reinterpret_cast<double *>(sS)[i] = reinterpret_cast<const double *>(P.st)[i];
        __syncthreads();
        if (tid == 0) {
            double tot[4] = {0.0, 0.0, 0.0, 0.0};
            const double *in = reinterpret_cast<const double *>(P.mine + mbox_off_tot()) + (size_t)P.tot_par * DistCtx::kMaxRanks * 4;
            for (int r = 0; r < R; ++r)
                for (int k = 0; k < 4; ++k) tot[k] += (r == P.rank) ? P.tot[k] : __ldcg(in + r * 4 + k);
            for (int k = 0; k < 4; ++k) P.tot[k] = tot[k];
            if (P.kind == 0) {
                if (P.m0 != MD_NONE && sS[0].active) finish_step(sS[0], P.m0, tot[0], tot[1]);
                if (P.m1 != MD_NONE && sS[1].active) finish_step(sS[1], P.m1, tot[2], tot[3]);
            } else {
                const bool is_init = (P.m0 == EW_INIT_LSQR || P.m0 == EW_INIT_CRAIG || P.m0 == EW_MINRES_INIT || P.m0 == EW_CGLS_INIT);
                if (is_init || sS[P.m1].active) finish_ew(sS[P.m1], P.m0, tot[0]);
            }
            if (!sS[0].active && !sS[1].active) *P.done_flag = 1;
        }
        __syncthreads();
        for (int i = tid; i < (int)(2 * sizeof(SlotState) / sizeof(double)); i += nt)
            reinterpret_cast<double *>(P.st)[i] = reinterpret_cast<const double *>(sS)[i];
    }
}

// ---- host side --------------------------------------------------------------------------------------
// FPSB_DIST_PROF=1: an event after every launch of the distributed loop; the mean time between consecutive
// events is printed per kind and rank at the end of the solve (tools/dist_bench.py; profiling only)
struct DistProf {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> tag;
    size_t used = 0;
    double last_us[4] = {0, 0, 0, 0};        // step_n, xchg_after_n, step_m, xchg_after_m of the last profiled solve
    long long last_cnt[4] = {0, 0, 0, 0};
    bool verbose = false;
    DistProf() { const char *e = getenv("FPSB_DIST_PROF"); on = verbose = e && *e && *e != '0'; }
    void mark(int t, cudaStream_t s) {
        if (!on || used >= 8192) return;
        if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); tag.push_back(0); }
        tag[used] = t;
        cudaEventRecord(ev[used++], s);
    }
    void report(int rank) {
        if (!on || used < 2) { used = 0; return; }
        const char *names[] = {"start", "step_n", "xchg_after_n", "step_m", "xchg_after_m", "other"};
        double sum[6] = {0, 0, 0, 0, 0, 0}; int cnt[6] = {0, 0, 0, 0, 0, 0};
        for (size_t i = 1; i < used; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            sum[tag[i]] += ms; cnt[tag[i]]++;
        }
        for (int t = 1; t < 5; ++t) { last_us[t - 1] = cnt[t] ? 1e3 * sum[t] / cnt[t] : 0.0; last_cnt[t - 1] = cnt[t]; }
        if (verbose) {
            fprintf(stderr, "[fpsb dist prof] rank %d:", rank);
            for (int t = 1; t < 6; ++t) if (cnt[t]) fprintf(stderr, "  %s %.1f us x %d", names[t], 1e3 * sum[t] / cnt[t], cnt[t]);
            fprintf(stderr, "\n");
        }
        used = 0;
    }
};
static DistProf g_dist_prof;
void dist_profile(bool on) { g_dist_prof.on = on || g_dist_prof.verbose; }
void dist_last_profile(double *us4, long long *cnt4) {
    for (int i = 0; i < 4; ++i) { us4[i] = g_dist_prof.last_us[i]; cnt4[i] = g_dist_prof.last_cnt[i]; }
}

struct DistEngine {
    Handle *h;
    DistCtx *D;
    NcclApi *api;
    Engine E;
    DistEngine(Handle *hh) : h(hh), D(hh->dist), api(nccl_api()), E(hh) { E.tot_out = D->tot.p; D->halo_fresh = false; }

    // one launch of the peer-memory exchange (see xchg_kernel)
    void xchg(bool do_scatter, bool do_epi, bool do_gather, int kind, int m0, int m1, const DistEpiParams *Q = nullptr) {
        XchgParams P{};
        P.nranks = D->nranks; P.rank = D->rank;
        P.do_scatter = do_scatter; P.do_epi = do_epi; P.do_gather = do_gather; P.kind = kind; P.m0 = m0; P.m1 = m1;
        P.sig = D->sig;
        P.sc_par = (int)(D->n_sc & 1); P.ga_par = (int)(D->n_ga & 1); P.tot_par = (int)(D->n_tot & 1);
        P.nsend = D->nsend; P.nrecv = D->nrecv;
        P.meta = reinterpret_cast<const long long *>(D->d_meta.p);
        P.send_idx = D->send_idx.p;
        P.S = D->S.p; P.pair = E.W->Gn.p;
        P.nbound = (int)D->nbound; P.bidx = D->bidx.p; P.bptr = D->bptr.p; P.bsrc = D->bsrc.p; P.bpeer = D->bpeer.p;
        P.xparts = D->xparts.p; P.bar = D->d_bar.p; P.bar_base = D->bar_base;
        if (Q) P.Q = *Q;
        P.tot = D->tot.p; P.st = E.W->st.p; P.done_flag = E.W->done.p; P.err = D->d_err.p;
        P.mine = D->mbox;
        for (int p = 0; p < D->nranks; ++p) {
            P.peer[p] = D->peer_mbox[p];
            P.sc_at_peer[p] = D->sc_at_peer[p]; P.ga_at_peer[p] = D->ga_at_peer[p];
            P.peer_nsend[p] = D->peer_nsend[p]; P.peer_nrecv[p] = D->peer_nrecv[p];
        }
        // the lists are short (two grid lines per neighbour for a strip partition): one entry per thread
        int64_t work = 0;
        if (do_scatter) work = std::max<int64_t>(work, std::max(D->nrecv, D->nbound));
        if (do_epi) work = std::max<int64_t>(work, D->nbound);
        if (do_gather) work = std::max<int64_t>(work, std::max(D->nsend, D->nrecv));
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(DistCtx::kXchgMaxGrid, (work + kBlock - 1) / kBlock));
        // (launching this kernel with programmatic dependent launch on both sides was measured: no gain)
        xchg_kernel<<<grid, kBlock, 0, h->stream>>>(P);
        h->launches += 1;
        D->bar_base += (do_scatter ? 2ull : 1ull) * (unsigned long long)grid;
        if (D->nranks > 1) {
            if (do_scatter) { D->sig += 1; D->n_sc += 1; }
            if (do_gather || kind >= 0) D->sig += 1;
        }
        if (do_gather) { D->n_ga += 1; D->halo_fresh = true; }
        if (kind >= 0) D->n_tot += 1;
    }
    void check_peer_error() {
        if (!D->peer) return;
        int e = 0;
        FPSB_CUDA(cudaMemcpyAsync(&e, D->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        FPSB_CUDA(cudaStreamSynchronize(h->stream));
        if (e) { set_error("row-partitioned run: a peer did not answer within 4 s (peer-memory exchange)"); throw CudaFail{FPSB_ECUDA}; }
    }

    void allreduce_finish(int kind, int m0, int m1) {
        if (D->peer) { xchg(false, false, false, kind, m0, m1); return; }
        FPSB_NCCL(api->AllReduce(D->tot.p, D->tot.p, 4, ncclDouble, ncclSum, D->comm, h->stream));
        finish_kernel<<<1, 64, 0, h->stream>>>(E.W->st.p, kind, m0, m1, D->tot.p, E.W->done.p);
        h->launches += 1;
    }
    // owners -> halo slots of an interleaved pair living in the extended n-space
    void gather_halo(double2 *pair) {
        if (D->nranks == 1) return;
        if (D->peer && pair == E.W->Gn.p) { xchg(false, false, true, -1, 0, 0); return; }
        if (D->nsend > 0) {
            pack_pairs_kernel<<<(unsigned)((D->nsend + 255) / 256), 256, 0, h->stream>>>((int)D->nsend, D->send_idx.p, pair, D->sendbuf.p);
            h->launches += 1;
        }
        FPSB_NCCL(api->GroupStart());
        for (int p = 0; p < D->nranks; ++p) {
            if (p == D->rank) continue;
            const int64_t ns = D->send_ptr[p + 1] - D->send_ptr[p];
            if (ns > 0) FPSB_NCCL(api->Send(D->sendbuf.p + D->send_ptr[p], (size_t)ns * 2, ncclDouble, p, D->comm, h->stream));
            if (D->recv_cnt[p] > 0) FPSB_NCCL(api->Recv(pair + D->recv_start[p], (size_t)D->recv_cnt[p] * 2, ncclDouble, p, D->comm, h->stream));
        }
        FPSB_NCCL(api->GroupEnd());
    }
    // halo partial sums -> owners, added in rank order
    void scatter_add(double2 *S) {
        if (D->nranks == 1) return;
        if (D->peer && S == D->S.p) { xchg(true, false, false, -1, 0, 0); return; }
        FPSB_NCCL(api->GroupStart());
        for (int p = 0; p < D->nranks; ++p) {
            if (p == D->rank) continue;
            const int64_t nr = D->send_ptr[p + 1] - D->send_ptr[p];
            if (D->recv_cnt[p] > 0) FPSB_NCCL(api->Send(S + D->recv_start[p], (size_t)D->recv_cnt[p] * 2, ncclDouble, p, D->comm, h->stream));
            if (nr > 0) FPSB_NCCL(api->Recv(D->recvbuf.p + D->send_ptr[p], (size_t)nr * 2, ncclDouble, p, D->comm, h->stream));
        }
        FPSB_NCCL(api->GroupEnd());
        for (int p = 0; p < D->nranks; ++p) {
            const int64_t nr = D->send_ptr[p + 1] - D->send_ptr[p];
            if (p == D->rank || nr == 0) continue;
            add_pairs_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, h->stream>>>((int)nr, D->send_idx.p + D->send_ptr[p],
                                                                                 D->recvbuf.p + D->send_ptr[p], S);
            h->launches += 1;
        }
    }
    // S = A_loc' [Gm.x Gm.y]  (raw sums of both columns, every extended row) followed by the exchange
    void jt_partials(const double2 *gm_pair) {
        StepParams P = E.base_n;
        P.io[0] = io_mode(MD_PLAIN); P.io[1] = io_mode(MD_PLAIN);
        P.gin2 = gm_pair; P.self2 = D->S.p;
        P.tot_out = nullptr; P.raw_out = nullptr;          // plain product: every row's sum goes to S through self2
        if (h->At.grid == 0) { D->S.zero(h->stream); return; }
        launch_step(h, h->At, P, true, 0);
        scatter_add(D->S.p);
    }
    // n-space half step.  Fused variant: the step kernel runs the epilogue of the interior owned rows
    // itself and leaves the raw sums of the halo / boundary rows; those are exchanged, added, and
    // finished by one small kernel.  Fallback (long rows in A', huge boundary): partial sums for every
    // row, exchange, separate epilogue over all owned rows.
    void step_n(const SlotIO &io0, const SlotIO &io1) {
        step_n_impl(io0, io1);
        g_dist_prof.mark(2, h->stream);
    }
    void step_n_impl(const SlotIO &io0, const SlotIO &io1) {
        DistEpiParams Q{};
        Q.n_own = (int)D->n_own; Q.own_off = (int)D->own_off;
        Q.S = D->S.p; Q.self2 = E.W->Gn.p;
        Q.io[0] = io0; Q.io[1] = io1;
        Q.st = E.W->st.p; Q.partials = E.W->partials.p; Q.counter = E.W->counter.p; Q.tot_out = D->tot.p;
        if (D->fused_n) {
            StepParams P = E.base_n;
            P.io[0] = io0; P.io[1] = io1;
            P.tot_out = D->tot.p; P.raw_out = D->S.p;
            P.st = E.W->st.p;
            if (h->At.grid == 0) return;
            launch_step(h, h->At, P, true, 1);
            g_dist_prof.mark(1, h->stream);
            if (D->peer) {
                // scatter-add + boundary epilogue + halo gather for the next m-space step + norms + recurrences
                xchg(D->nranks > 1, true, D->nranks > 1, 0, io0.mode, io1.mode, &Q);
                return;
            }
            scatter_add(D->S.p);
            if (D->nbound > 0) {
                boundary_epilogue_kernel<<<1, kBlock, 0, h->stream>>>((int)D->nbound, D->bidx.p, Q);
                h->launches += 1;
            }
        } else {
            jt_partials(E.W->Gm.p);
            const int grid = std::max(1, std::min(E.W->ew_grid, (int)((D->n_own + kBlock - 1) / kBlock)));
            dist_epilogue_kernel<<<grid, kBlock, 0, h->stream>>>(Q);
            h->launches += 1;
            D->halo_fresh = false;
        }
        allreduce_finish(0, io0.mode, io1.mode);
    }
    // m-space half step: halo gather, fused step kernel on A_loc (rows are local), all-reduce, recurrences
    void step_m(const SlotIO &io0, const SlotIO &io1) {
        if (!(D->peer && D->halo_fresh)) gather_halo(E.W->Gn.p);
        if (D->peer && D->tail && h->A.grid > 0 && h->A.nlong == 0) {
            // the step kernel's last CTA exchanges the four sums with the peers and runs the recurrences
            E.ptail = D->d_ptail.p; E.pt_sig = D->sig; E.pt_par = (int)(D->n_tot & 1);
            E.step(true, true, io0, io1);
            E.ptail = nullptr;
            if (D->nranks > 1) D->sig += 1;
            D->n_tot += 1;
            return;
        }
        E.step(true, true, io0, io1);
        g_dist_prof.mark(3, h->stream);
        allreduce_finish(0, io0.mode, io1.mode);
        g_dist_prof.mark(4, h->stream);
    }
    // The whole LSQR / CRAIG loop in the persistent kernel, exchanges included (gk_loop_kernel<true>); false when this
    // handle cannot (NCCL transport, long rows, an operator without tiles): the launch-per-half-iteration loop then runs
    bool run_loop(const SlotIO &n0, const SlotIO &n1, const SlotIO &m0, const SlotIO &m1) {
        if (!D->loop_ok || g_dist_prof.on) return false;     // (per-launch profiling measures the launch-per-half-iteration path)
        E.loop_dx = D->d_dloop.p; E.loop_raw = D->S.p;
        if (!E.can_persist()) { E.loop_dx = nullptr; return false; }
        unsigned long long seq[4] = {D->sig, D->n_sc, D->n_ga, D->n_tot};
        FPSB_CUDA(cudaMemcpyAsync(D->d_seq.p, seq, sizeof(seq), cudaMemcpyHostToDevice, h->stream));
        FPSB_CUDA(cudaStreamSynchronize(h->stream));
        E.run_loop(n0, n1, m0, m1);                     // (ends with the stream synchronised)
        FPSB_CUDA(cudaMemcpy(seq, D->d_seq.p, sizeof(seq), cudaMemcpyDeviceToHost));
        D->sig = seq[0]; D->n_sc = seq[1]; D->n_ga = seq[2]; D->n_tot = seq[3];
        D->halo_fresh = false;
        E.loop_dx = nullptr;
        return true;
    }
    void ew(int op, int slot, int n, const double *in0, double *v0, double *v1, double *v2, double2 *pair, int pair_slot, double c0,
            double *v3 = nullptr, double *v4 = nullptr) {
        E.ew(op, slot, n, in0, v0, v1, v2, v3, v4, pair, pair_slot, c0, 1);
        const bool npair = pair >= E.W->Gn.p && pair < E.W->Gn.p + h->nvar;      // the n-space pair changed
        if (npair) D->halo_fresh = false;
        allreduce_finish(1, op, slot);
    }
};

// plain distributed products (owned slices in / out); used by the parity tests of the exchange
void dist_jprod(Handle *h, const double *x_own, double *y_loc) {
    DistEngine X(h);
    IterWs *W = h->iter;
    DistCtx *D = h->dist;
    // stage x in column 0 of the gathered pair of the extended space
    W->Gn.zero(h->stream);
    X.E.tot_out = nullptr;
    X.E.ew(EW_INIT_LSQR, 0, (int)D->n_own, x_own, nullptr, nullptr, nullptr, nullptr, nullptr, W->Gn.p + D->own_off, 0, 1.0, 0);
    X.gather_halo(W->Gn.p);
    if (h->A.grid == 0) return;
    StepParams P = X.E.base_m;
    P.io[0] = io_mode(MD_PLAIN); P.io[1] = io_mode(MD_PLAIN);
    P.gin2 = W->Gn.p; P.self2 = W->Gm.p; P.tot_out = nullptr;
    launch_step(h, h->A, P, true, 0);
    unpack_cols_kernel<<<(unsigned)((h->ncon + 255) / 256), 256, 0, h->stream>>>((int)h->ncon, W->Gm.p, y_loc, nullptr);
    h->launches += 1;
    X.check_peer_error();
}
void dist_jtprod(Handle *h, const double *u_loc, double *y_own) {
    DistEngine X(h);
    IterWs *W = h->iter;
    DistCtx *D = h->dist;
    W->Gm.zero(h->stream);
    X.E.tot_out = nullptr;
    X.E.ew(EW_INIT_LSQR, 0, (int)h->ncon, u_loc, nullptr, nullptr, nullptr, nullptr, nullptr, W->Gm.p, 0, 1.0, 0);
    X.jt_partials(W->Gm.p);
    unpack_cols_kernel<<<(unsigned)((D->n_own + 255) / 256), 256, 0, h->stream>>>((int)D->n_own, D->S.p + D->own_off, y_own, nullptr);
    h->launches += 1;
    X.check_peer_error();
}

// p_i = rhs_i - (A' q_i) on the owned rows, for one or two columns
__global__ void residual_own_kernel(int n, const double2 *S, const double *rhs0, const double *rhs1, double *p0, double *p1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double2 s = S[i]; if (p0) p0[i] = rhs0[i] - s.x; if (p1) p1[i] = rhs1[i] - s.y; }
}

// Row-partitioned solve_two_mixed (src/solve_linear_system.jl:107-140): rhs1 / p1 / p2 are the OWNED
// n-space slices, rhs2 / q1 / q2 the local m-space slices
void dist_solve_two_mixed(Handle *h, double delta, const double *rhs1, const double *rhs2, double *p1, double *q1,
                          double *p2, double *q2, fpsb_krylov_stats *st, int64_t nvar_global, int64_t ncon_global) {
    DistEngine X(h);
    Engine &E = X.E;
    IterWs *W = h->iter;
    DistCtx *D = h->dist;
    const fpsb_iter_opts &o = h->iopts;
    const int64_t n = nvar_global, m = ncon_global;
    const int n_own = (int)D->n_own, m_loc = (int)h->ncon;
    E.begin(make_lsqr(sqrt(delta), o.ls_atol, o.ls_rtol, o.ls_itmax, n, m),
            make_craig(delta, o.ln_atol, o.ln_rtol, o.ln_btol, o.ln_conlim, o.ln_itmax, m, n));
    W->Gn.zero(h->stream); W->Gm.zero(h->stream);
    W->an[1][0].zero(h->stream); W->an[1][1].zero(h->stream);
    W->am[0][1].zero(h->stream);
    X.ew(EW_INIT_LSQR, 0, n_own, rhs1, nullptr, nullptr, nullptr, W->Gn.p + D->own_off, 0, 1.0);
    X.ew(EW_INIT_CRAIG, 1, m_loc, rhs2, W->am[1][0].p, W->am[1][1].p, nullptr, W->Gm.p, 1, -1.0);
    SlotIO l_init = io_mode(MD_LSQR_INIT_M), l_u = io_mode(MD_LSQR_U);
    SlotIO l_v = io_mode(MD_LSQR_V, W->am[0][0].p, W->am[0][1].p);
    SlotIO c_v = io_mode(MD_CRAIG_V, W->an[1][0].p, W->an[1][1].p);
    SlotIO c_u = io_mode(MD_CRAIG_U, W->am[1][0].p, W->am[1][1].p);
    E.mark_begin();
    X.step_m(l_init, io_none());
    g_dist_prof.mark(0, h->stream);
    if (!X.run_loop(l_u, c_v, l_v, c_u))
        E.loop([&](int) {
            X.step_n(l_u, c_v);
            X.step_m(l_v, c_u);
        }, kChunk);
    E.mark_end();
    g_dist_prof.report(D->rank);
    E.tot_out = nullptr;
    // outputs: q1 = x_lsqr ; p1 = rhs1 - A' q1 ; p2 = -(x_craig + pending) ; q2 = y_craig
    E.ew(EW_COPY, 0, m_loc, W->am[0][1].p, q1, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    E.ew(EW_COPY, 1, m_loc, W->am[1][1].p, q2, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    // the CRAIG flush reads the pending coefficients of the (replicated) state: owned rows only
    E.ew(EW_CRAIG_FLUSH, 1, n_own, nullptr, W->an[1][0].p + D->own_off, W->an[1][1].p + D->own_off, p2, nullptr, nullptr,
         W->Gn.p + D->own_off, 1, 1.0, 1);
    // A' q1 through the exchange (Gm column 0 <- q1 ; column 1 zero)
    W->Gm.zero(h->stream);
    E.ew(EW_INIT_LSQR, 0, m_loc, q1, nullptr, nullptr, nullptr, nullptr, nullptr, W->Gm.p, 0, 1.0, 0);
    X.jt_partials(W->Gm.p);
    residual_own_kernel<<<(unsigned)((n_own + 255) / 256), 256, 0, h->stream>>>(n_own, D->S.p + D->own_off, rhs1, nullptr, p1, nullptr);
    h->launches += 1;
    E.fetch(st);
    X.check_peer_error();
}

// Row-partitioned solve_two_least_squares (src/solve_linear_system.jl:79-105): two LSQR solves in lock step
void dist_solve_two_least_squares(Handle *h, double delta, const double *rhs1, const double *rhs2, double *p1, double *q1,
                                  double *p2, double *q2, fpsb_krylov_stats *st, int64_t nvar_global, int64_t ncon_global) {
    DistEngine X(h);
    Engine &E = X.E;
    IterWs *W = h->iter;
    DistCtx *D = h->dist;
    const fpsb_iter_opts &o = h->iopts;
    const int64_t n = nvar_global, m = ncon_global;
    const int n_own = (int)D->n_own, m_loc = (int)h->ncon;
    SlotState s = make_lsqr(sqrt(delta), o.ls_atol, o.ls_rtol, o.ls_itmax, n, m);
    E.begin(s, s);
    W->Gn.zero(h->stream); W->Gm.zero(h->stream);
    W->am[0][1].zero(h->stream); W->am[1][1].zero(h->stream);
    X.ew(EW_INIT_LSQR, 0, n_own, rhs1, nullptr, nullptr, nullptr, W->Gn.p + D->own_off, 0, 1.0);
    X.ew(EW_INIT_LSQR, 1, n_own, rhs2, nullptr, nullptr, nullptr, W->Gn.p + D->own_off, 1, 1.0);
    SlotIO init0 = io_mode(MD_LSQR_INIT_M), init1 = io_mode(MD_LSQR_INIT_M);
    SlotIO u0 = io_mode(MD_LSQR_U), u1 = io_mode(MD_LSQR_U);
    SlotIO v0 = io_mode(MD_LSQR_V, W->am[0][0].p, W->am[0][1].p);
    SlotIO v1 = io_mode(MD_LSQR_V, W->am[1][0].p, W->am[1][1].p);
    E.mark_begin();
    X.step_m(init0, init1);
    if (!X.run_loop(u0, u1, v0, v1))
        E.loop([&](int) {
            X.step_n(u0, u1);
            X.step_m(v0, v1);
        }, kChunk);
    E.mark_end();
    E.tot_out = nullptr;
    E.ew(EW_COPY, 0, m_loc, W->am[0][1].p, q1, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    E.ew(EW_COPY, 1, m_loc, W->am[1][1].p, q2, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    // p_i = rhs_i - A' q_i : one two-column product through the exchange
    {
        // interleave q1, q2 into Gm
        W->Gm.zero(h->stream);
        E.ew(EW_INIT_LSQR, 0, m_loc, q1, nullptr, nullptr, nullptr, nullptr, nullptr, W->Gm.p, 0, 1.0, 0);
        E.ew(EW_INIT_LSQR, 1, m_loc, q2, nullptr, nullptr, nullptr, nullptr, nullptr, W->Gm.p, 1, 1.0, 0);
    }
    X.jt_partials(W->Gm.p);
    residual_own_kernel<<<(unsigned)((n_own + 255) / 256), 256, 0, h->stream>>>(n_own, D->S.p + D->own_off, rhs1, rhs2, p1, p2);
    h->launches += 1;
    E.fetch(st);
    X.check_peer_error();
}

// one column of an interleaved pair -> the same column of another pair / a plain vector
__global__ void pair_set_col_kernel(int n, const double2 *src, double2 *dst, int col) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double2 v = src[i]; reinterpret_cast<double *>(dst + i)[col] = col ? v.y : v.x; }
}

// Row-partitioned solve_two_extras of IterativeSolver (src/solve_linear_system.jl:45-77):
//   u1 = LSQR(A', rhs1, lambda = sqrt(tau))          rhs1 = the OWNED n-space slice
//   u2 = MINRES(A A' + tau I, rhs2)                  rhs2 / u1 / u2 = the local m-space slices
// tau = max(delta, 1e-14).  LSQR is the one-slot case of dist_solve_two_least_squares (persistent loop kernel with the
// exchange inside when the handle allows it).  MINRES keeps all its vectors in the m-space (rows of A: local, no halo);
// only its operator needs the exchange: t = A' r2 is formed as partial sums of A_loc', scatter-added to the owners,
// gathered back into the halo slots, and the fused MINRES row epilogue then runs on A_loc t; the inner products are
// all-reduced after every step.
void dist_solve_two_extras(Handle *h, double delta, const double *rhs1, const double *rhs2, double *u1, double *u2,
                           fpsb_krylov_stats *st, int64_t nvar_global, int64_t ncon_global) {
    DistEngine X(h);
    Engine &E = X.E;
    IterWs *W = h->iter;
    DistCtx *D = h->dist;
    const fpsb_iter_opts &o = h->iopts;
    const int64_t n = nvar_global, m = ncon_global;
    const int n_own = (int)D->n_own, m_loc = (int)h->ncon, n_ext = (int)h->nvar;
    const double tau = std::max(delta, 1e-14);
    fpsb_krylov_stats tmp[2];
    // ---- LSQR(A', rhs1, lambda = sqrt(tau)) in slot 0
    E.begin(make_lsqr(sqrt(tau), o.ls_atol, o.ls_rtol, o.ls_itmax, n, m), make_none());
    W->Gn.zero(h->stream); W->Gm.zero(h->stream);
    W->am[0][1].zero(h->stream);
    X.ew(EW_INIT_LSQR, 0, n_own, rhs1, nullptr, nullptr, nullptr, W->Gn.p + D->own_off, 0, 1.0);
    {
        SlotIO l_init = io_mode(MD_LSQR_INIT_M), l_u = io_mode(MD_LSQR_U);
        SlotIO l_v = io_mode(MD_LSQR_V, W->am[0][0].p, W->am[0][1].p);
        X.step_m(l_init, io_none());
        if (!X.run_loop(l_u, io_none(), l_v, io_none()))
            E.loop([&](int) {
                X.step_n(l_u, io_none());
                X.step_m(l_v, io_none());
            }, kChunk);
    }
    E.tot_out = nullptr;
    E.ew(EW_COPY, 0, m_loc, W->am[0][1].p, u1, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    E.fetch(tmp);
    st[0] = tmp[0];
    // ---- MINRES(A A' + tau I, rhs2) in slot 1
    E.tot_out = D->tot.p;
    E.begin(make_none(), make_minres(tau, o.ne_atol, o.ne_rtol, o.ne_etol, o.ne_conlim, o.ne_itmax, m));
    const int slot = 1;
    double *r1 = W->am[slot][0].p, *r2 = W->am[slot][1].p, *wa = W->am[slot][2].p, *wb = W->am[slot][3].p,
           *x = W->am[slot][4].p, *y = W->ym.p, *t = W->an[slot][0].p;
    W->Gn.zero(h->stream); W->Gm.zero(h->stream);
    X.ew(EW_MINRES_INIT, slot, m_loc, rhs2, r1, r2, wa, nullptr, -1, 1.0, wb, x);
    double *w1 = wa, *w2 = wb;
    E.loop([&](int k) {
        // t = A' r2 on the owned + halo columns
        E.tot_out = nullptr;
        E.ew(EW_INIT_LSQR, slot, m_loc, r2, nullptr, nullptr, nullptr, nullptr, nullptr, W->Gm.p, slot, 1.0, 0);
        X.jt_partials(W->Gm.p);
        pair_set_col_kernel<<<(unsigned)((n_own + 255) / 256), 256, 0, h->stream>>>(n_own, D->S.p + D->own_off, W->Gn.p + D->own_off, slot);
        h->launches += 1;
        D->halo_fresh = false;
        X.gather_halo(W->Gn.p);
        unpack_cols_kernel<<<(unsigned)((n_ext + 255) / 256), 256, 0, h->stream>>>(n_ext, W->Gn.p, slot == 0 ? t : nullptr, slot == 1 ? t : nullptr);
        h->launches += 1;
        // y = (A t + lambda r2)/beta - (beta/oldbeta) r1 ; alpha = r2'y / beta
        E.tot_out = D->tot.p;
        SlotIO m_io = io_mode(MD_MINRES_M, r2, r1);
        m_io.gin = t; m_io.self = y;
        E.step(true, false, io_none(), m_io);
        X.allreduce_finish(0, MD_NONE, m_io.mode);
        X.ew(EW_MINRES_E1, slot, m_loc, y, r1, r2, w1, nullptr, -1, 1.0, w2, nullptr);
        double *wcur = (k == 1) ? w2 : w1;
        X.ew(EW_MINRES_E2, slot, m_loc, nullptr, nullptr, nullptr, wcur, nullptr, -1, 1.0, nullptr, x);
        if (k >= 2) std::swap(w1, w2);
    }, 2);
    E.tot_out = nullptr;
    E.ew(EW_COPY, slot, m_loc, x, u2, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    E.fetch(tmp);
    st[1] = tmp[1];
    X.check_peer_error();
}

