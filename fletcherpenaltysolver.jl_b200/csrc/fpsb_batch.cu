// fpsb_batch.cu — throughput mode: thousands of independent SMALL instances of the 2-RHS solve
// (BASELINE config C5; SURVEY §8e "batch of independent instances: static sharding, zero
// communication").  Every instance is the LDLtSolver path of the reference
// (src/solve_linear_system.jl:206-252 with the natural ordering P = 1..N): dense LDL' of
//     K = [I A'; A -delta I]   (N = nvar + ncon <= 32)
// with LDLFactorizations' dynamic regularisation rule, then the two permuted L / D / L' sweeps.
// One warp per instance, L in shared memory (see the kernel).  Multi-GPU: the caller shards the
// instances across ranks.
#include "fpsb_internal.h"
#include <algorithm>
#include <cmath>

namespace fpsb {

constexpr int kBatchMaxN = 32;

struct BatchParams {
    int64_t ninst;
    int nvar, ncon;
    const double *A;       // [ninst][ncon][nvar]
    double delta;
    const double *rhs1;    // [ninst][nvar]
    const double *rhs2;    // [ninst][ncon] (mixed) or [ninst][nvar] (least squares)
    double *p1, *q1, *p2, *q2;
    int *factorized;       // [ninst]
    int kind;              // 0 mixed, 1 least squares
    double tol, r1, r2;
};

// One WARP per instance, lane == row of K.  L lives in shared memory (row-major lower triangle),
// D and the two right-hand sides in registers (lane i holds D[i], y0[i], y1[i]).  The operation
// order of every accumulated quantity is exactly the sequential up-looking algorithm's (ascending
// index, no FMA contraction, file compiled with -fmad=false), so results match the oracle's scalar
// LDL' bit for bit also when the dynamic regularisation amplifies rounding by 1/sqrt(eps).
constexpr int kBatchWarps = 8;

__global__ void __launch_bounds__(kBatchWarps * 32) batch_kkt_kernel(BatchParams P) {
    extern __shared__ double s_L[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = P.nvar, m = P.ncon, N = n + m;
    const int ld = N | 1;                                   // odd leading dimension: conflict-free column reads
    double *L = s_L + (size_t)wid * N * ld;
    const unsigned full = 0xffffffffu;
    const bool dyn = (P.r1 != 0.0) || (P.r2 != 0.0);
    for (int64_t inst = (int64_t)blockIdx.x * kBatchWarps + wid; inst < P.ninst; inst += (int64_t)gridDim.x * kBatchWarps) {
        // assemble the lower triangle of K = [I A'; A -delta I]
        const double *A = P.A + inst * (int64_t)m * n;
        for (int e = lane; e < N * ld; e += 32) L[e] = 0.0;
        __syncwarp();
        if (lane < n) L[lane * ld + lane] = 1.0;
        for (int e = lane; e < m * n; e += 32) { const int r = e / n, c = e - r * n; L[(n + r) * ld + c] = A[e]; }
        if (lane < m) L[(n + lane) * ld + n + lane] = -P.delta;
        __syncwarp();
        // up-looking LDL': row k of L from the rows above, reference pivot rule
        double Dl = 1.0;                                    // D[lane]
        bool ok = true;
        for (int k = 0; k < N; ++k) {
            double y = (lane < k) ? L[k * ld + lane] : 0.0; // K[k][lane]
            double dk = L[k * ld + k];
            for (int i = 0; i < k; ++i) {
                const double yi = __shfl_sync(full, y, i);  // Y[i] is final: all i' < i already applied
                if (lane > i && lane < k) y = y - L[lane * ld + i] * yi;
            }
            const double l = (lane < k) ? y / Dl : 0.0;
            for (int i = 0; i < k; ++i) {
                const double li = __shfl_sync(full, l, i), yi = __shfl_sync(full, y, i);
                dk = dk - li * yi;
            }
            if (dyn && fabs(dk) < P.tol) {
                const double r = (k < n) ? P.r1 : P.r2;
                const double sg = (double)((r > 0.0) - (r < 0.0));
                dk = sg * fmax(fabs(dk + r), fabs(r));
            }
            if (dk == 0.0) { ok = false; dk = 1.0; }
            __syncwarp();
            if (lane < k) L[k * ld + lane] = l;
            if (lane == k) Dl = dk;
            __syncwarp();
        }
        if (lane == 0) P.factorized[inst] = ok ? 1 : 0;
        // two right-hand sides
        const double *b1 = P.rhs1 + inst * (int64_t)n;
        const double *b2 = P.rhs2 + inst * (int64_t)(P.kind == 0 ? m : n);
        double y0 = 0.0, y1 = 0.0;
        if (lane < N) {
            y0 = (lane < n) ? b1[lane] : 0.0;
            y1 = (P.kind == 0) ? ((lane < n) ? 0.0 : b2[lane - n]) : ((lane < n) ? b2[lane] : 0.0);
        }
        if (ok) {
            for (int j = 0; j < N; ++j) {                   // forward: lane i accumulates over ascending j
                const double a = __shfl_sync(full, y0, j), b = __shfl_sync(full, y1, j);
                if (lane > j && lane < N) { const double lij = L[lane * ld + j]; y0 = y0 - lij * a; y1 = y1 - lij * b; }
            }
            if (lane < N) { y0 = y0 / Dl; y1 = y1 / Dl; }
            for (int j = N - 1; j >= 0; --j) {              // backward: lane j accumulates over ascending i > j
                for (int i = j + 1; i < N; ++i) {
                    const double a = __shfl_sync(full, y0, i), b = __shfl_sync(full, y1, i);
                    if (lane == j) { const double lij = L[i * ld + j]; y0 = y0 - lij * a; y1 = y1 - lij * b; }
                }
            }
        }
        if (lane < n) { P.p1[inst * n + lane] = y0; P.p2[inst * n + lane] = y1; }
        else if (lane < N) { P.q1[inst * m + lane - n] = y0; P.q2[inst * m + lane - n] = y1; }
        __syncwarp();
    }
}

// device workspace reused across calls (host-buffer callers): growing, never shrinking
struct BatchWs {
    DevBuf<double> in, out;
    DevBuf<int> fac;
    size_t in_cap = 0, out_cap = 0, fac_cap = 0;
    int device = -1;
};
static BatchWs &batch_ws(int device) {
    static thread_local BatchWs ws;
    if (ws.device != device) { ws.in.release(); ws.out.release(); ws.fac.release(); ws.in_cap = ws.out_cap = ws.fac_cap = 0; ws.device = device; }
    return ws;
}

}  // namespace fpsb

using namespace fpsb;

extern "C" int fpsb_batch_solve_two(int64_t ninst, int nvar, int ncon, int kind, const double *A, double delta,
                                    const double *rhs1, const double *rhs2, double *p1, double *q1, double *p2,
                                    double *q2, int *factorized, const fpsb_ldlt_opts *opts, int loc, int device) {
    if (ninst < 0 || nvar < 0 || ncon < 0 || nvar + ncon > kBatchMaxN || (kind != 0 && kind != 1)) {
        set_error("fpsb_batch_solve_two: bad sizes (nvar + ncon must be <= %d)", kBatchMaxN);
        return FPSB_EINVAL;
    }
    if (ninst == 0) return FPSB_OK;
    if (!A || !rhs1 || !rhs2 || !p1 || !q1 || !p2 || !q2 || !factorized) { set_error("fpsb_batch_solve_two: NULL argument"); return FPSB_EINVAL; }
    if (fpsb_device_count() <= 0) { set_error("fpsb_batch_solve_two: no CUDA device available (no CPU fallback)"); return FPSB_ECUDA; }
    try {
        FPSB_CUDA(cudaSetDevice(device));
        fpsb_ldlt_opts o;
        if (opts) o = *opts; else fpsb_ldlt_default_opts(&o);
        const size_t n = (size_t)nvar, m = (size_t)ncon, ni = (size_t)ninst;
        const size_t n2 = kind == 0 ? m : n;
        BatchParams P{};
        P.ninst = ninst; P.nvar = nvar; P.ncon = ncon; P.delta = delta; P.kind = kind;
        P.tol = o.ldlt_tol; P.r1 = o.ldlt_r1; P.r2 = o.ldlt_r2;
        const size_t nout = ni * (2 * n + 2 * m);
        const size_t nin = ni * m * n + ni * n + ni * n2;
        BatchWs &W = batch_ws(device);
        if (loc == FPSB_HOST) {
            if (W.in_cap < nin) { W.in.alloc(nin + 8); W.in_cap = nin; }
            if (W.out_cap < nout) { W.out.alloc(nout + 8); W.out_cap = nout; }
            if (W.fac_cap < ni) { W.fac.alloc(ni + 8); W.fac_cap = ni; }
            double *dA = W.in.p, *d1 = dA + ni * m * n, *d2 = d1 + ni * n;
            FPSB_CUDA(cudaMemcpyAsync(dA, A, ni * m * n * sizeof(double), cudaMemcpyHostToDevice, 0));
            FPSB_CUDA(cudaMemcpyAsync(d1, rhs1, ni * n * sizeof(double), cudaMemcpyHostToDevice, 0));
            FPSB_CUDA(cudaMemcpyAsync(d2, rhs2, ni * n2 * sizeof(double), cudaMemcpyHostToDevice, 0));
            P.A = dA; P.rhs1 = d1; P.rhs2 = d2;
            P.p1 = W.out.p; P.q1 = P.p1 + ni * n; P.p2 = P.q1 + ni * m; P.q2 = P.p2 + ni * n;
            P.factorized = W.fac.p;
        } else {
            P.A = A; P.rhs1 = rhs1; P.rhs2 = rhs2; P.p1 = p1; P.q1 = q1; P.p2 = p2; P.q2 = q2; P.factorized = factorized;
        }
        const int N = nvar + ncon;
        const size_t smem = (size_t)kBatchWarps * N * (N | 1) * sizeof(double);
        FPSB_CUDA(cudaFuncSetAttribute(batch_kkt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const int64_t want = (ninst + kBatchWarps - 1) / kBatchWarps;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
        batch_kkt_kernel<<<grid, kBatchWarps * 32, smem>>>(P);
        FPSB_CUDA(cudaGetLastError());
        if (loc == FPSB_HOST) {
            FPSB_CUDA(cudaMemcpyAsync(p1, P.p1, ni * n * sizeof(double), cudaMemcpyDeviceToHost, 0));
            FPSB_CUDA(cudaMemcpyAsync(q1, P.q1, ni * m * sizeof(double), cudaMemcpyDeviceToHost, 0));
            FPSB_CUDA(cudaMemcpyAsync(p2, P.p2, ni * n * sizeof(double), cudaMemcpyDeviceToHost, 0));
            FPSB_CUDA(cudaMemcpyAsync(q2, P.q2, ni * m * sizeof(double), cudaMemcpyDeviceToHost, 0));
            FPSB_CUDA(cudaMemcpyAsync(factorized, W.fac.p, ni * sizeof(int), cudaMemcpyDeviceToHost, 0));
        }
        FPSB_CUDA(cudaDeviceSynchronize());
        return FPSB_OK;
    } catch (const fpsb::CudaFail &f) {
        return f.code;
    } catch (...) {
        set_error("fpsb_batch_solve_two: unexpected exception");
        return FPSB_ECUDA;
    }
}
