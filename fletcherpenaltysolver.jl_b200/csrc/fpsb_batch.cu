// fpsb_batch.cu — throughput mode: thousands of independent SMALL instances of the 2-RHS solve
// (BASELINE config C5; SURVEY §8e "batch of independent instances: static sharding, zero
// communication").  Every instance is the LDLtSolver path of the reference
// (src/solve_linear_system.jl:206-252 with the natural ordering P = 1..N): dense LDL' of
//     K = [I A'; A -delta I]   (N = nvar + ncon <= 32)
// with LDLFactorizations' dynamic regularisation rule, then the two permuted L / D / L' sweeps.
// One thread per instance, K kept in local memory; instances are laid out instance-major so a warp
// reads 32 consecutive instances' data.  Multi-GPU: the caller shards instances across ranks.
#include "fpsb_internal.h"
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

__global__ void __launch_bounds__(128) batch_kkt_kernel(BatchParams P) {
    const int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= P.ninst) return;
    const int n = P.nvar, m = P.ncon, N = n + m;
    double L[kBatchMaxN * kBatchMaxN];     // lower triangle, row-major [i][j], j <= i
    double D[kBatchMaxN];
    double y0[kBatchMaxN], y1[kBatchMaxN];
    // assemble K (lower triangle): I, A, -delta I
    for (int i = 0; i < N; ++i)
        for (int j = 0; j <= i; ++j) L[i * kBatchMaxN + j] = 0.0;
    for (int i = 0; i < n; ++i) L[i * kBatchMaxN + i] = 1.0;
    const double *A = P.A + inst * (int64_t)m * n;
    for (int r = 0; r < m; ++r) {
        for (int c = 0; c < n; ++c) L[(n + r) * kBatchMaxN + c] = A[r * n + c];
        L[(n + r) * kBatchMaxN + n + r] = -P.delta;
    }
    // up-looking LDL' (row k of L from the rows above), reference pivot rule
    const bool dyn = (P.r1 != 0.0) || (P.r2 != 0.0);
    bool ok = true;
    for (int k = 0; k < N; ++k) {
        double dk = L[k * kBatchMaxN + k];
        for (int i = 0; i < k; ++i) {
            // y_i = K[k][i] - sum_{j<i} y_j L[i][j]
            double yi = L[k * kBatchMaxN + i];
            for (int j = 0; j < i; ++j) yi -= y0[j] * L[i * kBatchMaxN + j];
            y0[i] = yi;
        }
        for (int i = 0; i < k; ++i) {
            const double l = y0[i] / D[i];
            dk -= l * y0[i];
            L[k * kBatchMaxN + i] = l;
        }
        if (dyn && fabs(dk) < P.tol) {
            const double r = (k < n) ? P.r1 : P.r2;
            const double sg = (double)((r > 0.0) - (r < 0.0));
            dk = sg * fmax(fabs(dk + r), fabs(r));
        }
        if (dk == 0.0) { ok = false; dk = 1.0; }
        D[k] = dk;
    }
    P.factorized[inst] = ok ? 1 : 0;
    const double *b1 = P.rhs1 + inst * (int64_t)n;
    const double *b2 = P.rhs2 + inst * (int64_t)(P.kind == 0 ? m : n);
    for (int i = 0; i < N; ++i) {
        y0[i] = (i < n) ? b1[i] : 0.0;
        y1[i] = (P.kind == 0) ? ((i < n) ? 0.0 : b2[i - n]) : ((i < n) ? b2[i] : 0.0);
    }
    if (ok) {
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < i; ++j) { y0[i] -= L[i * kBatchMaxN + j] * y0[j]; y1[i] -= L[i * kBatchMaxN + j] * y1[j]; }
        for (int i = 0; i < N; ++i) { y0[i] /= D[i]; y1[i] /= D[i]; }
        for (int j = N - 1; j >= 0; --j)
            for (int i = j + 1; i < N; ++i) { y0[j] -= L[i * kBatchMaxN + j] * y0[i]; y1[j] -= L[i * kBatchMaxN + j] * y1[i]; }
    }
    for (int i = 0; i < n; ++i) { P.p1[inst * n + i] = y0[i]; P.p2[inst * n + i] = y1[i]; }
    for (int i = 0; i < m; ++i) { P.q1[inst * m + i] = y0[n + i]; P.q2[inst * m + i] = y1[n + i]; }
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
        DevBuf<double> dA, d1, d2, dout;
        DevBuf<int> dfac;
        const size_t nout = ni * (2 * n + 2 * m);
        if (loc == FPSB_HOST) {
            dA.alloc(ni * m * n + 1); d1.alloc(ni * n + 1); d2.alloc(ni * n2 + 1); dout.alloc(nout + 1); dfac.alloc(ni + 1);
            FPSB_CUDA(cudaMemcpy(dA.p, A, ni * m * n * sizeof(double), cudaMemcpyHostToDevice));
            FPSB_CUDA(cudaMemcpy(d1.p, rhs1, ni * n * sizeof(double), cudaMemcpyHostToDevice));
            FPSB_CUDA(cudaMemcpy(d2.p, rhs2, ni * n2 * sizeof(double), cudaMemcpyHostToDevice));
            P.A = dA.p; P.rhs1 = d1.p; P.rhs2 = d2.p;
            P.p1 = dout.p; P.q1 = P.p1 + ni * n; P.p2 = P.q1 + ni * m; P.q2 = P.p2 + ni * n;
            P.factorized = dfac.p;
        } else {
            P.A = A; P.rhs1 = rhs1; P.rhs2 = rhs2; P.p1 = p1; P.q1 = q1; P.p2 = p2; P.q2 = q2; P.factorized = factorized;
        }
        const int grid = (int)((ninst + 127) / 128);
        batch_kkt_kernel<<<grid, 128>>>(P);
        FPSB_CUDA(cudaGetLastError());
        FPSB_CUDA(cudaDeviceSynchronize());
        if (loc == FPSB_HOST) {
            FPSB_CUDA(cudaMemcpy(p1, P.p1, ni * n * sizeof(double), cudaMemcpyDeviceToHost));
            FPSB_CUDA(cudaMemcpy(q1, P.q1, ni * m * sizeof(double), cudaMemcpyDeviceToHost));
            FPSB_CUDA(cudaMemcpy(p2, P.p2, ni * n * sizeof(double), cudaMemcpyDeviceToHost));
            FPSB_CUDA(cudaMemcpy(q2, P.q2, ni * m * sizeof(double), cudaMemcpyDeviceToHost));
            FPSB_CUDA(cudaMemcpy(factorized, dfac.p, ni * sizeof(int), cudaMemcpyDeviceToHost));
        }
        return FPSB_OK;
    } catch (const fpsb::CudaFail &f) {
        return f.code;
    } catch (...) {
        set_error("fpsb_batch_solve_two: unexpected exception");
        return FPSB_ECUDA;
    }
}
