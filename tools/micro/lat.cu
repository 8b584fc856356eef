// Micro-benchmark: dependent-issue latency of DFMA, LDS.128 and a SpMM-like inner loop from shared memory
// with 1..16 warps per SM, on B200.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
__global__ void dfma_chain(double *out, int n, long long *cyc) {
    double a = out[threadIdx.x], b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lds_chain(int *out, int n, long long *cyc) {
    __shared__ int4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_int4((i * 7 + 1) & 1023, 0, 0, 0);
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) idx = buf[idx].x;
    long long t1 = clock64();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// SpMM-like: per lane, npair pair rows from smem (val double2 + col int2), 2 gathers of double2, 4 DFMA
__global__ void spmm_like(double *out, int npair, int reps, long long *cyc) {
    extern __shared__ __align__(16) unsigned char sm[];
    double2 *val = (double2 *)sm;                     // [nw][npair][32]
    int2 *col = (int2 *)(sm + (size_t)(blockDim.x / 32) * npair * 32 * 16);
    double2 *win = (double2 *)(sm + (size_t)(blockDim.x / 32) * npair * 32 * 24);   // 1024 entries
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (blockDim.x / 32) * npair * 32; i += blockDim.x) { val[i] = make_double2(1.0, 0.5); col[i] = make_int2((i * 13) & 1023, (i * 29 + 5) & 1023); }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) win[i] = make_double2(1.0, 2.0);
    __syncthreads();
    double s0 = 0, s1 = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const double2 *sv = val + (size_t)w * npair * 32 + lane;
        const int2 *sc = col + (size_t)w * npair * 32 + lane;
#pragma unroll 5
        for (int p = 0; p < npair; ++p) {
            const double2 v = sv[p * 32];
            const int2 c = sc[p * 32];
            const double2 x0 = win[c.x], x1 = win[c.y];
            s0 = fma(v.x, x0.x, s0); s1 = fma(v.x, x0.y, s1);
            s0 = fma(v.y, x1.x, s0); s1 = fma(v.y, x1.y, s1);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = s0 + s1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *d; long long *cyc; int *di;
    CK(cudaMalloc(&d, 1 << 20)); CK(cudaMemset(d, 0, 1 << 20)); CK(cudaMalloc(&cyc, 1 << 16)); CK(cudaMalloc(&di, 1 << 20));
    long long h[4];
    const int n = 4096;
    dfma_chain<<<1, 32>>>(d, n, cyc); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("DFMA dependent latency        %.1f cycles\n", (double)h[0] / n);
    lds_chain<<<1, 32>>>(di, n, cyc); CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("LDS.128 dependent latency     %.1f cycles\n", (double)h[0] / n);
    CK(cudaFuncSetAttribute(spmm_like, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int nw : {1, 2, 4, 8, 12, 16, 24}) {
        const int npair = 10, reps = 200;
        size_t smem = (size_t)nw * npair * 32 * 24 + 1024 * 16;
        spmm_like<<<148, nw * 32, smem>>>(d, npair, reps, cyc); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("spmm-like inner loop, %2d warps/SM: %.0f cycles per slice of %d pair rows  -> %.2f cycles per warp-row per SM\n", nw,
               (double)h[0] / reps, npair, (double)h[0] / reps / (2.0 * npair * nw));
    }
    return 0;
}
