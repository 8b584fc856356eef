// Micro-benchmark: HBM read bandwidth of cp.async.bulk (1-D TMA) into a shared-memory ring versus
// plain vectorised LDG streaming, on B200.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t *bar, uint32_t ph) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(ph) : "memory");
    return ok;
}
// one producer thread streams `nchunk` chunks of `chunk` bytes (strided by grid) through `stages` ring
// slots; consumer warps just wait for each stage and release it (optionally touching the data).
__global__ void tma_stream(const unsigned char *src, size_t chunk, long nchunk, int stages, int split, int touch, double *sink) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ alignas(8) uint64_t full[16], empty[16];
    const int tid = threadIdx.x;
    const int ncons = blockDim.x / 32 - 1;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(ncons));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    double acc = 0;
    if (tid < 32) {
        if (tid == 0) {
            long k = 0;
            for (long c = blockIdx.x; c < nchunk; c += gridDim.x, ++k) {
                const int s = (int)(k % stages);
                if (k >= stages) while (!try_wait(&empty[s], (uint32_t)((k / stages - 1) & 1))) {}
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"((uint32_t)chunk) : "memory");
                const size_t piece = chunk / split;
                for (int j = 0; j < split; ++j)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + (size_t)s * chunk + j * piece)),
                                 "l"(src + (size_t)c * chunk + j * piece), "r"((uint32_t)piece), "r"(s32(&full[s])) : "memory");
            }
        }
    } else {
        long k = 0;
        const int lane = tid & 31;
        for (long c = blockIdx.x; c < nchunk; c += gridDim.x, ++k) {
            const int s = (int)(k % stages);
            while (!try_wait(&full[s], (uint32_t)((k / stages) & 1))) {}
            if (touch) {
                const double2 *p = reinterpret_cast<const double2 *>(ring + (size_t)s * chunk);
                for (size_t i = tid - 32; i < chunk / 16; i += blockDim.x - 32) { double2 v = p[i]; acc += v.x + v.y; }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
    }
    if (touch && acc == 1.2345) sink[0] = acc;
}

// NP producer warps (lane 0 of each issues), producer p owns ring slots k = p, p + NP, ...
__global__ void tma_stream_mp(const unsigned char *src, size_t chunk, long nchunk, int stages, int np, double *sink) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ alignas(8) uint64_t full[16], empty[16];
    const int tid = threadIdx.x;
    const int ncons = blockDim.x / 32 - np;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(ncons));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int wid = tid >> 5, lane = tid & 31;
    if (wid < np) {
        if (lane == 0) {
            long k = wid;
            for (long c = blockIdx.x + (long)wid * gridDim.x; c < nchunk; c += (long)gridDim.x * np, k += np) {
                const int s = (int)(k % stages);
                if (k >= stages) while (!try_wait(&empty[s], (uint32_t)((k / stages - 1) & 1))) {}
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"((uint32_t)chunk) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + (size_t)s * chunk)),
                             "l"(src + (size_t)c * chunk), "r"((uint32_t)chunk), "r"(s32(&full[s])) : "memory");
            }
        }
    } else {
        long k = 0;
        for (long c = blockIdx.x; c < nchunk; c += gridDim.x, ++k) {
            const int s = (int)(k % stages);
            while (!try_wait(&full[s], (uint32_t)((k / stages) & 1))) {}
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
    }
}
// LDG streaming: each thread reads 16 B, unroll U independent loads, grid-stride
template <int U>
__global__ void ldg_stream(const double2 *src, size_t n, double *sink) {
    double acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
    }
    if (acc == 1.2345) sink[0] = acc;
}
int main() {
    const size_t bytes = (size_t)512 << 20;
    unsigned char *d; double *sink;
    CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 1, bytes)); CK(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    auto timeit = [&](auto f, const char *name) {
        f(); CK(cudaDeviceSynchronize());
        float best = 1e9;
        for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
        printf("%-60s %8.3f ms %8.1f GB/s\n", name, best, bytes / best * 1e-6);
    };
    char nm[128];
    timeit([&] { ldg_stream<4><<<148 * 8, 256>>>((const double2 *)d, bytes / 16, sink); }, "ldg U=4 148x8x256");
    timeit([&] { ldg_stream<8><<<148 * 8, 256>>>((const double2 *)d, bytes / 16, sink); }, "ldg U=8 148x8x256");
    for (int touch = 0; touch < 0; ++touch)
    for (int cps = 1; cps <= 2; ++cps)
        for (size_t chunk : {(size_t)8192, (size_t)16384, (size_t)32768}) {
            for (int stages : {2, 4, 6}) {
                for (int split : {1, 4}) {
                    if ((size_t)stages * chunk * cps > 200 * 1024) continue;
                    snprintf(nm, sizeof nm, "tma touch=%d ctas/sm=%d chunk=%zu stages=%d split=%d", touch, cps, chunk, stages, split);
                    timeit([&] { tma_stream<<<148 * cps, 160, stages * chunk>>>(d, chunk, (long)(bytes / chunk), stages, split, touch, sink); }, nm);
                }
            }
        }
    CK(cudaFuncSetAttribute(tma_stream_mp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int ncons : {1, 2, 4, 8, 16, 24})
        for (size_t chunk : {(size_t)16384, (size_t)32768}) {
            const int np = 1, stages = 5;
            snprintf(nm, sizeof nm, "tma_mp producers=%d consumers(polling warps)=%d chunk=%zu stages=%d", np, ncons, chunk, stages);
            timeit([&] { tma_stream_mp<<<148, 32 * (np + ncons), stages * chunk>>>(d, chunk, (long)(bytes / chunk), stages, np, sink); }, nm);
        }
    CK(cudaDeviceSynchronize());
    return 0;
}
