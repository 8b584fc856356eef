// fpsb_device.cuh — sm_100a device helpers: TMA bulk copies + mbarrier, reductions, Givens.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fpsb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a mis-programmed transaction count must not hang the GPU box
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); ++it)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
// A parity wait is only meaningful while the waiter is at most ONE phase ahead of the barrier.  The ring's stages are
// consumed by rotating consumer groups (stage count and group count are coprime in general), so a group that runs two
// tiles ahead of its neighbours could ask for round r of a stage whose round r - 1 has not even been filled: the parity
// test then passes at once and the group reads the stage's OLD contents (measured: wrong results at n = 450-550 K with
// early row sums, where two groups get a head start of a tile at every phase boundary).  So the first warp of a group
// publishes, per stage, the number of rounds it has SEEN filled (stage_seen, right after its own full-barrier wait), and
// a warp only parity-waits for round r once round r - 1 was seen (stage_turn_wait): one broadcast shared-memory load per
// tile and warp, one store per tile and group.  The chain only points backwards in the ring, so it cannot deadlock.
__device__ __forceinline__ void stage_turn_wait(const int *seen, int round) {     // every lane (same address: one broadcast)
    int it = 0;
    while (*reinterpret_cast<const volatile int *>(seen) < round) {
        if (++it > (1 << 24)) __trap();          // ~0.3 s: a mis-counted ring must end the launch, not hang the box
    }
}
__device__ __forceinline__ void stage_seen(int *seen, int round) {                // one lane of the group's first warp
    *reinterpret_cast<volatile int *>(seen) = round + 1;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// L2 eviction-priority policies (createpolicy): the matrix stream is evict-first, the gathered
// vectors evict-last, so the stream cannot push the gather working set out of the 126 MB L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar,
                                                  uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ double2 ldg_evict_last(const double2 *p, uint64_t policy) {
    double2 r;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ double ldg_evict_last(const double *p, uint64_t policy) {
    double r;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ void stg_evict_last(double2 *p, double2 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ void stg_evict_last(double *p, double v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fixed-tree block reduction of NV values per thread (<= 32 warps); result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch /* NV*32 */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nw = (int)((blockDim.x + 31) >> 5);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = warp_sum(v[k]);
        if (lane == 0) scratch[k * 32 + w] = x;
    }
    __syncthreads();
    if (w == 0) {
        // second level: one lane per warp, same shuffle tree
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = warp_sum(lane < nw ? scratch[k * 32 + lane] : 0.0);
    }
    __syncthreads();
}

__device__ __forceinline__ double sgn(double a) { return (double)((a > 0.0) - (a < 0.0)); }

// Krylov.jl sym_givens
__device__ __forceinline__ void sym_givens(double a, double b, double &c, double &s, double &rho) {
    if (b == 0.0) {
        c = (a == 0.0) ? 1.0 : sgn(a);
        s = 0.0;
        rho = fabs(a);
    } else if (a == 0.0) {
        c = 0.0;
        s = sgn(b);
        rho = fabs(b);
    } else if (fabs(b) > fabs(a)) {
        double t = a / b;
        s = sgn(b) / sqrt(1.0 + t * t);
        c = s * t;
        rho = b / s;
    } else {
        double t = b / a;
        c = sgn(a) / sqrt(1.0 + t * t);
        s = c * t;
        rho = a / c;
    }
}

}  // namespace fpsb
