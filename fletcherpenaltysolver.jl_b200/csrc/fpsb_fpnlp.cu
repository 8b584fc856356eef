// fpsb_fpnlp.cu — device-resident evaluation glue of FletcherPenaltyNLP: the vector
// combinations either side of the 2-RHS solves as fused kernels on the handle's stream, so that x, g,
// c, the multipliers and the penalty gradient never leave HBM across the outer algo.jl loop.
//
// Reference code replaced (file:line under /root/reference):
//   _compute_ys_gs!   gs = p1 + sigma p2, ys = q1 + sigma q2, v = p2, w = q2   src/model-Fletcherpenaltynlp.jl:244-248
//                     memo key hash(x)                                          :235
//   obj               f - c'ys + rho/2 |c|^2 (+ eta/2 |x - xk|^2)               :364-367
//   grad!             gs - Hsv + sigma v + Sstw (+ rho J'c) (+ eta (x - xk))    :382-398
//   hprod! (Val 2)    Ptv = v - p1 ; Hv = p2 - HsPtv + 2 sigma Ptv (+ Hcv + rho J'Jv) (+ eta v), times obj_weight   :543-568
// All pointers are DEVICE pointers; scalar results come back through a pinned host slot.
#include "fpsb_internal.h"
#include "fpsb_device.cuh"

namespace fpsb {

__global__ void fp_ys_gs_kernel(int n, int m, double sigma, const double *p1, const double *q1, const double *p2,
                                const double *q2, double *gs, double *ys, double *v, double *w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double b = p2[i]; gs[i] = p1[i] + sigma * b; v[i] = b; }
    if (i < m) { const double b = q2[i]; ys[i] = q1[i] + sigma * b; w[i] = b; }
}

// fixed-order two-level reduction of up to 3 sums: out[0..2] (device) ; the last CTA finishes
__global__ void __launch_bounds__(256) fp_obj_kernel(int m, int n, const double *c, const double *ys, const double *x,
                                                      const double *xk, double *partials, unsigned *counter, double *out) {
    __shared__ double s_red[3 * 32];
    __shared__ int s_last;
    double acc[3] = {0.0, 0.0, 0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) { const double ci = c[i]; acc[0] += ci * ys[i]; acc[1] += ci * ci; }
    if (xk != nullptr)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { const double d = x[i] - xk[i]; acc[2] += d * d; }
    block_sum<3>(acc, s_red);
    if (threadIdx.x == 0) {
        double *pp = partials + (size_t)blockIdx.x * 3;
        pp[0] = acc[0]; pp[1] = acc[1]; pp[2] = acc[2];
        __threadfence();
        s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) {
        tot[0] += __ldcg(partials + (size_t)i * 3); tot[1] += __ldcg(partials + (size_t)i * 3 + 1); tot[2] += __ldcg(partials + (size_t)i * 3 + 2);
    }
    block_sum<3>(tot, s_red);
    if (threadIdx.x == 0) { out[0] = tot[0]; out[1] = tot[1]; out[2] = tot[2]; *counter = 0; }
}

__global__ void fp_grad_kernel(int n, double sigma, double rho, double eta, const double *gs, const double *Hsv,
                               const double *v, const double *Sstw, const double *Jtc, const double *x, const double *xk,
                               double *g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = gs[i] - Hsv[i] + sigma * v[i] + Sstw[i];
    if (rho > 0.0 && Jtc != nullptr) r = r + rho * Jtc[i];
    if (eta > 0.0 && xk != nullptr) r = r + eta * (x[i] - xk[i]);
    g[i] = r;
}

__global__ void fp_ptv_kernel(int n, const double *v, const double *p1, double *Ptv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Ptv[i] = v[i] - p1[i];
}

__global__ void fp_hprod2_kernel(int n, double sigma, double rho, double eta, double obj_weight, const double *p2,
                                 const double *HsPtv, const double *Ptv, const double *Hcv, const double *JtJv,
                                 const double *v, double *Hv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = p2[i] - HsPtv[i] + 2.0 * sigma * Ptv[i];
    if (rho > 0.0 && Hcv != nullptr && JtJv != nullptr) r = r + Hcv[i] + rho * JtJv[i];
    if (eta > 0.0) r = r + eta * v[i];
    Hv[i] = obj_weight * r;
}

// Val(1): the two extra terms of the exact Hessian (src/model-Fletcherpenaltynlp.jl:572-634), same operation order as the host mirror
__global__ void fp_hprod1_kernel(int n, double sigma, double rho, double eta, double obj_weight, const double *p2,
                                 const double *HsPtv, const double *Ptv, const double *JtinvJtJSsv, const double *SsinvJtJJv,
                                 const double *Hcv, const double *JtJv, const double *v, double *Hv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = p2[i] - HsPtv[i] + 2.0 * sigma * Ptv[i] - JtinvJtJSsv[i];
    r = r - SsinvJtJJv[i];
    if (rho > 0.0 && Hcv != nullptr && JtJv != nullptr) r = r + rho * (Hcv[i] + JtJv[i]);
    if (eta > 0.0) r = r + eta * v[i];
    Hv[i] = obj_weight * r;
}

// memo key of x: order-independent sum of per-element mixes of (bit pattern, index) — any 64-bit key
// that changes when x changes serves the reference's purpose (it memoises on hash(x) alone)
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z ^= z >> 33; z *= 0xff51afd7ed558ccdULL; z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ULL; z ^= z >> 33;
    return z;
}
__global__ void __launch_bounds__(256) fp_hash_kernel(int n, const double *x, unsigned long long *out) {
    unsigned long long acc = 0;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc += mix64((unsigned long long)__double_as_longlong(x[i]) ^ (0x9e3779b97f4a7c15ULL * (unsigned long long)(i + 1)));
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);      // integer addition: order-independent, exact
}

struct FpWs {
    DevBuf<double> partials, out;
    DevBuf<unsigned> counter;
    DevBuf<unsigned long long> key;
    double *h_out = nullptr;               // pinned: 3 doubles + 1 u64
};

static FpWs *fp_ws(Handle *h) {
    if (!h->fp) {
        FpWs *W = new FpWs();
        W->partials.alloc(3 * 1024 + 8);
        W->out.alloc(8);
        W->counter.alloc(4);
        W->key.alloc(2);
        W->counter.zero(h->stream);
        FPSB_CUDA(cudaMallocHost((void **)&W->h_out, 8 * sizeof(double)));
        h->fp = W;
    }
    return h->fp;
}
void fp_free(Handle *h) {
    if (!h->fp) return;
    if (h->fp->h_out) cudaFreeHost(h->fp->h_out);
    delete h->fp;
    h->fp = nullptr;
}

void fp_ys_gs(Handle *h, int64_t n, int64_t m, double sigma, const double *p1, const double *q1, const double *p2,
              const double *q2, double *gs, double *ys, double *v, double *w) {
    const int64_t N = n > m ? n : m;
    if (N == 0) return;
    fp_ys_gs_kernel<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>((int)n, (int)m, sigma, p1, q1, p2, q2, gs, ys, v, w);
    h->launches += 1;
    FPSB_CUDA(cudaGetLastError());
}
double fp_obj(Handle *h, int64_t n, int64_t m, double fx, double rho, double eta, const double *c, const double *ys,
              const double *x, const double *xk) {
    FpWs *W = fp_ws(h);
    const bool prox = eta > 0.0 && x && xk;
    const int64_t N = (prox && n > m) ? n : m;
    int grid = (int)std::min<int64_t>(1024, std::max<int64_t>(1, (N + 255) / 256));
    fp_obj_kernel<<<grid, 256, 0, h->stream>>>((int)m, (int)n, c, ys, x, prox ? xk : nullptr, W->partials.p, W->counter.p, W->out.p);
    h->launches += 1;
    FPSB_CUDA(cudaMemcpyAsync(W->h_out, W->out.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FPSB_CUDA(cudaStreamSynchronize(h->stream));
    double phi = fx - W->h_out[0] + rho / 2 * W->h_out[1];
    if (prox) phi += eta / 2 * W->h_out[2];
    return phi;
}
void fp_grad(Handle *h, int64_t n, double sigma, double rho, double eta, const double *gs, const double *Hsv, const double *v,
             const double *Sstw, const double *Jtc, const double *x, const double *xk, double *g) {
    if (n == 0) return;
    fp_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((int)n, sigma, rho, eta, gs, Hsv, v, Sstw, Jtc, x, xk, g);
    h->launches += 1;
    FPSB_CUDA(cudaGetLastError());
}
void fp_ptv(Handle *h, int64_t n, const double *v, const double *p1, double *Ptv) {
    if (n == 0) return;
    fp_ptv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((int)n, v, p1, Ptv);
    h->launches += 1;
}
void fp_hprod2(Handle *h, int64_t n, double sigma, double rho, double eta, double obj_weight, const double *p2, const double *HsPtv,
               const double *Ptv, const double *Hcv, const double *JtJv, const double *v, double *Hv) {
    if (n == 0) return;
    fp_hprod2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((int)n, sigma, rho, eta, obj_weight, p2, HsPtv, Ptv, Hcv, JtJv, v, Hv);
    h->launches += 1;
}
void fp_hprod1(Handle *h, int64_t n, double sigma, double rho, double eta, double obj_weight, const double *p2, const double *HsPtv,
               const double *Ptv, const double *JtinvJtJSsv, const double *SsinvJtJJv, const double *Hcv, const double *JtJv,
               const double *v, double *Hv) {
    if (n == 0) return;
    fp_hprod1_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((int)n, sigma, rho, eta, obj_weight, p2, HsPtv, Ptv, JtinvJtJSsv,
                                                                        SsinvJtJJv, Hcv, JtJv, v, Hv);
    h->launches += 1;
}
uint64_t fp_hash(Handle *h, int64_t n, const double *x) {
    FpWs *W = fp_ws(h);
    FPSB_CUDA(cudaMemsetAsync(W->key.p, 0, sizeof(unsigned long long), h->stream));
    if (n > 0) {
        int grid = (int)std::min<int64_t>(1024, (n + 255) / 256);
        fp_hash_kernel<<<grid, 256, 0, h->stream>>>((int)n, x, W->key.p);
        h->launches += 1;
    }
    unsigned long long *hk = reinterpret_cast<unsigned long long *>(W->h_out + 4);
    FPSB_CUDA(cudaMemcpyAsync(hk, W->key.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    FPSB_CUDA(cudaStreamSynchronize(h->stream));
    return (uint64_t)*hk;
}

}  // namespace fpsb
