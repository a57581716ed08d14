// Kernel (d): pseudo-log-likelihood stage, plus the optimiser / input-conversion kernels.
//   pll_count   histogram over (variable, code, y)   reference core/model.py:58-82
//   cpt         Laplace-smoothed table, float64       reference core/model.py:88
//   pll_reduce  float64 log-likelihood reduction      reference core/model.py:93-96
//   adam_step   Keras Adam (ResourceApplyAdam form)   reference run.py:60
//   y_to_f32    uint8 data matrix -> fp32 operand     replaces make_xs, reference run.py:46-50
#include <cuda_bf16.h>

#include "common.cuh"
#include "ops.cuh"

namespace {

constexpr int PLL_THREADS = 512;
constexpr int PLL_TILE = 512;    // rows staged per step

// One CTA owns a tile of VT variables and a contiguous range of at most 32768 rows.  Its histogram
// lives in shared memory as ONE 32-bit word per (variable, code): n0 in the low half, n1 in the high
// half, so a sample is a single shared-memory atomic (+1 or +65536) and 32 variables x 512 codes fit
// 64 KB (three CTAs per SM).  idx is read along its contiguous sample axis (128 bytes per warp
// request, four requests in flight per thread); y goes through a [PLL_TILE][VT] byte tile.
__global__ void __launch_bounds__(PLL_THREADS) pll_count_kernel(
    const int32_t* __restrict__ idx, long long idx_gs, const uint8_t* __restrict__ y, int ldy, int g0,
    unsigned long long* __restrict__ n1, unsigned long long* __restrict__ n0, int G, int B, int K, int VT,
    int rows_per_cta) {
    extern __shared__ __align__(16) unsigned int hist[];              // [VT][K] packed (n1 << 16 | n0)
    unsigned char* ytile = reinterpret_cast<unsigned char*>(hist + (size_t)VT * K);    // [PLL_TILE][VT]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int gv0 = blockIdx.y * VT;
    const int nv = min(VT, G - gv0);
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(B, r0 + rows_per_cta);
    for (int i = t; i < VT * K; i += PLL_THREADS) hist[i] = 0u;
    __syncthreads();
    for (int rb = r0; rb < r1; rb += PLL_TILE) {
        const int nr = min(PLL_TILE, r1 - rb);
        for (int i = t; i < nr * nv; i += PLL_THREADS) {
            const int row = i / nv, c = i - row * nv;
            ytile[row * VT + c] = y[(long long)(rb + row) * ldy + g0 + gv0 + c];
        }
        __syncthreads();
        for (int c = warp; c < nv; c += PLL_THREADS / 32) {           // one variable per warp at a time
            const int32_t* ic = idx + (long long)(gv0 + c) * idx_gs + rb;
            unsigned int* hc = hist + (size_t)c * K;
            int rr = lane;
            for (; rr + 96 < nr; rr += 128) {                          // four independent loads in flight
                const int k0 = ic[rr], k1 = ic[rr + 32], k2 = ic[rr + 64], k3 = ic[rr + 96];
                atomicAdd(&hc[k0], ytile[rr * VT + c] ? 65536u : 1u);
                atomicAdd(&hc[k1], ytile[(rr + 32) * VT + c] ? 65536u : 1u);
                atomicAdd(&hc[k2], ytile[(rr + 64) * VT + c] ? 65536u : 1u);
                atomicAdd(&hc[k3], ytile[(rr + 96) * VT + c] ? 65536u : 1u);
            }
            for (; rr < nr; rr += 32) atomicAdd(&hc[ic[rr]], ytile[rr * VT + c] ? 65536u : 1u);
        }
        __syncthreads();
    }
    for (int i = t; i < nv * K; i += PLL_THREADS) {
        const unsigned int h = hist[i];
        if (h) {
            const int c = i / K, k = i - c * K;
            const long long o = (long long)(gv0 + c) * K + k;
            if (h & 0xffffu) atomicAdd(&n0[o], (unsigned long long)(h & 0xffffu));
            if (h >> 16) atomicAdd(&n1[o], (unsigned long long)(h >> 16));
        }
    }
}

__global__ void cpt_kernel(const unsigned long long* __restrict__ n1, const unsigned long long* __restrict__ n0,
                           double* __restrict__ dist, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double a = (double)n1[i], b = (double)n0[i];
        dist[i] = (a + 0.8) / (a + b + 1.6);
    }
}

__global__ void __launch_bounds__(256) pll_reduce_kernel(const unsigned long long* __restrict__ n1,
                                                         const unsigned long long* __restrict__ n0,
                                                         const double* __restrict__ dist, long long n,
                                                         double* out) {
    __shared__ double red[8];
    double part = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double p = dist[i];
        part += (double)n1[i] * log(p + 1e-5) + (double)n0[i] * log(1.0 - p + 1e-5);
    }
    part = pg_warp_sum_d(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int i = 0; i < 8; ++i) a += red[i];
        atomicAdd(out, a);
    }
}

// wb (optional): bf16 mirror of the first nwb parameters (the operand copy of the bf16 tensor-core path), written
// with the updated values so that no separate conversion pass re-reads the weights
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   float alpha, float omb1, float omb2, float eps,
                                                   __nv_bfloat16* __restrict__ wb, long long nwb) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
#define PG_ADAM1(c)                                        \
        mm.c += (gg.c - mm.c) * omb1;                      \
        vv.c += (gg.c * gg.c - vv.c) * omb2;               \
        pp.c -= (mm.c * alpha) / (sqrtf(vv.c) + eps);
        PG_ADAM1(x) PG_ADAM1(y) PG_ADAM1(z) PG_ADAM1(w)
#undef PG_ADAM1
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (wb && (i << 2) + 4 <= nwb) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
            reinterpret_cast<uint2*>(wb)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float mm = m[i], vv = v[i];
        const float gg = g[i];
        mm += (gg - mm) * omb1;
        vv += (gg * gg - vv) * omb2;
        p[i] -= (mm * alpha) / (sqrtf(vv) + eps);
        m[i] = mm;
        v[i] = vv;
        if (wb && i < nwb) wb[i] = __float2bfloat16_rn(p[i]);
    }
}

__global__ void __launch_bounds__(256) flat_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 t = reinterpret_cast<const float4*>(src)[i];
        const __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
        reinterpret_cast<uint2*>(dst)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void y_to_f32_kernel(const uint8_t* __restrict__ y, int ldy, float* __restrict__ out, int ld, int B, int V) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * ld) return;
    const int b = (int)(i / ld), c = (int)(i - (long long)b * ld);
    out[i] = c < V ? (y[(long long)b * ldy + c] != 0 ? 1.0f : 0.0f) : 0.0f;
}

}  // namespace

extern "C" {

int pgmvae_pll_count(pgmvae_ctx* ctx, void* stream, const int32_t* idx, int64_t idx_gs, const uint8_t* y, int ldy,
                     int g0, unsigned long long* n1, unsigned long long* n0, int G, int B, int K) {
    PG_CHECK_ARG(ctx && idx && y && n1 && n0);
    PG_CHECK_ARG(K > 0 && G >= 0 && B >= 0);
    if (G == 0 || B == 0) return PGMVAE_OK;
    int VT = (int)((64 * 1024) / ((size_t)K * 4));
    if (VT > 32) VT = 32;
    if (VT > G) VT = G;
    if (VT < 1) {
        pgmvae_set_error("pll_count: K=%d too large for the shared-memory histogram", K);
        return PGMVAE_EINVAL;
    }
    const size_t smem = (size_t)VT * K * 4 + (size_t)PLL_TILE * VT;
    const int vtiles = (int)pg_cdiv(G, VT);
    // enough row splits to give every SM ~3 CTAs; at most 32768 rows per CTA (16-bit halves of the packed counters)
    int splits = (int)pg_cdiv(3 * ctx->sm_count, vtiles);
    const int max_splits = (int)pg_cdiv(B, 4 * PLL_TILE);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rows_per_cta = pg_round_up((int)pg_cdiv(B, splits), PLL_TILE);
    if (rows_per_cta > 32768) rows_per_cta = 32768;
    splits = (int)pg_cdiv(B, rows_per_cta);
    static size_t configured[16] = {};          // per device: the attribute is set per device
    if (smem > 48 * 1024 && smem > configured[ctx->device & 15]) {
        PG_CUDA(cudaFuncSetAttribute(pll_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15] = smem;
    }
    dim3 grid((unsigned)splits, (unsigned)vtiles);
    PG_KERNEL(ctx, pg_stream(ctx, stream), "pll_count", (double)G * B * 5.0 + 2.0 * G * K * 8.0, (double)G * B);
    pll_count_kernel<<<grid, PLL_THREADS, smem, pg_stream(ctx, stream)>>>(idx, idx_gs, y, ldy, g0, n1, n0, G, B, K, VT,
                                                                      rows_per_cta);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

int pgmvae_cpt(pgmvae_ctx* ctx, void* stream, const unsigned long long* n1, const unsigned long long* n0, double* dist,
               int64_t count) {
    PG_CHECK_ARG(ctx && n1 && n0 && dist && count >= 0);
    if (count == 0) return PGMVAE_OK;
    PG_KERNEL(ctx, pg_stream(ctx, stream), "cpt", 24.0 * count, 3.0 * count);
    cpt_kernel<<<(unsigned)pg_cdiv(count, 256), 256, 0, pg_stream(ctx, stream)>>>(n1, n0, dist, count);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

int pgmvae_pll_reduce(pgmvae_ctx* ctx, void* stream, const unsigned long long* n1, const unsigned long long* n0,
                      const double* dist, int64_t count, double* out_sum) {
    PG_CHECK_ARG(ctx && n1 && n0 && dist && out_sum && count >= 0);
    cudaStream_t st = pg_stream(ctx, stream);
    PG_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(double), st));
    if (count == 0) return PGMVAE_OK;
    int blocks = (int)std::min<int64_t>(pg_cdiv(count, 256 * 4), ctx->sm_count);
    if (blocks < 1) blocks = 1;
    PG_KERNEL(ctx, st, "pll_reduce", 24.0 * count, 6.0 * count);
    pll_reduce_kernel<<<blocks, 256, 0, st>>>(n1, n0, dist, count, out_sum);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

int pgmvae_adam_step(pgmvae_ctx* ctx, void* stream, float* p, const float* g, float* m, float* v, int64_t n,
                     float alpha, double b1, double b2, double eps) {
    return pg_adam_step_shadow(ctx, pg_stream(ctx, stream), p, g, m, v, n, alpha, b1, b2, eps, nullptr, 0);
}

}  // extern "C"

int pg_flat_to_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* src, __nv_bfloat16* dst, int64_t n) {
    if (n <= 0) return PGMVAE_OK;
    int blocks = (int)std::min<int64_t>(pg_cdiv(n, 256 * 4 * 4), (int64_t)ctx->sm_count * 8);
    if (blocks < 1) blocks = 1;
    PG_KERNEL(ctx, st, "weights_to_bf16", 6.0 * n, 0.0);
    flat_to_bf16_kernel<<<blocks, 256, 0, st>>>(src, dst, n);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

int pg_adam_step_shadow(pgmvae_ctx* ctx, cudaStream_t stream, float* p, const float* g, float* m, float* v, int64_t n,
                        float alpha, double b1, double b2, double eps, __nv_bfloat16* wb, int64_t nwb) {
    PG_CHECK_ARG(ctx && p && g && m && v && n >= 0);
    PG_CHECK_ARG(((uintptr_t)p & 15) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)m & 15) == 0 &&
                 ((uintptr_t)v & 15) == 0);
    if (n == 0) return PGMVAE_OK;
    int blocks = (int)std::min<int64_t>(pg_cdiv(n, 256 * 4 * 4), (int64_t)ctx->sm_count * 8);
    if (blocks < 1) blocks = 1;
    PG_KERNEL(ctx, stream, "adam", 28.0 * n + (wb ? 2.0 * nwb : 0.0), 10.0 * n);
    adam_kernel<<<blocks, 256, 0, stream>>>(p, g, m, v, n, alpha, (float)(1.0 - b1), (float)(1.0 - b2), (float)eps, wb,
                                            (long long)nwb);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

extern "C" {

int pgmvae_y_to_f32(pgmvae_ctx* ctx, void* stream, const uint8_t* y, int ldy, float* out, int ld, int B, int V) {
    PG_CHECK_ARG(ctx && y && out && ld >= V && ldy >= V && B >= 0);
    if (B == 0) return PGMVAE_OK;
    const long long n = (long long)B * ld;
    PG_KERNEL(ctx, pg_stream(ctx, stream), "y_to_f32", (double)B * V + 4.0 * n, 0.0);
    y_to_f32_kernel<<<(unsigned)pg_cdiv(n, 256), 256, 0, pg_stream(ctx, stream)>>>(y, ldy, out, ld, B, V);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

}  // extern "C"
