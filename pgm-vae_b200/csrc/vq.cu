// Kernels (b) and (c), CUDA-core fp32 flavour:
//   vq_assign    distances + argmin                 reference core/quantizer.py:44-47, :135-138
//   vq_quantize  gather + loss + straight-through   reference core/quantizer.py:49-53, :141-142,156
//   ema_stats    counts / per-code sums             reference core/quantizer.py:144-146
//   ema_apply    TF assign_moving_average x2 + Laplace + normalise  core/quantizer.py:144-152
// The codebook is held CODE-MAJOR ([K, D] per variable) so that a code row is one
// contiguous vector for the gather, the scatter and the tensor-core B operand alike.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "ops.cuh"

namespace {

constexpr int VQ_ROWS = 128;   // rows (samples) per CTA, one per thread
constexpr int VQ_KC = 64;      // codes staged per chunk

// dynamic smem: zs[D][VQ_ROWS] | es[D][VQ_KC] | ee[VQ_KC]
__global__ void __launch_bounds__(VQ_ROWS) vq_assign_kernel(
    const float* __restrict__ z, long long z_gs, int ldz, const float* __restrict__ e, long long e_gs, int lde,
    int32_t* __restrict__ idx, long long idx_gs, float* __restrict__ best_out, float* __restrict__ gap_out,
    int B, int D, int K, const int* __restrict__ gidx) {
    extern __shared__ __align__(16) float smem[];
    float* zs = smem;
    float* es = zs + (size_t)D * VQ_ROWS;
    float* ee = es + (size_t)D * VQ_KC;
    const int g = blockIdx.y, t = threadIdx.x;
    const int b0 = blockIdx.x * VQ_ROWS;
    const float* zg = z + (long long)g * z_gs;
    const float* eg = e + (long long)(gidx ? __ldg(gidx + g) : g) * e_gs;        // codebook of network gidx[g] (fts path)

    // stage the z tile transposed: zs[d][row]
    for (int i = t; i < VQ_ROWS * D; i += VQ_ROWS) {
        const int row = i / D, d = i - row * D;
        const int b = b0 + row;
        zs[d * VQ_ROWS + row] = b < B ? zg[(long long)b * ldz + d] : 0.f;
    }
    __syncthreads();
    float zz = 0.f;
    for (int d = 0; d < D; ++d) {
        const float v = zs[d * VQ_ROWS + t];
        zz = fmaf(v, v, zz);
    }
    float best = INFINITY, second = INFINITY;
    int bi = 0;
    for (int k0 = 0; k0 < K; k0 += VQ_KC) {
        __syncthreads();
        for (int i = t; i < VQ_KC * D; i += VQ_ROWS) {
            const int kk = i / D, d = i - kk * D;
            es[d * VQ_KC + kk] = (k0 + kk) < K ? eg[(long long)(k0 + kk) * lde + d] : 0.f;
        }
        __syncthreads();
        if (t < VQ_KC) {
            float s = 0.f;
            for (int d = 0; d < D; ++d) {
                const float v = es[d * VQ_KC + t];
                s = fmaf(v, v, s);
            }
            ee[t] = s;
        }
        __syncthreads();
        const int kn = min(VQ_KC, K - k0);
        for (int kk = 0; kk < kn; kk += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            for (int d = 0; d < D; ++d) {
                const float zv = zs[d * VQ_ROWS + t];
                const float4 e0 = *reinterpret_cast<const float4*>(&es[d * VQ_KC + kk]);
                const float4 e1 = *reinterpret_cast<const float4*>(&es[d * VQ_KC + kk + 4]);
                acc[0] = fmaf(zv, e0.x, acc[0]); acc[1] = fmaf(zv, e0.y, acc[1]);
                acc[2] = fmaf(zv, e0.z, acc[2]); acc[3] = fmaf(zv, e0.w, acc[3]);
                acc[4] = fmaf(zv, e1.x, acc[4]); acc[5] = fmaf(zv, e1.y, acc[5]);
                acc[6] = fmaf(zv, e1.z, acc[6]); acc[7] = fmaf(zv, e1.w, acc[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (kk + j < kn) {
                    // reference association: (|z|^2 - 2 z.e) + |e|^2
                    const float dist = (zz - 2.0f * acc[j]) + ee[kk + j];
                    if (dist < best) { second = best; best = dist; bi = k0 + kk + j; }
                    else if (dist < second) second = dist;
                }
            }
        }
    }
    const int b = b0 + t;
    if (b < B) {
        idx[(long long)g * idx_gs + b] = bi;
        if (best_out) best_out[(long long)g * idx_gs + b] = best;
        if (gap_out) gap_out[(long long)g * idx_gs + b] = second - best;
    }
}

__global__ void __launch_bounds__(256) vq_quantize_kernel(
    const float* __restrict__ z, long long z_gs, int ldz, const float* __restrict__ e, long long e_gs, int lde,
    const int32_t* __restrict__ idx, long long idx_gs, float* __restrict__ q, float* __restrict__ st,
    long long q_gs, int ldq, double* loss_acc, int B, int D) {
    __shared__ double red[8];
    const int g = blockIdx.y;
    const long long n = (long long)B * D;
    float part = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / D), d = (int)(i - (long long)b * D);
        const int k = idx[(long long)g * idx_gs + b];
        const float zv = z[(long long)g * z_gs + (long long)b * ldz + d];
        const float qv = __ldg(e + (long long)g * e_gs + (long long)k * lde + d);
        const float diff = qv - zv;
        part = fmaf(diff, diff, part);
        const long long o = (long long)g * q_gs + (long long)b * ldq + d;
        if (q) q[o] = qv;
        if (st) st[o] = zv + diff;      // inputs + stop_gradient(quantized - inputs)
    }
    double s = pg_warp_sum_d((double)part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0 && loss_acc) {
        double a = 0;
        for (int i = 0; i < 8; ++i) a += red[i];
        atomicAdd(loss_acc, a);
    }
}

// Segmented scatter-add of rows into per-code accumulators.
//   MODE 0 (EMA statistics): acc[k,:] += z[b,:], cnt[k] += 1
//   MODE 1 (codebook gradient): acc[k,:] += scale * (q[b,:] - z[b,:])
// SMEM=true: CTA-private accumulators in shared memory (K*D floats), flushed once per CTA
// with one global atomic per touched element; SMEM=false: direct global reductions
// (codebooks too large for shared memory).  One warp per row, lanes across D.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// Codebooks too large for shared memory (cfg4: 8192 x 64): every row is read once with 128-bit loads
// (LPR lanes per row, 32 / LPR rows per warp request, eight requests in flight) and added to its code
// with 128-bit vector reductions resolved in L2 -- one RED per 16 bytes instead of one per float.
constexpr int SCATTER_UNROLL = 8;       // 128-bit row loads in flight per thread

template <int MODE>
__global__ void __launch_bounds__(256) scatter_rows_vec_kernel(
    const float* __restrict__ z, const float* __restrict__ q, long long z_gs, int ldz,
    const int32_t* __restrict__ idx, long long idx_gs, float* __restrict__ cnt, long long c_gs,
    float* __restrict__ acc, long long a_gs, int lda, float scale, int B, int D4, int LPR) {
    const int g = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, l = lane - sub * LPR;          // row within the warp request, float4 within the row
    const int rpw = 32 / LPR;                                  // rows per warp request
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float* zg = z + (long long)g * z_gs;
    const float* qg = q ? q + (long long)g * z_gs : nullptr;
    const int32_t* ig = idx + (long long)g * idx_gs;
    float* ag = acc + (long long)g * a_gs;
    float* cg = cnt ? cnt + (long long)g * c_gs : nullptr;
    for (long long b0 = warp_id * rpw * SCATTER_UNROLL; b0 < B; b0 += nwarps * rpw * SCATTER_UNROLL) {
        float4 v[SCATTER_UNROLL];
        int k[SCATTER_UNROLL];
#pragma unroll
        for (int u = 0; u < SCATTER_UNROLL; ++u) {
            const long long b = b0 + u * rpw + sub;
            k[u] = -1;
            if (b < B && l < D4) {
                k[u] = ig[b];
                v[u] = *reinterpret_cast<const float4*>(zg + b * ldz + 4 * l);
                if (MODE == 1) {
                    const float4 qq = *reinterpret_cast<const float4*>(qg + b * ldz + 4 * l);
                    v[u] = make_float4(scale * (qq.x - v[u].x), scale * (qq.y - v[u].y), scale * (qq.z - v[u].z),
                                       scale * (qq.w - v[u].w));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < SCATTER_UNROLL; ++u) {
            if (k[u] >= 0) {
                red_add_v4(ag + (long long)k[u] * lda + 4 * l, v[u]);
                if (MODE == 0 && l == 0) atomicAdd(&cg[k[u]], 1.0f);
            }
        }
    }
}

// (A counting sort of the rows by code in front of the scatter, so that runs of one code are summed in registers, was
// built and measured in round 2 at the cfg4 shape: 2.67 TB/s against the 3.50 TB/s of the direct vector reductions
// above -- the three extra passes over idx and the random 256-byte row reads cost more than the L2 reductions they
// save -- and was removed again.)

template <int MODE, bool SMEM>
__global__ void __launch_bounds__(256) scatter_rows_kernel(
    const float* __restrict__ z, const float* __restrict__ q, long long z_gs, int ldz,
    const int32_t* __restrict__ idx, long long idx_gs, float* __restrict__ cnt, long long c_gs,
    float* __restrict__ acc, long long a_gs, int lda, float scale, int B, int D, int K, int rows_per_cta) {
    extern __shared__ __align__(16) float sm[];
    float* sacc = sm;                       // [K][D]
    float* scnt = sm + (size_t)K * D;       // [K]
    const int g = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(B, r0 + rows_per_cta);
    const float* zg = z + (long long)g * z_gs;
    const float* qg = q ? q + (long long)g * z_gs : nullptr;
    const int32_t* ig = idx + (long long)g * idx_gs;
    float* ag = acc + (long long)g * a_gs;
    float* cg = cnt ? cnt + (long long)g * c_gs : nullptr;
    if (SMEM) {
        for (int i = threadIdx.x; i < K * D + K; i += blockDim.x) sm[i] = 0.f;
        __syncthreads();
    }
    for (int b = r0 + warp; b < r1; b += nwarp) {
        const int k = ig[b];
        for (int d = lane; d < D; d += 32) {
            float v = zg[(long long)b * ldz + d];
            if (MODE == 1) v = scale * (qg[(long long)b * ldz + d] - v);
            if (SMEM) atomicAdd(&sacc[k * D + d], v);
            else atomicAdd(&ag[(long long)k * lda + d], v);
        }
        if (MODE == 0 && lane == 0) {
            if (SMEM) atomicAdd(&scnt[k], 1.0f);
            else atomicAdd(&cg[k], 1.0f);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
            const float v = sacc[i];
            if (v != 0.f) {
                const int k = i / D, d = i - k * D;
                atomicAdd(&ag[(long long)k * lda + d], v);
            }
        }
        if (MODE == 0)
            for (int k = threadIdx.x; k < K; k += blockDim.x)
                if (scnt[k] != 0.f) atomicAdd(&cg[k], scnt[k]);
    }
}

// EMA update, step 1 (one CTA per variable): the per-code counts
__global__ void __launch_bounds__(256) ema_counts_kernel(
    const float* __restrict__ counts, float* __restrict__ biased_c, float* __restrict__ ema_c, int K,
    float one_minus, float bias_factor, int zero_debias) {
    const long long co = (long long)blockIdx.x * K;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float u;
        if (zero_debias) {
            float bc = biased_c[co + k];
            bc = bc - (bc - counts[co + k]) * one_minus;
            biased_c[co + k] = bc;
            u = bc / bias_factor;
        } else {
            const float c = ema_c[co + k];
            u = c - (c - counts[co + k]) * one_minus;
        }
        ema_c[co + k] = u;
    }
}

// EMA update, step 2 (grid: slices of K*D x variables): n = sum_k ema_c (re-derived per CTA in the same
// order, so every CTA of a variable sees the identical value), per-code sums, Laplace smoothing and the
// normalised codebook write-back.  Elements are walked with float4 where the layout allows.
__global__ void __launch_bounds__(256) ema_apply_kernel(
    const float* __restrict__ dw, float* __restrict__ biased_w, const float* __restrict__ ema_c,
    float* __restrict__ ema_w, float* __restrict__ e, int K, int D, int ld, float one_minus, float epsilon,
    float bias_factor, int zero_debias, int per_cta) {
    __shared__ float red[8];
    __shared__ float n_sh;
    const int g = blockIdx.y;
    const long long co = (long long)g * K, wo = (long long)g * K * ld;
    float part = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) part += ema_c[co + k];
    part = pg_warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float n = 0.f;
        for (int i = 0; i < 8; ++i) n += red[i];
        n_sh = n;
    }
    __syncthreads();
    const float n = n_sh;
    const float denom = n + (float)K * epsilon;
    const int i0 = blockIdx.x * per_cta, i1 = min(K * D, i0 + per_cta);
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const int k = i / D, d = i - k * D;
        const long long o = wo + (long long)k * ld + d;
        float u;
        if (zero_debias) {
            float bw = biased_w[o];
            bw = bw - (bw - dw[o]) * one_minus;
            biased_w[o] = bw;
            u = bw / bias_factor;
        } else {
            const float w = ema_w[o];
            u = w - (w - dw[o]) * one_minus;
        }
        ema_w[o] = u;
        const float size = (ema_c[co + k] + epsilon) / denom * n;   // core/quantizer.py:149-150
        e[o] = u / size;                                             // :151-152
    }
}

}  // namespace

int pg_vq_assign_fp32(pgmvae_ctx* ctx, cudaStream_t st, const float* z, int64_t z_gs, int ldz, const float* e,
                      int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt, int G,
                      int B, int D, int K, const int* gidx) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    const size_t smem = ((size_t)D * VQ_ROWS + (size_t)D * VQ_KC + VQ_KC) * sizeof(float);
    if (smem > ctx->smem_optin) {
        pgmvae_set_error("vq_assign: embedding dim %d too large for the fp32 kernel", D);
        return PGMVAE_EINVAL;
    }
    static size_t configured[16] = {};          // per device: the attribute is set per device
    if (smem > 48 * 1024 && smem > configured[ctx->device & 15]) {
        PG_CUDA(cudaFuncSetAttribute(vq_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15] = smem;
    }
    dim3 grid((unsigned)pg_cdiv(B, VQ_ROWS), (unsigned)G);
    PG_KERNEL(ctx, st, "vq_assign_fp32", 4.0 * ((double)G * B * D + (double)G * K * D + (double)G * B),
              2.0 * G * B * (double)D * K);
    vq_assign_kernel<<<grid, VQ_ROWS, smem, st>>>(z, z_gs, ldz, e, e_gs, lde, idx, idx_gs, best_opt, gap_opt, B, D, K, gidx);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

template <int MODE>
static int scatter_launch(pgmvae_ctx* ctx, cudaStream_t st, const float* z, const float* q, int64_t z_gs, int ldz,
                          const int32_t* idx, int64_t idx_gs, float* cnt, int64_t c_gs, float* acc, int64_t a_gs,
                          int lda, float scale, int G, int B, int D, int K) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    const size_t smem = ((size_t)K * D + K) * sizeof(float);
    // CTA-private accumulators only while several CTAs fit an SM; beyond that the shared-memory atomics of one
    // resident CTA are slower (196 GB/s at K = 512, D = 64) than vector reductions in L2
    const bool vec_ok = D % 4 == 0 && D <= 128 && ldz % 4 == 0 && lda % 4 == 0 && z_gs % 4 == 0 && a_gs % 4 == 0 &&
                        !((uintptr_t)z & 15) && !((uintptr_t)acc & 15) && (!q || !((uintptr_t)q & 15));
    const bool use_smem = smem <= (vec_ok ? 32 : 160) * 1024;
    int rows = use_smem ? max(1024, 4 * K) : 2048;
    // keep at least ~2 CTAs per SM in flight when the problem allows it
    while (rows > 256 && pg_cdiv(B, rows) * G < 2 * ctx->sm_count) rows >>= 1;
    dim3 grid((unsigned)pg_cdiv(B, rows), (unsigned)G);
    // algorithmic bytes: read z (+q) and idx once, write the per-code sums and counts
    PG_KERNEL(ctx, st, MODE == 0 ? "ema_stats_scatter" : "vq_codebook_grad_scatter",
              (double)G * B * (4.0 * D * (MODE == 1 ? 2 : 1) + 4.0) + 4.0 * G * K * (D + 1.0), (double)G * B * D);
    if (use_smem) {
        static size_t configured[16][2] = {};          // per device
        if (smem > 48 * 1024 && smem > configured[ctx->device & 15][MODE]) {
            PG_CUDA(cudaFuncSetAttribute(scatter_rows_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
            configured[ctx->device & 15][MODE] = smem;
        }
        scatter_rows_kernel<MODE, true><<<grid, 256, smem, st>>>(z, q, z_gs, ldz, idx, idx_gs, cnt, c_gs, acc, a_gs,
                                                                 lda, scale, B, D, K, rows);
    } else if (vec_ok) {
        int lpr = 1;
        while (lpr < D / 4) lpr <<= 1;                       // lanes per row: power of two >= D / 4
        int per_sm = 8;
        if (const char* ev = getenv("PGMVAE_SCATTER_CTAS")) per_sm = atoi(ev) > 0 ? atoi(ev) : 8;
        dim3 vgrid((unsigned)(ctx->sm_count * per_sm), (unsigned)G);
        scatter_rows_vec_kernel<MODE><<<vgrid, 256, 0, st>>>(z, q, z_gs, ldz, idx, idx_gs, cnt, c_gs, acc, a_gs, lda,
                                                              scale, B, D / 4, lpr);
    } else {
        scatter_rows_kernel<MODE, false><<<grid, 256, 0, st>>>(z, q, z_gs, ldz, idx, idx_gs, cnt, c_gs, acc, a_gs,
                                                               lda, scale, B, D, K, rows);
    }
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

extern "C" {

int pgmvae_vq_quantize(pgmvae_ctx* ctx, void* stream, const float* z, int64_t z_gs, int ldz, const float* e,
                       int64_t e_gs, int lde, const int32_t* idx, int64_t idx_gs, float* q, float* st, int64_t q_gs,
                       int ldq, double* loss_acc, int G, int B, int D, int K) {
    PG_CHECK_ARG(ctx && z && e && idx);
    PG_CHECK_ARG(G >= 0 && B >= 0 && D > 0 && K > 0);
    if (G == 0 || B == 0) return PGMVAE_OK;
    const long long n = (long long)B * D;
    int bx = (int)std::min<long long>(pg_cdiv(n, 256 * 4), 1024);
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)G);
    PG_KERNEL(ctx, pg_stream(ctx, stream), "vq_quantize", 4.0 * G * B * (3.0 * D + 1.0), 3.0 * G * B * D);
    vq_quantize_kernel<<<grid, 256, 0, pg_stream(ctx, stream)>>>(z, z_gs, ldz, e, e_gs, lde, idx, idx_gs, q, st, q_gs,
                                                                ldq, loss_acc, B, D);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

int pgmvae_vq_codebook_grad(pgmvae_ctx* ctx, void* stream, const float* z, const float* q, int64_t zq_gs, int ldzq,
                            const int32_t* idx, int64_t idx_gs, float* de, int64_t de_gs, int ldde, float scale,
                            int G, int B, int D, int K) {
    PG_CHECK_ARG(ctx && z && q && idx && de);
    return scatter_launch<1>(ctx, pg_stream(ctx, stream), z, q, zq_gs, ldzq, idx, idx_gs, nullptr, 0, de, de_gs, ldde,
                             scale, G, B, D, K);
}

int pgmvae_ema_stats(pgmvae_ctx* ctx, void* stream, const float* z, int64_t z_gs, int ldz, const int32_t* idx,
                     int64_t idx_gs, float* counts, int64_t c_gs, float* dw, int64_t dw_gs, int lddw, int G, int B,
                     int D, int K) {
    PG_CHECK_ARG(ctx && z && idx && counts && dw);
    PG_CHECK_ARG(D > 0 && K > 0);
    return scatter_launch<0>(ctx, pg_stream(ctx, stream), z, nullptr, z_gs, ldz, idx, idx_gs, counts, c_gs, dw, dw_gs,
                             lddw, 1.0f, G, B, D, K);
}

int pgmvae_ema_apply(pgmvae_ctx* ctx, void* stream, const float* counts, const float* dw, float* biased_c,
                     float* biased_w, float* ema_c, float* ema_w, float* e, int G, int K, int D, int ld, double decay,
                     double epsilon, int step, int zero_debias) {
    PG_CHECK_ARG(ctx && counts && dw && ema_c && ema_w && e);
    PG_CHECK_ARG(!zero_debias || (biased_c && biased_w && step >= 1));
    if (G <= 0) return PGMVAE_OK;
    // TF: decay tensor = float32(1.0 - decay); bias_factor = 1 - pow(1.0 - decay_tensor, local_step) in fp32
    const float one_minus = (float)(1.0 - decay);
    const float bias_factor = 1.0f - powf(1.0f - one_minus, (float)step);
    PG_KERNEL(ctx, pg_stream(ctx, stream), "ema_apply", 4.0 * G * K * D * 5.0 + 4.0 * 4.0 * G * K, 6.0 * G * K * D);
    ema_counts_kernel<<<G, 256, 0, pg_stream(ctx, stream)>>>(counts, biased_c, ema_c, K, one_minus, bias_factor,
                                                           zero_debias);
    ctx->launches++;
    // slices of >= 4096 elements; enough CTAs to fill the machine when K * D is large (cfg4: 8192 x 64)
    int slices = (int)pg_cdiv((int64_t)K * D, 4096);
    const int want = (int)pg_cdiv(4 * ctx->sm_count, G);
    if (slices > want) slices = want;
    if (slices < 1) slices = 1;
    const int per_cta = (int)pg_cdiv((int64_t)K * D, slices);
    dim3 grid((unsigned)pg_cdiv((int64_t)K * D, per_cta), (unsigned)G);
    ema_apply_kernel<<<grid, 256, 0, pg_stream(ctx, stream)>>>(dw, biased_w, ema_c, ema_w, e, K, D, ld, one_minus,
                                                             (float)epsilon, bias_factor, zero_debias, per_cta);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

}  // extern "C"
