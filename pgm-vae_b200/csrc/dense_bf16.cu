// Kernel (a), bf16 tensor-core flavour (PGMVAE_PREC_BF16): the packed per-variable dense layer
// (reference core/dense.py:99-111) and the gradient GEMMs of its autodiff (run.py:62) as ONE
// persistent, warp-specialised grouped GEMM on tcgen05 kind::f16 (bf16 operands, fp32
// accumulation in TMEM).  This is the path for networks too wide for the TMEM-resident chain
// kernels (cfg3: 1556 variables, 1555 -> 400 -> ... -> 400 -> 1555), where the step is bound by
// the tensor pipe.
//
//   forward : C[B,out] = X[B,in]   (K-major A)  x  Wt[out,in]  (K-major B; bf16 shadow, transposed)
//   dgrad   : C[B,in]  = dY[B,out] (K-major A)  x  W[in,out]   (K-major B; bf16 shadow as stored)
//   wgrad   : C = X^T dY over the batch, both operands MN-major (the row-major bf16 activations as
//             they lie in HBM), in whichever orientation pads less:
//               direct      C[in,out]  thread = weight row, 128-bit row stores
//               transposed  C[out,in]  lanes = consecutive out columns of dW[in][out]: coalesced
//             K = the whole batch per CTA: no split-K, no atomics, dW is written exactly once.
//
// CTA (one per SM, persistent over tiles; 384 threads):
//   warp 0      TMA producer: A [128 x 64] + B [BN x 64] bf16 k-blocks (128-byte swizzle) into a ring
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BN <= 256, K = 16
//   warp 2      owns the TMEM allocation (all 512 columns: two accumulator buffers of 256)
//   warps 4-11  two epilogue warpgroups; warpgroup w drains accumulator buffer w, so the epilogue of
//               tile i runs under the main loop of tile i+1 (tmem_full / tmem_empty barriers)
// Tiles are numbered N fastest, then M, then variable: the CTAs in flight work on the same two or
// three variables, so weights and activations are shared through L2.
//
// Epilogues (thread = accumulator row, 32 columns per tcgen05.ld, the next chunk in flight while
// the current one is processed):
//   FWD          bias + selu / sigmoid / none -> bf16 row (next layer's operand) and/or fp32 row
//   SIGMOID_MSE  fd9: bias + sigmoid + squared / absolute error sums + d(loss)/d(pre-activation) in
//                bf16 (leave-one-out column masked), core/model.py:53 + run.py:61
//   DGRAD        (+ commitment gradient at the VQ boundary) x act'(activation below) -> bf16 row
//   WGRAD_D / WGRAD_T  fp32 weight gradient
#include <cuda_bf16.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "ops.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;                     // rows per MMA (TMEM lanes)
constexpr int BK = 64;                      // bf16 elements per k-block = one 128-byte swizzle row
constexpr int A_BYTES = TM * 128;           // 16 KB per stage
constexpr int PANEL_BYTES = BK * 128;       // MN-major operand: [64 k][64 mn] panel
constexpr int MAX_STAGES = 8;
constexpr int THREADS = 384;
constexpr int ACC_COLS = 256;               // TMEM columns per accumulator buffer

enum { EPI_FWD = 0, EPI_SIGMOID_MSE = 1, EPI_DGRAD = 2, EPI_WGRAD_D = 3, EPI_WGRAD_T = 4 };

struct Bf16P {
    int G, M, N, K, BN, tiles_m, tiles_n, kblocks, stages, total_tiles;
    int a_mn, b_mn, a_shared, b_shared, b_panels;
    unsigned a_bytes, b_bytes;
    int vec;                                              // rows allow 16-byte vector access
    __nv_bfloat16* cb; long long cb_gs; int ldcb;         // bf16 output rows (may be null)
    float* cf; long long cf_gs; int ldcf;                 // fp32 output rows (may be null)
    const float* bias; long long bias_gs; int act;
    const __nv_bfloat16* yb; int ldyb; double* acc; float gscale; int g0;      // SIGMOID_MSE
    const __nv_bfloat16* hb; long long hb_gs; int ldhb;   // DGRAD: activation below (bf16) ...
    const float* hf; long long hf_gs; int ldhf;           // ... or fp32
    const float* z; const float* q; long long zq_gs; int ldzq; float cscale;
    float* dw; long long dw_gs; int lddw; int zero_row_base, accum;           // WGRAD (accum: dW += instead of =)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float LOG2E = 1.4426950408889634f;
// selu with a bare MUFU.EX2; the second term is exactly 0 for x >= 0
__device__ __forceinline__ float selu_fast(float x) {
    return fmaf(PG_SELU_SCALE, fmaxf(x, 0.f), fmaf(PG_SELU_SCALE_ALPHA, ex2_approx(fminf(x, 0.f) * LOG2E), -PG_SELU_SCALE_ALPHA));
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-x * LOG2E)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// 32 consecutive values of a thread's row (nv of them valid); 16-byte accesses where whole groups are valid
__device__ __forceinline__ void store_bf16_row(__nv_bfloat16* dst, const float (&v)[32], int nv, int vec) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        if (vec && j + 8 <= nv) {
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                            pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (j + i < nv) dst[j + i] = __float2bfloat16_rn(v[j + i]);
        }
    }
}
__device__ __forceinline__ void store_f32_row(float* dst, const float (&v)[32], int nv, int vec) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        if (vec && j + 4 <= nv) {
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (j + i < nv) dst[j + i] = v[j + i];
        }
    }
}
__device__ __forceinline__ void load_bf16_row(const __nv_bfloat16* src, float (&v)[32], int nv, int vec) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        if (vec && j + 8 <= nv) {
            const uint4 t = *reinterpret_cast<const uint4*>(src + j);
            v[j] = bf16_lo(t.x); v[j + 1] = bf16_hi(t.x); v[j + 2] = bf16_lo(t.y); v[j + 3] = bf16_hi(t.y);
            v[j + 4] = bf16_lo(t.z); v[j + 5] = bf16_hi(t.z); v[j + 6] = bf16_lo(t.w); v[j + 7] = bf16_hi(t.w);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j + i] = j + i < nv ? __bfloat162float(src[j + i]) : 0.f;
        }
    }
}
__device__ __forceinline__ void load_f32_row(const float* src, float (&v)[32], int nv, int vec) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        if (vec && j + 4 <= nv) {
            const float4 t = *reinterpret_cast<const float4*>(src + j);
            v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[j + i] = j + i < nv ? src[j + i] : 0.f;
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ Bf16P p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, shared space
    uint8_t* sA = smem;                                           // [stages][16 KB]
    uint8_t* sB = sA + (size_t)p.stages * A_BYTES;                // [stages][b_bytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * p.b_bytes);
    uint64_t* full = bars;                        // [MAX_STAGES]   TMA -> MMA
    uint64_t* empty = bars + MAX_STAGES;          // [MAX_STAGES]   MMA -> TMA
    uint64_t* tmem_full = bars + 2 * MAX_STAGES;  // [2]            MMA -> epilogue warpgroup
    uint64_t* tmem_empty = tmem_full + 2;         // [2]            epilogue warpgroup -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&mapA);
        tc::tma_prefetch_desc(&mapB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tmem_full[b], 1);
            tc::mbar_init(&tmem_empty[b], 4);      // one arrival per warp of the warpgroup
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, 512u);
        tc::tmem_relinquish();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_g = p.tiles_m * p.tiles_n;

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        const uint32_t stage_tx = p.a_bytes + p.b_bytes;
        uint32_t s = 0, ph = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int g = tile / tiles_per_g, r = tile - g * tiles_per_g;
            const int mt = r / p.tiles_n, nt = r - mt * p.tiles_n;
            const int m0 = mt * TM, n0 = nt * p.BN;
            const int ga = p.a_shared ? 0 : g, gb = p.b_shared ? 0 : g;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                tc::mbar_wait(&empty[s], ph ^ 1);
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&full[s], stage_tx);
                    uint8_t* a_dst = sA + (size_t)s * A_BYTES;
                    uint8_t* b_dst = sB + (size_t)s * p.b_bytes;
                    if (!p.a_mn) {
                        tc::tma_load_3d(a_dst, &mapA, &full[s], kb * BK, m0, ga);                   // [128 m][64 k]
                    } else {
                        tc::tma_load_3d(a_dst, &mapA, &full[s], m0, kb * BK, ga);                   // 2 panels [64 k][64 m]
                        tc::tma_load_3d(a_dst + PANEL_BYTES, &mapA, &full[s], m0 + 64, kb * BK, ga);
                    }
                    if (!p.b_mn) {
                        tc::tma_load_3d(b_dst, &mapB, &full[s], kb * BK, n0, gb);                   // [BN n][64 k]
                    } else {
                        for (int pn = 0; pn < p.b_panels; ++pn)                                      // panels [64 k][64 n]
                            tc::tma_load_3d(b_dst + (size_t)pn * PANEL_BYTES, &mapB, &full[s], n0 + pn * 64, kb * BK, gb);
                    }
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        const uint32_t idesc = tc::make_idesc(1, TM, p.BN, p.a_mn, p.b_mn);
        // K-major : rows of 128 B along k, 8-row atoms 1024 B apart (SBO); one MMA (k = 16) advances 32 B
        // MN-major: panels [64 k][128 B along m/n], 8-k atoms 1024 B apart (SBO), panels PANEL_BYTES apart
        //           (LBO); one MMA (k = 16) advances two atoms = 2048 B
        const uint64_t dA0 = p.a_mn ? tc::make_smem_desc(tc::smem_u32(sA), PANEL_BYTES, 1024, 2)
                                    : tc::make_smem_desc(tc::smem_u32(sA), 16, 1024, 2);
        const uint64_t dB0 = p.b_mn ? tc::make_smem_desc(tc::smem_u32(sB), PANEL_BYTES, 1024, 2)
                                    : tc::make_smem_desc(tc::smem_u32(sB), 16, 1024, 2);
        const uint32_t a_step = (p.a_mn ? 2048u : 32u) >> 4, b_step = (p.b_mn ? 2048u : 32u) >> 4;
        const uint32_t a_stage = (uint32_t)A_BYTES >> 4, b_stage = p.b_bytes >> 4;
        uint32_t s = 0, ph = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1, use = it >> 1;
            tc::mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);          // the epilogue has drained this buffer
            tc::fence_after_thread_sync();
            const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                tc::mbar_wait(&full[s], ph);
                tc::fence_after_thread_sync();
                if (tc::elect_one()) {
                    const uint64_t dA = dA0 + (uint64_t)(s * a_stage), dB = dB0 + (uint64_t)(s * b_stage);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        tc::mma_f16(d_tmem, dA + (uint64_t)(k4 * a_step), dB + (uint64_t)(k4 * b_step), idesc,
                                    (kb > 0 || k4 > 0) ? 1u : 0u);
                    tc::mma_commit(&empty[s]);
                    if (kb == p.kblocks - 1) tc::mma_commit(&tmem_full[buf]);
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: two warpgroups, thread = accumulator row =====================
        const int wg = (warp - 4) >> 2;                // drains accumulator buffer wg (tiles it with it & 1 == wg)
        const int qd = warp & 3;                       // TMEM lane quarter of this warp
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + wg * ACC_COLS;
        double dsq = 0.0, dab = 0.0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            if ((int)(it & 1) != wg) continue;
            const uint32_t use = it >> 1;
            const int g = tile / tiles_per_g, r = tile - g * tiles_per_g;
            const int mt = r / p.tiles_n, nt = r - mt * p.tiles_n;
            const int m0 = mt * TM, n0 = nt * p.BN;
            const int row = m0 + qd * 32 + lane;
            const bool rvalid = row < p.M;
            const int ncols = min(p.BN, p.N - n0);
            const int nch = (ncols + 31) >> 5;
            float sq = 0.f, ab = 0.f;

            auto process = [&](float (&v)[32], int c) {
                const int nb = n0 + c * 32;
                const int nv = min(32, ncols - c * 32);            // valid columns of this chunk (tile and tensor bounds)
                if (EPI == EPI_FWD) {
                    if (!rvalid) return;
                    if (p.bias) {
                        float bv[32];
                        load_f32_row(p.bias + (long long)g * p.bias_gs + nb, bv, nv, p.vec);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += bv[j];
                    }
                    if (p.act == PGMVAE_ACT_SELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = selu_fast(v[j]);
                    } else if (p.act == PGMVAE_ACT_SIGMOID) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = sigmoid_fast(v[j]);
                    }
                    if (p.cb) store_bf16_row(p.cb + (long long)g * p.cb_gs + (long long)row * p.ldcb + nb, v, nv, p.vec);
                    if (p.cf) store_f32_row(p.cf + (long long)g * p.cf_gs + (long long)row * p.ldcf + nb, v, nv, p.vec);
                } else if (EPI == EPI_SIGMOID_MSE) {
                    if (!rvalid) return;
                    float t[32];
                    if (p.bias) {
                        load_f32_row(p.bias + (long long)g * p.bias_gs + nb, t, nv, p.vec);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += t[j];
                    }
                    load_bf16_row(p.yb + (long long)row * p.ldyb + nb, t, nv, p.vec);       // targets (0 / 1)
                    const int self = p.g0 + g - nb;                                         // masked column of this net
                    float o[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        o[j] = sigmoid_fast(v[j]);
                        float d = o[j] - t[j];
                        if (j >= nv || j == self) d = 0.f;
                        sq = fmaf(d, d, sq);
                        ab += fabsf(d);
                        const float u = p.gscale * o[j];
                        v[j] = d * fmaf(-u, o[j], u);                                       // gscale * d * o * (1 - o)
                    }
                    if (p.cb) store_bf16_row(p.cb + (long long)g * p.cb_gs + (long long)row * p.ldcb + nb, v, nv, p.vec);
                    if (p.cf) store_f32_row(p.cf + (long long)g * p.cf_gs + (long long)row * p.ldcf + nb, o, nv, p.vec);
                } else if (EPI == EPI_DGRAD) {
                    if (!rvalid) return;
                    float t[32];
                    if (p.z) {
                        const long long zo = (long long)g * p.zq_gs + (long long)row * p.ldzq + nb;
                        float qv[32];
                        load_f32_row(p.z + zo, t, nv, p.vec);
                        load_f32_row(p.q + zo, qv, nv, p.vec);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaf(p.cscale, t[j] - qv[j], v[j]);
                    }
                    if (p.hb || p.hf) {
                        if (p.hb) load_bf16_row(p.hb + (long long)g * p.hb_gs + (long long)row * p.ldhb + nb, t, nv, p.vec);
                        else load_f32_row(p.hf + (long long)g * p.hf_gs + (long long)row * p.ldhf + nb, t, nv, p.vec);
                        if (p.act == PGMVAE_ACT_SELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= pg_dselu_from_out(t[j]);
                        } else if (p.act == PGMVAE_ACT_SIGMOID) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= t[j] * (1.0f - t[j]);
                        }
                    }
                    if (p.cb) store_bf16_row(p.cb + (long long)g * p.cb_gs + (long long)row * p.ldcb + nb, v, nv, p.vec);
                    if (p.cf) store_f32_row(p.cf + (long long)g * p.cf_gs + (long long)row * p.ldcf + nb, v, nv, p.vec);
                } else if (EPI == EPI_WGRAD_D) {
                    // C[m = in][n = out] -> dW[in][out]: the thread owns (a piece of) a weight row
                    if (!rvalid) return;
                    float* dst = p.dw + (long long)g * p.dw_gs + (long long)row * p.lddw + nb;
                    if (row == p.zero_row_base + g) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
                    if (p.accum) {
                        float t[32];
                        load_f32_row(dst, t, nv, p.vec);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += t[j];
                    }
                    store_f32_row(dst, v, nv, p.vec);
                } else {
                    // C[m = out][n = in] -> dW[in][out]: lanes = consecutive out columns (coalesced per n)
                    if (!rvalid) return;
                    float* dst = p.dw + (long long)g * p.dw_gs + (long long)nb * p.lddw + row;
                    const int zr = p.zero_row_base + g - nb;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < nv) {
                            const float val = j == zr ? 0.f : v[j];
                            float* d1 = dst + (long long)j * p.lddw;
                            *d1 = p.accum ? *d1 + val : val;
                        }
                }
            };

            tc::mbar_wait(&tmem_full[wg], use & 1);
            tc::fence_after_thread_sync();
            float va[32], vb[32];
            tc::tmem_ld_32x32(lane_addr, va);
            for (int c = 0; c < nch; c += 2) {
                tc::tmem_ld_wait(va);
                if (c + 1 < nch) {
                    tc::tmem_ld_32x32(lane_addr + (c + 1) * 32, vb);
                } else {                                   // the whole accumulator sits in registers: hand it back
                    tc::fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&tmem_empty[wg]);
                }
                process(va, c);
                if (c + 1 < nch) {
                    tc::tmem_ld_wait(vb);
                    if (c + 2 < nch) {
                        tc::tmem_ld_32x32(lane_addr + (c + 2) * 32, va);
                    } else {
                        tc::fence_before_thread_sync();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&tmem_empty[wg]);
                    }
                    process(vb, c + 1);
                }
            }
            if (EPI == EPI_SIGMOID_MSE) { dsq += (double)sq; dab += (double)ab; }
        }
        if (EPI == EPI_SIGMOID_MSE) {
            dsq = pg_warp_sum_d(dsq);
            dab = pg_warp_sum_d(dab);
            if (lane == 0 && p.acc) {
                atomicAdd(p.acc, dsq);
                atomicAdd(p.acc + 1, dab);
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 2) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc(tmem_base, 512u);
    }
}

// ---- operand preparation ---------------------------------------------------------------------------------------
// fp32 [G][rows][ld] -> bf16 [G][rows][ldo] (columns >= cols are written as zero up to ldo)
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, long long s_gs, int lds, __nv_bfloat16* __restrict__ dst,
                                   long long d_gs, int ldd, int rows, int cols) {
    const int g = blockIdx.y;
    const long long n = (long long)rows * ldd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / ldd), c = (int)(i - (long long)r * ldd);
        const float v = c < cols ? src[(long long)g * s_gs + (long long)r * lds + c] : 0.f;
        dst[(long long)g * d_gs + i] = __float2bfloat16_rn(v);
    }
}

// W fp32 [G][rows][lds] -> Wt bf16 [G][cols][ldt] (transposed; row pad of Wt written as zero) and, optionally, the
// straight bf16 copy Wc [G][rows][ldc].  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) shadow_kernel(const float* __restrict__ src, long long s_gs, int lds, int rows, int cols,
                                                     __nv_bfloat16* __restrict__ wt, long long t_gs, int ldt,
                                                     __nv_bfloat16* __restrict__ wc, long long c_gs, int ldc) {
    __shared__ float tile[32][33];
    const int g = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        const float v = (r < rows && c < cols) ? src[(long long)g * s_gs + (long long)r * lds + c] : 0.f;
        tile[i][tx] = v;
        if (wc && r < rows && c < ldc) wc[(long long)g * c_gs + (long long)r * ldc + c] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    if (wt) {
        for (int i = ty; i < 32; i += 8) {
            const int c = c0 + i, r = r0 + tx;                        // Wt[c][r]
            if (c < cols && r < ldt) wt[(long long)g * t_gs + (long long)c * ldt + r] = __float2bfloat16_rn(tile[tx][i]);
        }
    }
}

__global__ void y_to_bf16_kernel(const uint8_t* __restrict__ y, int ldy, __nv_bfloat16* __restrict__ out, int ld, int B, int V) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * ld) return;
    const int b = (int)(i / ld), c = (int)(i - (long long)b * ld);
    out[i] = __float2bfloat16_rn(c < V && y[(long long)b * ldy + c] != 0 ? 1.0f : 0.0f);
}

// bias gradient: db[g][n] = sum_b dY[g][b][n]; one CTA per (64 columns, variable): 8 row groups x 32 column pairs
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dy, long long dy_gs, int lddy,
                                                          float* __restrict__ db, long long db_gs, int B, int N, int accum) {
    __shared__ float red[8][64];
    const int g = blockIdx.y, c = blockIdx.x * 64 + (threadIdx.x & 31) * 2, rg = threadIdx.x >> 5;
    float s0 = 0.f, s1 = 0.f;
    if (c < N) {
        const __nv_bfloat16* base = dy + (long long)g * dy_gs + c;
        for (int b = rg; b < B; b += 8) {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(base + (long long)b * lddy);   // c even, lddy even: aligned
            s0 += bf16_lo(u);
            s1 += bf16_hi(u);
        }
    }
    red[rg][(threadIdx.x & 31) * 2] = s0;
    red[rg][(threadIdx.x & 31) * 2 + 1] = s1;
    __syncthreads();
    if (threadIdx.x < 64) {
        const int n = blockIdx.x * 64 + threadIdx.x;
        if (n < N) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
            float* d = db + (long long)g * db_gs + n;
            *d = accum ? *d + s : s;
        }
    }
}

// ---- host ------------------------------------------------------------------------------------------------------
inline bool al16(const void* p) { return !((uintptr_t)p & 15); }

// tile width: a multiple of 16 (<= 256).  Cost of covering N with tiles of bn columns: the padded width plus a fixed
// per-tile overhead worth ~48 columns (the A tile is re-read for every N tile; narrow tiles are shared-memory bound)
int pick_bn(int N) {
    int best = 16;
    long long best_cost = -1;
    for (int bn = 256; bn >= 16; bn -= 16) {
        const long long cost = pg_cdiv(N, bn) * (bn + 48);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

struct Operand {
    const __nv_bfloat16* p; int64_t gs; int ld;     // rows x ld elements per group (gs == 0: shared)
};

// C[M,N] = A . B over K.  K-major operand: [rows = M or N][K] row-major.  MN-major operand: [K][M or N] row-major.
template <int EPI>
int launch(pgmvae_ctx* ctx, cudaStream_t st, Bf16P& p, const Operand& A, const Operand& B, const char* name, double bytes) {
    if (p.G <= 0 || p.M <= 0 || p.N <= 0 || p.K <= 0) return PGMVAE_OK;
    if (!al16(A.p) || !al16(B.p) || A.ld % 8 || B.ld % 8 || A.gs % 8 || B.gs % 8) {
        pgmvae_set_error("%s: bf16 operands must be 16-byte aligned with row strides that are multiples of 8", name);
        return PGMVAE_EINVAL;
    }
    p.BN = pick_bn(p.N);
    p.tiles_m = (int)pg_cdiv(p.M, TM);
    p.tiles_n = (int)pg_cdiv(p.N, p.BN);
    p.kblocks = (int)pg_cdiv(p.K, BK);
    const int64_t total = (int64_t)p.G * p.tiles_m * p.tiles_n;
    if (total > 0x7fffffff) {
        pgmvae_set_error("%s: too many tiles", name);
        return PGMVAE_EINVAL;
    }
    p.total_tiles = (int)total;
    p.a_shared = A.gs == 0; p.b_shared = B.gs == 0;
    p.a_bytes = A_BYTES;
    p.b_panels = (int)pg_cdiv(p.BN, 64);
    p.b_bytes = p.b_mn ? (unsigned)(p.b_panels * PANEL_BYTES) : (unsigned)pg_round_up(p.BN * 128, 1024);
    const size_t stage = (size_t)A_BYTES + p.b_bytes;
    const size_t fixed = 1024 + 256;
    int stages = (int)((ctx->smem_optin - fixed) / stage);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) {
        pgmvae_set_error("%s: shared memory too small", name);
        return PGMVAE_EINVAL;
    }
    p.stages = stages;
    // always more than half an SM's shared memory: one CTA per SM, which owns all 512 TMEM columns
    size_t smem = fixed + stages * stage;
    if (smem < (size_t)120 * 1024) smem = (size_t)120 * 1024;
    CUtensorMap mA, mB;
    if (!p.a_mn) PG_TRY(tc::make_map(&mA, A.p, 2, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)p.G, (uint64_t)A.ld, (uint64_t)A.gs, BK, TM));
    else PG_TRY(tc::make_map(&mA, A.p, 2, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.G, (uint64_t)A.ld, (uint64_t)A.gs, 64, BK));
    if (!p.b_mn) PG_TRY(tc::make_map(&mB, B.p, 2, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)p.G, (uint64_t)B.ld, (uint64_t)B.gs, BK, (uint32_t)p.BN));
    else PG_TRY(tc::make_map(&mB, B.p, 2, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.G, (uint64_t)B.ld, (uint64_t)B.gs, 64, BK));
    static size_t configured[16] = {};
    const int dev = ctx->device & 15;
    if (smem > configured[dev]) {
        PG_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    const int grid = p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count;
    PG_KERNEL(ctx, st, name, bytes, 2.0 * p.G * (double)p.M * p.N * p.K);
    gemm_bf16_kernel<EPI><<<grid, THREADS, smem, st>>>(mA, mB, p);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

}  // namespace

// ---- internal entry points (ops.cuh) ---------------------------------------------------------------------------
int pg_bf16_fwd(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* x, int64_t x_gs, int ldx, const __nv_bfloat16* wt,
                int64_t wt_gs, int ldwt, const float* bias, int64_t bias_gs, __nv_bfloat16* outb, int64_t outb_gs, int ldob,
                float* outf, int64_t outf_gs, int ldof, int G, int B, int in, int out_dim, int act) {
    Bf16P p{};
    p.G = G; p.M = B; p.N = out_dim; p.K = in;
    p.cb = outb; p.cb_gs = outb_gs; p.ldcb = ldob; p.cf = outf; p.cf_gs = outf_gs; p.ldcf = ldof;
    p.bias = bias; p.bias_gs = bias_gs; p.act = act;
    p.vec = (!outb || (al16(outb) && ldob % 8 == 0 && outb_gs % 8 == 0)) &&
            (!outf || (al16(outf) && ldof % 4 == 0 && outf_gs % 4 == 0)) && (!bias || (al16(bias) && bias_gs % 4 == 0));
    const double xg = x_gs == 0 ? 1.0 : (double)G;
    return launch<EPI_FWD>(ctx, st, p, Operand{x, x_gs, ldx}, Operand{wt, wt_gs, ldwt}, "dense_fwd_bf16",
                           2.0 * (xg * B * in + (double)G * in * out_dim) + 4.0 * G * out_dim +
                               (double)G * B * out_dim * ((outb ? 2.0 : 0.0) + (outf ? 4.0 : 0.0)));
}

int pg_bf16_fwd_sigmoid_mse(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* x, int64_t x_gs, int ldx,
                            const __nv_bfloat16* wt, int64_t wt_gs, int ldwt, const float* bias, int64_t bias_gs,
                            const __nv_bfloat16* yb, int ldyb, __nv_bfloat16* dpre, int64_t dpre_gs, int ldd, float* out_opt,
                            int64_t out_gs, int ldo, double* acc2, int G, int g0, int B, int in, int V, float grad_scale) {
    Bf16P p{};
    p.G = G; p.M = B; p.N = V; p.K = in;
    p.cb = dpre; p.cb_gs = dpre_gs; p.ldcb = ldd; p.cf = out_opt; p.cf_gs = out_gs; p.ldcf = ldo;
    p.bias = bias; p.bias_gs = bias_gs; p.yb = yb; p.ldyb = ldyb; p.acc = acc2; p.gscale = grad_scale; p.g0 = g0;
    p.vec = (!dpre || (al16(dpre) && ldd % 8 == 0 && dpre_gs % 8 == 0)) && al16(yb) && ldyb % 8 == 0 &&
            (!out_opt || (al16(out_opt) && ldo % 4 == 0 && out_gs % 4 == 0)) && (!bias || (al16(bias) && bias_gs % 4 == 0));
    return launch<EPI_SIGMOID_MSE>(ctx, st, p, Operand{x, x_gs, ldx}, Operand{wt, wt_gs, ldwt}, "dense_fwd_sigmoid_mse_bf16",
                                   2.0 * ((double)G * B * in + (double)G * in * V + (double)B * V + (double)G * B * V) +
                                       4.0 * G * V + (out_opt ? 4.0 * G * B * V : 0.0));
}

int pg_bf16_dgrad(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* dy, int64_t dy_gs, int lddy, const __nv_bfloat16* w,
                  int64_t w_gs, int ldw, const __nv_bfloat16* hb, int64_t hb_gs, int ldhb, const float* hf, int64_t hf_gs,
                  int ldhf, const float* z, const float* q, int64_t zq_gs, int ldzq, float cscale, __nv_bfloat16* dxb,
                  int64_t dxb_gs, int lddxb, float* dxf, int64_t dxf_gs, int lddxf, int G, int B, int in, int out_dim,
                  int act_below) {
    Bf16P p{};
    p.G = G; p.M = B; p.N = in; p.K = out_dim;
    p.cb = dxb; p.cb_gs = dxb_gs; p.ldcb = lddxb; p.cf = dxf; p.cf_gs = dxf_gs; p.ldcf = lddxf;
    p.hb = hb; p.hb_gs = hb_gs; p.ldhb = ldhb; p.hf = hf; p.hf_gs = hf_gs; p.ldhf = ldhf; p.act = act_below;
    p.z = z; p.q = q; p.zq_gs = zq_gs; p.ldzq = ldzq; p.cscale = cscale;
    p.vec = (!dxb || (al16(dxb) && lddxb % 8 == 0 && dxb_gs % 8 == 0)) && (!dxf || (al16(dxf) && lddxf % 4 == 0 && dxf_gs % 4 == 0)) &&
            (!hb || (al16(hb) && ldhb % 8 == 0 && hb_gs % 8 == 0)) && (!hf || (al16(hf) && ldhf % 4 == 0 && hf_gs % 4 == 0)) &&
            (!z || (al16(z) && al16(q) && ldzq % 4 == 0 && zq_gs % 4 == 0));
    return launch<EPI_DGRAD>(ctx, st, p, Operand{dy, dy_gs, lddy}, Operand{w, w_gs, ldw}, "dense_dgrad_bf16",
                             2.0 * ((double)G * B * out_dim + (double)G * in * out_dim + (double)G * B * in * 2.0) +
                                 (z ? 8.0 * G * B * in : 0.0));
}

// dW[in][out] (+)= X^T dY; the orientation that pads less.  db is NOT computed here (pg_bf16_colsum).
int pg_bf16_wgrad(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* x, int64_t x_gs, int ldx, const __nv_bfloat16* dy,
                  int64_t dy_gs, int lddy, float* dw, int64_t dw_gs, int lddw, int G, int B, int in, int out_dim,
                  int zero_row_base, int accumulate) {
    auto padded = [](int m, int n) { return (double)pg_round_up(m, TM) * (double)(pg_cdiv(n, pick_bn(n)) * (pick_bn(n) + 48)); };
    bool direct = padded(in, out_dim) <= padded(out_dim, in);
    if (const char* ev = getenv("PGMVAE_WGRAD_ORIENT")) direct = ev[0] == 'd' ? true : (ev[0] == 't' ? false : direct);
    Bf16P p{};
    p.G = G; p.K = B; p.a_mn = 1; p.b_mn = 1;
    p.dw = dw; p.dw_gs = dw_gs; p.lddw = lddw; p.zero_row_base = zero_row_base >= 0 ? zero_row_base : -(1 << 30);
    p.vec = al16(dw) && lddw % 4 == 0 && dw_gs % 4 == 0;
    p.accum = accumulate;
    const double xg = x_gs == 0 ? 1.0 : (double)G;
    const double bytes = 2.0 * (xg * B * in + (double)G * B * out_dim) + 4.0 * G * (double)in * out_dim;
    if (direct) {
        p.M = in; p.N = out_dim;
        return launch<EPI_WGRAD_D>(ctx, st, p, Operand{x, x_gs, ldx}, Operand{dy, dy_gs, lddy}, "dense_wgrad_bf16", bytes);
    }
    p.M = out_dim; p.N = in;
    return launch<EPI_WGRAD_T>(ctx, st, p, Operand{dy, dy_gs, lddy}, Operand{x, x_gs, ldx}, "dense_wgrad_bf16", bytes);
}

int pg_bf16_colsum(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* dy, int64_t dy_gs, int lddy, float* db, int64_t db_gs,
                   int G, int B, int N, int accumulate) {
    if (G <= 0 || N <= 0) return PGMVAE_OK;
    if (((uintptr_t)dy & 3) || lddy % 2 || dy_gs % 2) {
        pgmvae_set_error("colsum (bf16): rows must be 4-byte aligned");
        return PGMVAE_EINVAL;
    }
    dim3 grid((unsigned)pg_cdiv(N, 64), (unsigned)G);
    PG_KERNEL(ctx, st, "bias_grad_colsum_bf16", 2.0 * G * (double)B * N + 4.0 * G * N, (double)G * B * N);
    colsum_bf16_kernel<<<grid, 256, 0, st>>>(dy, dy_gs, lddy, db, db_gs, B, N, accumulate);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

int pg_f32_to_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* src, int64_t s_gs, int lds, __nv_bfloat16* dst, int64_t d_gs,
                   int ldd, int G, int rows, int cols) {
    if (G <= 0 || rows <= 0) return PGMVAE_OK;
    const long long n = (long long)rows * ldd;
    int bx = (int)std::min<long long>(pg_cdiv(n, 256 * 4), 4096);
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)G);
    PG_KERNEL(ctx, st, "f32_to_bf16", 6.0 * G * (double)rows * cols, 0.0);
    f32_to_bf16_kernel<<<grid, 256, 0, st>>>(src, s_gs, lds, dst, d_gs, ldd, rows, cols);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

// bf16 shadows of a weight tensor W [G][rows][lds]: wt [G][cols][ldt] (transposed) and / or wc [G][rows][ldc]
int pg_bf16_shadow(pgmvae_ctx* ctx, cudaStream_t st, const float* w, int64_t w_gs, int lds, int rows, int cols,
                   __nv_bfloat16* wt, int64_t t_gs, int ldt, __nv_bfloat16* wc, int64_t c_gs, int ldc, int G) {
    if (G <= 0) return PGMVAE_OK;
    // cover the padded extents so that the pad of both shadows is (re)written as zero
    const int rr = wt ? std::max(rows, ldt) : rows, cc = wc ? std::max(cols, ldc) : cols;
    for (int g0 = 0; g0 < G; g0 += 65535) {
        const int gn = std::min(65535, G - g0);
        dim3 grid((unsigned)pg_cdiv(cc, 32), (unsigned)pg_cdiv(rr, 32), (unsigned)gn);
        PG_KERNEL(ctx, st, "bf16_shadow", (double)gn * rows * cols * (4.0 + (wt ? 2.0 : 0.0) + (wc ? 2.0 : 0.0)), 0.0);
        shadow_kernel<<<grid, 256, 0, st>>>(w + (size_t)g0 * w_gs, w_gs, lds, rows, cols, wt ? wt + (size_t)g0 * t_gs : nullptr,
                                            t_gs, ldt, wc ? wc + (size_t)g0 * c_gs : nullptr, c_gs, ldc);
        PG_LAUNCHED(ctx);
    }
    return PGMVAE_OK;
}

int pg_y_to_bf16(pgmvae_ctx* ctx, cudaStream_t st, const uint8_t* y, int ldy, __nv_bfloat16* out, int ld, int B, int V) {
    if (B <= 0) return PGMVAE_OK;
    const long long n = (long long)B * ld;
    PG_KERNEL(ctx, st, "y_to_bf16", (double)B * V + 2.0 * n, 0.0);
    y_to_bf16_kernel<<<(unsigned)pg_cdiv(n, 256), 256, 0, st>>>(y, ldy, out, ld, B, V);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}
