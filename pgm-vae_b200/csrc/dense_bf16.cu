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
// the current one is processed; bf16 rows leave / arrive as staged TMA tiles, the tile's bias is
// staged in shared memory once per tile, fp32 rows leave through a transposing tile):
//   FWD          bias + selu / sigmoid / none -> bf16 row (next layer's operand) and/or fp32 row
//   SIGMOID_MSE  fd9: bias + sigmoid + squared / absolute error sums + d(loss)/d(pre-activation) in
//                bf16 (leave-one-out column masked), core/model.py:53 + run.py:61
//   DGRAD        (+ commitment gradient at the VQ boundary) x act'(activation below) -> bf16 row
//   WGRAD_D / WGRAD_T  fp32 weight gradient
// Every instantiation exists twice: the full-register one, and a SLIM one (144 registers, 3 KB less
// shared memory) that leaves room on the SM for a CTA of the data-parallel exchange kernel
// (model.cu: p2p_shard_adam_kernel) next to it.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

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
constexpr int ONES_COL = 240;               // column of the bias-gradient accumulator (WGRAD_T, BN <= 240)
constexpr int ONES_BYTES = 2048;            // [16 n][64 k] bf16 tile of 1.0 (K-major)

enum { EPI_FWD = 0, EPI_SIGMOID_MSE = 1, EPI_DGRAD = 2, EPI_WGRAD_D = 3, EPI_WGRAD_T = 4 };

struct Bf16P {
    int G, M, N, K, BN, tiles_m, tiles_n, kblocks, stages, total_tiles;
    int a_mn, b_mn, a_shared, b_shared, b_panels;
    int pair, total_ptiles, tiles_mp, tiles_np;           // 2-CTA clusters: 0 none, 1 pair along M (B multicast), 2 pair along N (A multicast),
                                                          // 3 pair along M with ONE cta_group::2 MMA over both SMs (each holds half of B),
                                                          // 4 cluster of 2 x 2 tiles: A multicast along N, B multicast along M
    unsigned a_bytes, b_bytes;
    int vec;                                              // rows allow 16-byte vector access
    int tma_out, tma_in, in_shared;                       // bf16 output / epilogue operand move through staged TMA tiles
    int out_db;                                           // dgrad: two output staging tiles next to the two operand tiles
    int stage_f32;                                        // forward: fp32 rows leave through a transposing 16 KB staging tile
    __nv_bfloat16* cb; long long cb_gs; int ldcb;         // bf16 output rows (may be null)
    float* cf; long long cf_gs; int ldcf;                 // fp32 output rows (may be null)
    const float* bias; long long bias_gs; int act;
    const uint32_t* ybits; int ldbits; double* acc; float gscale; int g0;      // SIGMOID_MSE: targets bit-packed, 32 columns per word
    const __nv_bfloat16* hb; long long hb_gs; int ldhb;   // DGRAD: activation below (bf16) ...
    const float* hf; long long hf_gs; int ldhf;           // ... or fp32
    const float* z; const float* q; long long zq_gs; int ldzq; float cscale;
    float* dw; long long dw_gs; int lddw; int zero_row_base, accum;           // WGRAD (accum: dW += instead of =)
    float* db; long long db_gs; int ones;                 // WGRAD_T: bias gradient = row sums of A through a ones-tile MMA
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

// Tensor maps of the epilogue's bf16 tiles: one map PER N TILE (the tensor seen through it starts at the tile's first
// column and is as wide as the tile), so that TMA clips the columns a 32-wide chunk overhangs the tile as well as the
// ragged edges of the tensor.  [64-byte swizzle, box = 128 rows x 32 columns]
constexpr int MAX_NT_MAPS = 8;
struct EpiMaps {
    CUtensorMap out[MAX_NT_MAPS];      // bf16 output rows (activation / gradient)
    CUtensorMap in[MAX_NT_MAPS];       // bf16 epilogue operand rows (targets of the MSE stage / activation below)
};
constexpr int STG_TILE = TM * 64;      // one staged chunk: 128 rows x 32 bf16 (64 B, 64-byte swizzle) = 8 KB
constexpr int STG_WG = 3 * STG_TILE;   // per epilogue warpgroup: out | in[0] | in[1]
constexpr int STG_WG4 = 4 * STG_TILE;  // ... out[0] | out[1] | in[0] | in[1] (dgrad of the short-K layers: p.out_db)

// (variable, M tile, N tile) of pair-tile `pt` for the CTA of rank `crank` in its cluster.  N fastest, then M, then
// the variable; the two CTAs of a pair take neighbouring M tiles of one N tile (pair == 1: they share the B tile) or
// neighbouring N tiles of one M tile (pair == 2: they share the A tile).  Tiles past the edge are phantoms: their
// loads are zero-filled and nothing of them is stored, but they keep the pair in lock step.
__device__ __forceinline__ void decode_tile(const Bf16P& p, int pt, int crank, int& g, int& mt, int& nt) {
    if (p.pair == 1 || p.pair == 3) {
        const int per_g = p.tiles_mp * p.tiles_n;
        g = pt / per_g;
        const int r = pt - g * per_g;
        const int mtp = r / p.tiles_n;
        nt = r - mtp * p.tiles_n;
        mt = 2 * mtp + crank;
    } else if (p.pair == 4) {
        const int per_g = p.tiles_mp * p.tiles_np;
        g = pt / per_g;
        const int r = pt - g * per_g;
        const int mtp = r / p.tiles_np;
        mt = 2 * mtp + (crank >> 1);
        nt = 2 * (r - mtp * p.tiles_np) + (crank & 1);
    } else if (p.pair == 2) {
        const int per_g = p.tiles_m * p.tiles_np;
        g = pt / per_g;
        const int r = pt - g * per_g;
        mt = r / p.tiles_np;
        nt = 2 * (r - mt * p.tiles_np) + crank;
    } else {
        const int per_g = p.tiles_m * p.tiles_n;
        g = pt / per_g;
        const int r = pt - g * per_g;
        mt = r / p.tiles_n;
        nt = r - mt * p.tiles_n;
    }
}

// CTA2: the instantiation that contains the cta_group::2 instructions (pair == 3).  It is a separate kernel because a
// kernel holding such instructions can only be launched as clusters of CTA pairs ("cluster misconfiguration" otherwise).
// SLIM: 144 registers x 384 threads and 3 KB of shared memory less than the maximum leave room on every SM for one CTA
// of the data-parallel exchange kernel (model.cu: p2p_shard_adam_kernel, 128 threads x 80 registers) NEXT TO this one
// (ctx->coresident, set by the training step while that exchange is in use).  The cap costs the epilogues a few spills
// (the MSE stage ~25 %), so everything else runs the full-register instantiation.
template <int EPI, bool CTA2, bool SLIM>
__global__ void __maxnreg__(SLIM ? 144 : 168)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ EpiMaps em, const __grid_constant__ Bf16P p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, shared space
    uint8_t* sA = smem;                                           // [stages][16 KB]
    uint8_t* sB = sA + (size_t)p.stages * A_BYTES;                // [stages][b_bytes]
    uint8_t* sStage = sB + (size_t)p.stages * p.b_bytes;          // [2 warpgroups][STG_WG] (only with tma_out / tma_in)
    uint8_t* sOnes = sStage + ((p.tma_out || p.tma_in || p.stage_f32 || EPI == EPI_SIGMOID_MSE) ? 2 * (p.out_db ? STG_WG4 : STG_WG) : 0);      // (only with p.ones)
    float* sBias = reinterpret_cast<float*>(sOnes + (p.ones ? ONES_BYTES : 0));      // [2 warpgroups][256] bias of the tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 2 * ACC_COLS);
    uint64_t* full = bars;                        // [MAX_STAGES]   TMA -> MMA
    uint64_t* empty = bars + MAX_STAGES;          // [MAX_STAGES]   MMA -> TMA
    uint64_t* tmem_full = bars + 2 * MAX_STAGES;  // [2]            MMA -> epilogue warpgroup
    uint64_t* tmem_empty = tmem_full + 2;         // [2]            epilogue warpgroup -> MMA
    uint64_t* in_full = tmem_empty + 2;           // [2 warpgroups][2 buffers]  epilogue operand chunk has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&mapA);
        tc::tma_prefetch_desc(&mapB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            // multicast pairs: a stage is free when BOTH CTAs' MMAs have read it.  cta_group::2: the leader's full barrier
            // collects the producers of both CTAs, and its one MMA stream frees the stage in both.
            tc::mbar_init(&full[s], CTA2 ? 2 : 1);
            tc::mbar_init(&empty[s], p.pair == 4 ? 3 : ((p.pair == 1 || p.pair == 2) ? 2 : 1));
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tmem_full[b], 1);
            tc::mbar_init(&tmem_empty[b], CTA2 ? 8 : 4);      // one arrival per warp of the warpgroup (of both CTAs)
        }
        for (int b = 0; b < 4; ++b) tc::mbar_init(&in_full[b], 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        if (CTA2) {
            tc::tmem_alloc_2sm(tmem_slot, 512u);
            tc::tmem_relinquish_2sm();
        } else {
            tc::tmem_alloc(tmem_slot, 512u);
            tc::tmem_relinquish();
        }
    }
    if (EPI == EPI_WGRAD_T && p.ones && warp == 3) {
        // B tile of ones: D2[m][0..16) += A[m][k] * 1 leaves sum_k A[m][k] -- the bias gradient when A = dY^T -- in
        // 16 spare accumulator columns, for 8 extra tensor cycles per k-step on one N tile (all-ones is swizzle-invariant)
        for (int i = lane; i < ONES_BYTES / 16; i += 32)
            reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        tc::fence_proxy_async_smem();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (p.pair) tc::cluster_sync_all();                   // the peer's barriers exist before anything is multicast to them
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int crank = p.pair ? (int)tc::cluster_ctarank() : 0;
    const int csize = p.pair == 4 ? 4 : (p.pair ? 2 : 1);
    // who shares which operand with this CTA (multicast masks over cluster ranks).  2 x 2 cluster: rank = 2 rm + rn
    const int rm = p.pair == 4 ? crank >> 1 : (p.pair == 1 ? crank : 0);       // position along M = which half of B this CTA loads
    const int rn = p.pair == 4 ? crank & 1 : (p.pair == 2 ? crank : 0);        // position along N = which half of A this CTA loads
    const uint16_t a_mask = p.pair == 4 ? (uint16_t)(3u << (2 * rm)) : 3;      // CTAs with the same M tile
    const uint16_t b_mask = p.pair == 4 ? (uint16_t)(5u << rn) : 3;            // CTAs with the same N tile
    const bool share_a = p.pair == 2 || p.pair == 4, share_b = p.pair == 1 || p.pair == 4;
    const int cid = (int)blockIdx.x / csize, ncl = (int)gridDim.x / csize;

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        const uint32_t stage_tx = p.a_bytes + p.b_bytes;
        // pairs: this CTA loads ITS HALF of the shared operand and multicasts it into both CTAs (rows of a K-major
        // tile, 64-wide panels of an MN-major one); every CTA still receives -- and expects -- the whole stage
        const int a_rows_half = TM / 2, b_rows_half = p.BN / 2;
        const int bp0 = share_b ? (rm == 0 ? 0 : (p.b_panels + 1) / 2) : 0;
        const int bp1 = share_b ? (rm == 0 ? (p.b_panels + 1) / 2 : p.b_panels) : p.b_panels;
        uint32_t s = 0, ph = 0;
        for (int pt = cid; pt < p.total_ptiles; pt += ncl) {
            int g, mt, nt;
            decode_tile(p, pt, crank, g, mt, nt);
            const int m0 = mt * TM, n0 = nt * p.BN;
            const int ga = p.a_shared ? 0 : g, gb = p.b_shared ? 0 : g;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                tc::mbar_wait(&empty[s], ph ^ 1);
                if (CTA2) {
                    // cta_group::2: own A tile and own HALF of the B tile into own shared memory; the bytes of both CTAs
                    // complete on the leader's barrier, which the leader arms for both
                    if (tc::elect_one()) {
                        if (crank == 0) tc::mbar_arrive_expect_tx(&full[s], 2 * stage_tx);
                        else tc::mbar_arrive_cluster(&full[s], 0);
                        uint8_t* a_dst = sA + (size_t)s * A_BYTES;
                        uint8_t* b_dst = sB + (size_t)s * p.b_bytes;
                        if (!p.a_mn) {
                            tc::tma_load_3d_2sm(a_dst, &mapA, &full[s], kb * BK, m0, ga);
                        } else {
                            tc::tma_load_3d_2sm(a_dst, &mapA, &full[s], m0, kb * BK, ga);
                            tc::tma_load_3d_2sm(a_dst + PANEL_BYTES, &mapA, &full[s], m0 + 64, kb * BK, ga);
                        }
                        const int nh = n0 + crank * (p.BN / 2);
                        if (!p.b_mn) {
                            tc::tma_load_3d_2sm(b_dst, &mapB, &full[s], kb * BK, nh, gb);             // [BN/2 n][64 k]
                        } else {
                            for (int pn = 0; pn < p.b_panels; ++pn)                                     // panels of the half
                                tc::tma_load_3d_2sm(b_dst + (size_t)pn * PANEL_BYTES, &mapB, &full[s], nh + pn * 64, kb * BK, gb);
                        }
                    }
                    __syncwarp();
                    if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
                    continue;
                }
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&full[s], stage_tx);
                    uint8_t* a_dst = sA + (size_t)s * A_BYTES;
                    uint8_t* b_dst = sB + (size_t)s * p.b_bytes;
                    if (share_a) {                         // A shared along N: this CTA loads half rn for all of them
                        if (!p.a_mn)
                            tc::tma_load_3d_mc(a_dst + rn * a_rows_half * 128, &mapA, &full[s], kb * BK, m0 + rn * a_rows_half, ga,
                                               a_mask);
                        else
                            tc::tma_load_3d_mc(a_dst + rn * PANEL_BYTES, &mapA, &full[s], m0 + rn * 64, kb * BK, ga, a_mask);
                    } else if (!p.a_mn) {
                        tc::tma_load_3d(a_dst, &mapA, &full[s], kb * BK, m0, ga);                   // [128 m][64 k]
                    } else {
                        tc::tma_load_3d(a_dst, &mapA, &full[s], m0, kb * BK, ga);                   // 2 panels [64 k][64 m]
                        tc::tma_load_3d(a_dst + PANEL_BYTES, &mapA, &full[s], m0 + 64, kb * BK, ga);
                    }
                    if (share_b) {                         // B shared along M: this CTA loads half rm for all of them
                        if (!p.b_mn) {
                            tc::tma_load_3d_mc(b_dst + rm * b_rows_half * 128, &mapB, &full[s], kb * BK, n0 + rm * b_rows_half, gb,
                                               b_mask);
                        } else {
                            for (int pn = bp0; pn < bp1; ++pn)
                                tc::tma_load_3d_mc(b_dst + (size_t)pn * PANEL_BYTES, &mapB, &full[s], n0 + pn * 64, kb * BK, gb, b_mask);
                        }
                    } else if (!p.b_mn) {
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
        const uint32_t idesc = tc::make_idesc(1, CTA2 ? 2 * TM : TM, p.BN, p.a_mn, p.b_mn);
        // K-major : rows of 128 B along k, 8-row atoms 1024 B apart (SBO); one MMA (k = 16) advances 32 B
        // MN-major: panels [64 k][128 B along m/n], 8-k atoms 1024 B apart (SBO), panels PANEL_BYTES apart
        //           (LBO); one MMA (k = 16) advances two atoms = 2048 B
        const uint64_t dA0 = p.a_mn ? tc::make_smem_desc(tc::smem_u32(sA), PANEL_BYTES, 1024, 2)
                                    : tc::make_smem_desc(tc::smem_u32(sA), 16, 1024, 2);
        const uint64_t dB0 = p.b_mn ? tc::make_smem_desc(tc::smem_u32(sB), PANEL_BYTES, 1024, 2)
                                    : tc::make_smem_desc(tc::smem_u32(sB), 16, 1024, 2);
        const uint32_t a_step = (p.a_mn ? 2048u : 32u) >> 4, b_step = (p.b_mn ? 2048u : 32u) >> 4;
        const uint32_t a_stage = (uint32_t)A_BYTES >> 4, b_stage = p.b_bytes >> 4;
        const uint32_t idesc1 = tc::make_idesc(1, CTA2 ? 2 * TM : TM, 16, p.a_mn, 0);
        const uint64_t dOnes = tc::make_smem_desc(tc::smem_u32(sOnes), 16, 1024, 2);
        uint32_t s = 0, ph = 0, it = 0;
        for (int pt = cid; pt < p.total_ptiles && !(CTA2 && crank != 0); pt += ncl, ++it) {     // cta_group::2: the leader issues for both
            const uint32_t buf = it & 1, use = it >> 1;
            int g, mt, nt;
            decode_tile(p, pt, crank, g, mt, nt);
            const bool ones = EPI == EPI_WGRAD_T && p.ones && nt == 0;
            tc::mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);          // the epilogue has drained this buffer
            tc::fence_after_thread_sync();
            const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                tc::mbar_wait(&full[s], ph);
                tc::fence_after_thread_sync();
                if (CTA2) {
                    if (tc::elect_one()) {
                        const uint64_t dA = dA0 + (uint64_t)(s * a_stage), dB = dB0 + (uint64_t)(s * b_stage);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc::mma_f16_2sm(d_tmem, dA + (uint64_t)(k4 * a_step), dB + (uint64_t)(k4 * b_step), idesc,
                                            (kb > 0 || k4 > 0) ? 1u : 0u);
                        if (ones) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                tc::mma_f16_2sm(d_tmem + ONES_COL, dA + (uint64_t)(k4 * a_step), dOnes + (uint64_t)(k4 * 2), idesc1,
                                                (kb > 0 || k4 > 0) ? 1u : 0u);
                        }
                        tc::mma_commit_2sm(&empty[s], 3);                                   // frees the stage in both CTAs
                        if (kb == p.kblocks - 1) tc::mma_commit_2sm(&tmem_full[buf], 3);   // both CTAs' epilogues
                    }
                    __syncwarp();
                    if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
                    continue;
                }
                if (tc::elect_one()) {
                    const uint64_t dA = dA0 + (uint64_t)(s * a_stage), dB = dB0 + (uint64_t)(s * b_stage);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        tc::mma_f16(d_tmem, dA + (uint64_t)(k4 * a_step), dB + (uint64_t)(k4 * b_step), idesc,
                                    (kb > 0 || k4 > 0) ? 1u : 0u);
                    if (ones) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc::mma_f16(d_tmem + ONES_COL, dA + (uint64_t)(k4 * a_step), dOnes + (uint64_t)(k4 * 2), idesc1,
                                        (kb > 0 || k4 > 0) ? 1u : 0u);
                    }
                    if (p.pair) tc::mma_commit_mc(&empty[s], p.pair == 4 ? (uint16_t)(a_mask | b_mask) : (uint16_t)3);
                    else tc::mma_commit(&empty[s]);
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
        const int rt = qd * 32 + lane;                 // row within the tile
        const bool leader = rt == 0;                   // issues this warpgroup's TMA loads / stores
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + wg * ACC_COLS;
        // staged chunk [128 rows][64 B], 64-byte swizzle: the 16-byte piece j of row r sits at piece j ^ ((r >> 1) & 3)
        // per warpgroup: out | in[0] | in[1]; the MSE stage needs no operand tiles (its targets are bits) and uses the
        // room for a second output tile and its target words: out[0] | out[1] | words
        uint8_t* sOut = sStage + (size_t)wg * (p.out_db ? STG_WG4 : STG_WG);
        uint8_t* sIn = sOut + (p.out_db ? 2 : 1) * STG_TILE;
        uint32_t* sY = reinterpret_cast<uint32_t*>(sOut + 2 * STG_TILE) + rt * 9;       // 9 target words of this thread's row
        const uint32_t swz = (uint32_t)((rt >> 1) & 3);
        uint64_t* my_in_full = in_full + wg * 2;
        uint32_t in_uses[2] = {0u, 0u};
        uint32_t out_cnt = 0;
        const bool tin = p.tma_in != 0, tout = p.tma_out != 0;
        double dsq = 0.0, dab = 0.0;
        // hand an accumulator buffer back to the MMA issuer: under cta_group::2 that is the leader CTA's warp 1 for both CTAs
        auto release_acc = [&]() {
            if (CTA2) tc::mbar_arrive_cluster(&tmem_empty[wg], 0);
            else tc::mbar_arrive(&tmem_empty[wg]);
        };
        uint32_t it = 0;
        for (int pt = cid; pt < p.total_ptiles; pt += ncl, ++it) {
            if ((int)(it & 1) != wg) continue;
            const uint32_t use = it >> 1;
            int g, mt, nt;
            decode_tile(p, pt, crank, g, mt, nt);
            const int m0 = mt * TM, n0 = nt * p.BN;
            const int row = m0 + rt;
            const bool rvalid = row < p.M;
            const int ncols = min(p.BN, p.N - n0);
            const int nch = (ncols + 31) >> 5;
            if (ncols <= 0) {                              // phantom N tile of a pair: nothing to store, hand the buffer back
                tc::mbar_wait_parked(&tmem_full[wg], use & 1);      // parked: eight polling epilogue warps would take issue slots from the producer / MMA warps
                tc::fence_after_thread_sync();
                tc::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) release_acc();
                continue;
            }
            const int gin = p.in_shared ? 0 : g;
            float sq = 0.f, ab = 0.f;
            // the tile's bias through shared memory: two predicated loads per thread and tile instead of 32 bounds-checked
            // ones per chunk (which were 40 % of the instructions of a chunk).  The MSE stage keeps -log2(e) * bias: its
            // sigmoid is 1 / (1 + 2^(-log2(e) * (acc + bias))).  The warpgroup's reads of the previous tile's bias
            // precede the barriers of that tile's last write_out.
            const bool sf32 = EPI == EPI_FWD && p.stage_f32 != 0;
            const bool sbias = (tout || sf32) && p.bias && (EPI == EPI_FWD || EPI == EPI_SIGMOID_MSE);
            float* myBias = sBias + wg * ACC_COLS;
            if (sbias) {
                const float* bp = p.bias + (long long)g * p.bias_gs + n0;
                const float sc = EPI == EPI_SIGMOID_MSE ? -LOG2E : 1.0f;
                myBias[rt] = rt < ncols ? sc * __ldg(bp + rt) : 0.f;
                myBias[rt + TM] = rt + TM < ncols ? sc * __ldg(bp + rt + TM) : 0.f;
                tc::named_bar_sync(1 + wg, 128);
            }

            // epilogue operand (targets / activation below) of chunk c -> staging buffer c & 1
            auto fetch_in = [&](int c) {
                if (leader) {
                    tc::mbar_arrive_expect_tx(&my_in_full[c & 1], (uint32_t)STG_TILE);
                    tc::tma_load_3d(sIn + (size_t)(c & 1) * STG_TILE, &em.in[nt], &my_in_full[c & 1], c * 32, m0, gin);
                }
            };
            // the thread's 32 operand values of chunk c (bf16 -> fp32)
            auto read_in = [&](int c, float (&t)[32]) {
                const int b = c & 1;
                tc::mbar_wait(&my_in_full[b], in_uses[b] & 1);
                ++in_uses[b];
                const uint8_t* rowp = sIn + (size_t)b * STG_TILE + rt * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 u = *reinterpret_cast<const uint4*>(rowp + ((j ^ swz) << 4));
                    t[8 * j] = bf16_lo(u.x); t[8 * j + 1] = bf16_hi(u.x); t[8 * j + 2] = bf16_lo(u.y); t[8 * j + 3] = bf16_hi(u.y);
                    t[8 * j + 4] = bf16_lo(u.z); t[8 * j + 5] = bf16_hi(u.z); t[8 * j + 6] = bf16_lo(u.w); t[8 * j + 7] = bf16_hi(u.w);
                }
            };
            // the thread's 32 bf16 results of chunk c -> staging tile -> one TMA store per warpgroup
            auto write_out = [&](int c, const float (&v)[32]) {
                // (forward stages have no operand tiles: the room holds a second output tile)
                const bool two_out = EPI == EPI_SIGMOID_MSE || EPI == EPI_FWD || p.out_db;
                uint8_t* tile = sOut + (two_out ? (size_t)(out_cnt++ & 1u) * STG_TILE : 0);     // alternates ACROSS tiles as well
                if (leader) {                                       // the store that last used this staging tile has read it
                    if (two_out) tc::bulk_wait_read1();
                    else tc::bulk_wait_read();
                }
                tc::named_bar_sync(1 + wg, 128);
                uint8_t* rowp = tile + rt * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(rowp + ((j ^ swz) << 4)) =
                        make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                   pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                tc::fence_proxy_async_smem();
                tc::named_bar_sync(1 + wg, 128);
                if (leader) {
                    tc::tma_store_3d(&em.out[nt], tile, c * 32, m0, g);
                    tc::bulk_commit();
                }
            };

            // fp32 rows [128][32 floats] through a staging tile of 16 KB (16-byte piece j of row r at j ^ (r & 7)): a thread
            // owns a row, so direct 16-byte accesses touch 32 lines per warp instruction; through the tile a warp moves four
            // whole 128-byte rows per instruction (fd4's latent: 0.118 -> 0.076 ms per group at cfg3).  `tile16k` must not
            // hold anything else of this warpgroup.  (The same for the fp32 z / q rows READ by the dgrad at the VQ boundary
            // was measured slower -- 0.34 vs 0.22 ms, four more barriers per chunk in a latency-bound tile loop -- and removed.)
            auto f32_rows_out = [&](uint8_t* tile16k, float* dst, int ld, int c, const float (&v)[32]) {
                const int nv = min(32, ncols - c * 32);
                tc::named_bar_sync(1 + wg, 128);                    // the tile's previous contents have been read
                float4* rowp = reinterpret_cast<float4*>(tile16k + rt * 128);
#pragma unroll
                for (int j = 0; j < 8; ++j) rowp[j ^ (rt & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                tc::named_bar_sync(1 + wg, 128);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = rt + 128 * k, r = i >> 3, pc = i & 7;
                    if (m0 + r >= p.M) continue;
                    const float4 val = reinterpret_cast<const float4*>(tile16k + r * 128)[pc ^ (r & 7)];
                    float* d = dst + (long long)(m0 + r) * ld + n0 + c * 32 + pc * 4;
                    if (pc * 4 + 4 <= nv) *reinterpret_cast<float4*>(d) = val;
                    else {
                        const float e[4] = {val.x, val.y, val.z, val.w};
                        for (int q4 = 0; q4 < 4; ++q4) if (pc * 4 + q4 < nv) d[q4] = e[q4];
                    }
                }
            };
            auto process = [&](float (&v)[32], int c) {
                const int nb = n0 + c * 32;
                const int nv = min(32, ncols - c * 32);            // valid columns of this chunk (tile and tensor bounds)
                if (EPI == EPI_FWD) {
                    if (sbias) {
                        const float4* b4 = reinterpret_cast<const float4*>(myBias + c * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = b4[j];
                            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                        }
                    } else if (p.bias) {
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
                    if (sf32) f32_rows_out(sOut, p.cf + (long long)g * p.cf_gs, p.ldcf, c, v);      // (out | in[0]: no operand tiles here)
                    else if (p.cf && rvalid) store_f32_row(p.cf + (long long)g * p.cf_gs + (long long)row * p.ldcf + nb, v, nv, p.vec);
                    if (tout) write_out(c, v);
                    else if (p.cb && rvalid) store_bf16_row(p.cb + (long long)g * p.cb_gs + (long long)row * p.ldcb + nb, v, nv, p.vec);
                } else if (EPI == EPI_SIGMOID_MSE) {
                    float t[32];
                    // v <- -log2(e) * (acc + bias)
                    if (sbias) {
                        const float4* b4 = reinterpret_cast<const float4*>(myBias + c * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = b4[j];
                            v[4 * j] = fmaf(v[4 * j], -LOG2E, b.x); v[4 * j + 1] = fmaf(v[4 * j + 1], -LOG2E, b.y);
                            v[4 * j + 2] = fmaf(v[4 * j + 2], -LOG2E, b.z); v[4 * j + 3] = fmaf(v[4 * j + 3], -LOG2E, b.w);
                        }
                    } else {
                        if (p.bias) {
                            load_f32_row(p.bias + (long long)g * p.bias_gs + nb, t, nv, p.vec);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += t[j];
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= -LOG2E;
                    }
                    // targets (0 / 1): bit j of `bits` belongs to column nb + j
                    const uint32_t bits = (n0 & 31) ? __funnelshift_r(sY[c], sY[c + 1], n0 & 31) : sY[c];
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] = (bits >> j) & 1u ? 1.0f : 0.0f;
                    const int self = rvalid ? p.g0 + g - nb : -1;                           // masked column of this net
                    const int nvr = rvalid ? nv : 0;                                        // rows past the batch count nothing
                    float o[32];
                    if (nvr == 32 && (unsigned)self >= 32u) {                               // interior chunk: nothing to mask
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            o[j] = rcp_approx(1.0f + ex2_approx(v[j]));
                            const float d = o[j] - t[j];
                            sq = fmaf(d, d, sq);
                            ab += fabsf(d);
                            const float u = p.gscale * o[j];
                            v[j] = d * fmaf(-u, o[j], u);                                   // gscale * d * o * (1 - o)
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            o[j] = rcp_approx(1.0f + ex2_approx(v[j]));
                            float d = o[j] - t[j];
                            if (j >= nvr || j == self) d = 0.f;
                            sq = fmaf(d, d, sq);
                            ab += fabsf(d);
                            const float u = p.gscale * o[j];
                            v[j] = d * fmaf(-u, o[j], u);
                        }
                    }
                    if (p.cf && rvalid) store_f32_row(p.cf + (long long)g * p.cf_gs + (long long)row * p.ldcf + nb, o, nv, p.vec);
                    if (tout) write_out(c, v);
                    else if (p.cb && rvalid) store_bf16_row(p.cb + (long long)g * p.cb_gs + (long long)row * p.ldcb + nb, v, nv, p.vec);
                } else if (EPI == EPI_DGRAD) {
                    float t[32];
                    if (p.z && rvalid) {
                        const long long zo = (long long)g * p.zq_gs + (long long)row * p.ldzq + nb;
                        float qv[32];
                        load_f32_row(p.z + zo, t, nv, p.vec);
                        load_f32_row(p.q + zo, qv, nv, p.vec);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaf(p.cscale, t[j] - qv[j], v[j]);
                    }
                    if (p.hb || p.hf) {
                        if (tin) read_in(c, t);
                        else if (p.hb) load_bf16_row(p.hb + (long long)g * p.hb_gs + (long long)(rvalid ? row : 0) * p.ldhb + nb, t, nv, p.vec);
                        else load_f32_row(p.hf + (long long)g * p.hf_gs + (long long)(rvalid ? row : 0) * p.ldhf + nb, t, nv, p.vec);
                        if (p.act == PGMVAE_ACT_SELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= pg_dselu_from_out(t[j]);
                        } else if (p.act == PGMVAE_ACT_SIGMOID) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= t[j] * (1.0f - t[j]);
                        }
                    }
                    if (p.cf && rvalid) store_f32_row(p.cf + (long long)g * p.cf_gs + (long long)row * p.ldcf + nb, v, nv, p.vec);
                    if (tout) write_out(c, v);
                    else if (p.cb && rvalid) store_bf16_row(p.cb + (long long)g * p.cb_gs + (long long)row * p.ldcb + nb, v, nv, p.vec);
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
                    if (p.accum) {
                        float t[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) t[j] = j < nv ? dst[(long long)j * p.lddw] : 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += t[j];
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < nv) dst[(long long)j * p.lddw] = j == zr ? 0.f : v[j];
                }
            };

            if (tin) fetch_in(0);
            if (EPI == EPI_SIGMOID_MSE) {                  // this row's target words of the tile, long before they are needed
                const uint32_t* yrow = p.ybits + (long long)(rvalid ? row : 0) * p.ldbits;
                const int w0 = n0 >> 5;
#pragma unroll
                for (int i = 0; i < 9; ++i) sY[i] = w0 + i < p.ldbits ? __ldg(yrow + w0 + i) : 0u;
            }
            tc::mbar_wait_parked(&tmem_full[wg], use & 1);      // parked: eight polling epilogue warps would take issue slots from the producer / MMA warps
            tc::fence_after_thread_sync();
            float va[32], vb[32];
            if (EPI == EPI_WGRAD_T && p.ones && nt == 0) {          // bias gradient of weight column `row` (uniform branch)
                tc::tmem_ld_32x32(lane_addr + ONES_COL - 16, va);
                tc::tmem_ld_wait(va);
                if (rvalid) {
                    float* d = p.db + (long long)g * p.db_gs + row;
                    *d = p.accum ? *d + va[16] : va[16];
                }
            }
            tc::tmem_ld_32x32(lane_addr, va);
            for (int c = 0; c < nch; c += 2) {
                tc::tmem_ld_wait(va);
                if (c + 1 < nch) {
                    tc::tmem_ld_32x32(lane_addr + (c + 1) * 32, vb);
                    if (tin) fetch_in(c + 1);              // buffer (c + 1) & 1 was last read before the barriers of chunk c - 1
                } else {                                   // the whole accumulator sits in registers: hand it back
                    tc::fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) release_acc();
                }
                process(va, c);
                if (c + 1 < nch) {
                    tc::tmem_ld_wait(vb);
                    if (c + 2 < nch) {
                        tc::tmem_ld_32x32(lane_addr + (c + 2) * 32, va);
                        if (tin) fetch_in(c + 2);
                    } else {
                        tc::fence_before_thread_sync();
                        __syncwarp();
                        if (lane == 0) release_acc();
                    }
                    process(vb, c + 1);
                }
            }
            if (EPI == EPI_SIGMOID_MSE) { dsq += (double)sq; dab += (double)ab; }
        }
        if (tout && leader) tc::bulk_wait_all();           // every store has landed before the CTA gives up its shared memory
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
    if (p.pair) tc::cluster_sync_all();                   // no CTA leaves while its peer may still signal its barriers
    if (warp == 2) {
        tc::fence_after_thread_sync();
        if (CTA2) tc::tmem_dealloc_2sm(tmem_base, 512u);
        else tc::tmem_dealloc(tmem_base, 512u);
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

// targets of the MSE stage, bit-packed: word w of row b holds columns 32 w .. 32 w + 31 (columns >= V read as 0)
template <class T>
__global__ void to_bits_kernel(const T* __restrict__ y, int ldy, uint32_t* __restrict__ bits, int ldbits, int B, int V) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * ldbits) return;
    const int b = (int)(i / ldbits), w = (int)(i - (long long)b * ldbits);
    uint32_t word = 0u;
    const T* row = y + (long long)b * ldy + 32 * w;
#pragma unroll 8
    for (int j = 0; j < 32; ++j)
        if (32 * w + j < V && row[j] != (T)0) word |= 1u << j;
    bits[i] = word;
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
    // 2-CTA clusters with the shared operand multicast: the main loop is bound by L2 -> SM traffic (ncu: lts 69 %,
    // tensor pipe 33 % with one CTA per tile), and a pair halves the traffic of the operand it shares.  Pair along M
    // (neighbouring M tiles share B) or along N (neighbouring N tiles share A), whichever moves fewer bytes per CTA and
    // k-block; an odd tile count costs a phantom tile, accepted up to 10 % of the work.
    p.tiles_mp = (int)pg_cdiv(p.tiles_m, 2);
    p.tiles_np = (int)pg_cdiv(p.tiles_n, 2);
    p.pair = 0;
    {
        const double none = (double)p.a_bytes + p.b_bytes;
        const double wm = 2.0 * p.tiles_mp / p.tiles_m, wn = 2.0 * p.tiles_np / p.tiles_n;
        const double cm = wm <= 1.10 ? ((double)p.a_bytes + 0.5 * p.b_bytes) * wm : 1e30;
        const double cn = wn <= 1.10 ? (0.5 * p.a_bytes + (double)p.b_bytes) * wn : 1e30;
        if (std::min(cm, cn) < 0.95 * none) p.pair = cm <= cn ? 1 : 2;
        if (p.total_tiles < 2 * ctx->sm_count) p.pair = 0;          // too little work to fill the machine with pairs anyway
        // pairs along M can go one step further: ONE cta_group::2 MMA over both SMs, each holding only ITS half of the
        // B tile -- the shared-memory fill per SM and k-block drops from A + B to A + B / 2 as well (what bounds the
        // multicast pair: every SM still receives the whole B tile)
        if (p.pair == 1 && getenv("PGMVAE_BF16_2SM") != nullptr) p.pair = 3;        // opt-in: measured equal to the multicast pair
        // clusters of 2 x 2 tiles share both operands (per CTA and k-block: A / 2 + B / 2 from L2) at the price of the SMs
        // that clusters of four cannot use (132 of 148 on B200): for the L2-bound main loops with long K
        {
            const double w4 = wm * wn;
            if (w4 <= 1.10 && p.kblocks >= 16 && p.total_tiles >= 4 * ctx->sm_count && getenv("PGMVAE_BF16_NO_QUAD") == nullptr &&
                0.5 * none * w4 * (148.0 / 132.0) < 0.9 * std::min(std::min(cm, cn), none))
                p.pair = 4;
        }
        if (const char* ev = getenv("PGMVAE_BF16_PAIR")) p.pair = atoi(ev) >= 0 && atoi(ev) <= 4 ? atoi(ev) : p.pair;
        if ((p.pair == 1 || p.pair == 3 || p.pair == 4) && (p.BN / 2) % 8) p.pair = 0;   // half tiles keep whole 8-row swizzle atoms
    }
    p.total_ptiles = (p.pair == 1 || p.pair == 3) ? p.G * p.tiles_mp * p.tiles_n
                     : (p.pair == 2 ? p.G * p.tiles_m * p.tiles_np
                                    : (p.pair == 4 ? p.G * p.tiles_mp * p.tiles_np : p.total_tiles));
    if (p.pair == 3) {                       // this CTA's half of the B tile
        p.b_panels = (int)pg_cdiv(p.BN / 2, 64);
        p.b_bytes = p.b_mn ? (unsigned)(p.b_panels * PANEL_BYTES) : (unsigned)pg_round_up((p.BN / 2) * 128, 1024);
    }
    // bf16 rows leave (and the epilogue's bf16 operand rows arrive) as staged tiles through TMA: a thread owns a row, so
    // direct 16-byte accesses touch 32 different lines per warp instruction and the epilogue is bound by LSU wavefronts
    EpiMaps em;
    memset(&em, 0, sizeof(em));
    const __nv_bfloat16* inp = EPI == EPI_DGRAD ? p.hb : nullptr;
    const int ld_in = p.ldhb;
    const int64_t gs_in = p.hb_gs;
    const bool epi_rows = EPI == EPI_FWD || EPI == EPI_SIGMOID_MSE || EPI == EPI_DGRAD;
    p.in_shared = 0;
    p.tma_out = epi_rows && p.cb && p.tiles_n <= MAX_NT_MAPS && al16(p.cb) && p.ldcb % 8 == 0 && p.cb_gs % 8 == 0 &&
                getenv("PGMVAE_BF16_DIRECT_EPI") == nullptr;
    p.tma_in = p.tma_out && inp && al16(inp) && ld_in % 8 == 0 && gs_in % 8 == 0;
    for (int nt = 0; nt < p.tiles_n && p.tma_out; ++nt) {
        const int n0 = nt * p.BN, ncols = std::min(p.BN, p.N - n0);
        PG_TRY(tc::make_map(&em.out[nt], p.cb + n0, 2, (uint64_t)ncols, (uint64_t)p.M, (uint64_t)p.G, (uint64_t)p.ldcb,
                            (uint64_t)p.cb_gs, 32, TM, false, true));
        if (p.tma_in)
            PG_TRY(tc::make_map(&em.in[nt], inp + n0, 2, (uint64_t)ncols, (uint64_t)p.M, (uint64_t)(p.in_shared ? 1 : p.G),
                                (uint64_t)ld_in, (uint64_t)gs_in, 32, TM, false, true));
    }
    p.stage_f32 = EPI == EPI_FWD && p.cf && !p.tma_out && p.vec && getenv("PGMVAE_BF16_DIRECT_EPI") == nullptr &&
                  getenv("PGMVAE_BF16_NO_F32_STAGE") == nullptr;
    const size_t stage = (size_t)A_BYTES + p.b_bytes;
    const bool staged = p.tma_out || p.stage_f32 || EPI == EPI_SIGMOID_MSE;       // (the MSE stage keeps its target words there)
    // short-K dgrad layers are bound by their epilogue, not by the depth of the operand ring: a second output tile
    p.out_db = EPI == EPI_DGRAD && p.tma_in && p.kblocks <= 8;
    const size_t fixed = 1024 + 256 + (staged ? 2 * (size_t)(p.out_db ? STG_WG4 : STG_WG) : 0) + (p.ones ? ONES_BYTES : 0) +
                         2 * ACC_COLS * sizeof(float);
    const bool slim = ctx->coresident;
    const size_t smem_cap = ctx->smem_optin - (slim ? 3072 : 0);    // (room for a co-resident exchange CTA, see the kernel)
    int stages = (int)((smem_cap - fixed) / stage);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (const char* ev = getenv("PGMVAE_BF16_STAGES")) stages = std::max(2, std::min(stages, atoi(ev)));   // (experiments)
    if (stages < 2) {
        pgmvae_set_error("%s: shared memory too small", name);
        return PGMVAE_EINVAL;
    }
    p.stages = stages;
    // always more than half an SM's shared memory: one CTA per SM, which owns all 512 TMEM columns
    size_t smem = fixed + stages * stage;
    if (smem < (size_t)120 * 1024) smem = (size_t)120 * 1024;
    CUtensorMap mA, mB;
    // (a K-major operand shared by a pair is loaded in two half-height boxes, one per CTA)
    const uint32_t a_box_rows = (p.pair == 2 || p.pair == 4) ? TM / 2 : TM;
    const uint32_t b_box_rows = (uint32_t)((p.pair == 1 || p.pair == 3 || p.pair == 4) ? p.BN / 2 : p.BN);
    if (!p.a_mn) PG_TRY(tc::make_map(&mA, A.p, 2, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)p.G, (uint64_t)A.ld, (uint64_t)A.gs, BK, a_box_rows));
    else PG_TRY(tc::make_map(&mA, A.p, 2, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.G, (uint64_t)A.ld, (uint64_t)A.gs, 64, BK));
    if (!p.b_mn) PG_TRY(tc::make_map(&mB, B.p, 2, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)p.G, (uint64_t)B.ld, (uint64_t)B.gs, BK, b_box_rows));
    else PG_TRY(tc::make_map(&mB, B.p, 2, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.G, (uint64_t)B.ld, (uint64_t)B.gs, 64, BK));
    static size_t configured[16][2][2] = {};
    const int dev = ctx->device & 15, two = p.pair == 3 ? 1 : 0;
    using Kern = void (*)(const CUtensorMap, const CUtensorMap, const EpiMaps, const Bf16P);
    const Kern kern = two ? (slim ? (Kern)gemm_bf16_kernel<EPI, true, true> : (Kern)gemm_bf16_kernel<EPI, true, false>)
                          : (slim ? (Kern)gemm_bf16_kernel<EPI, false, true> : (Kern)gemm_bf16_kernel<EPI, false, false>);
    if (smem > configured[dev][two][slim]) {
        PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev][two][slim] = smem;
    }
    const int csize = p.pair == 4 ? 4 : (p.pair ? 2 : 1);
    int nclusters = std::min(p.total_ptiles, ctx->sm_count / csize);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(nclusters * csize));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (csize == 4) {
        // tiles are assigned statically: never launch more clusters than can be resident at once (clusters of four do
        // not tile every GPC)
        static int max_quads[16][5][2] = {};
        int& mq = max_quads[dev][EPI][slim];
        if (mq == 0) {
            int n = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
            mq = (e == cudaSuccess && n > 0) ? n : 30;
        }
        nclusters = std::min(nclusters, mq);
        cfg.gridDim = dim3((unsigned)(nclusters * csize));
    }
    PG_KERNEL(ctx, st, name, bytes, 2.0 * p.G * (double)p.M * p.N * p.K);
    PG_CUDA(cudaLaunchKernelEx(&cfg, kern, mA, mB, em, p));
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

}  // namespace

// ---- internal entry points (ops.cuh) ---------------------------------------------------------------------------
int pg_bf16_fwd(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* x, int64_t x_gs, int ldx, const __nv_bfloat16* wt,
                int64_t wt_gs, int ldwt, const float* bias, int64_t bias_gs, __nv_bfloat16* outb, int64_t outb_gs, int ldob,
                float* outf, int64_t outf_gs, int ldof, int G, int B, int in, int out_dim, int act, int w_mn) {
    Bf16P p{};
    p.G = G; p.M = B; p.N = out_dim; p.K = in; p.b_mn = w_mn ? 1 : 0;
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
                            const uint32_t* ybits, int ldbits, __nv_bfloat16* dpre, int64_t dpre_gs, int ldd, float* out_opt,
                            int64_t out_gs, int ldo, double* acc2, int G, int g0, int B, int in, int V, float grad_scale,
                            int w_mn) {
    Bf16P p{};
    p.G = G; p.M = B; p.N = V; p.K = in; p.b_mn = w_mn ? 1 : 0;
    p.cb = dpre; p.cb_gs = dpre_gs; p.ldcb = ldd; p.cf = out_opt; p.cf_gs = out_gs; p.ldcf = ldo;
    p.bias = bias; p.bias_gs = bias_gs; p.ybits = ybits; p.ldbits = ldbits; p.acc = acc2; p.gscale = grad_scale; p.g0 = g0;
    p.vec = (!dpre || (al16(dpre) && ldd % 8 == 0 && dpre_gs % 8 == 0)) &&
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

// dW[in][out] (+)= X^T dY in the orientation that pads less, and db[out] (+)= column sums of dY (nullable): through
// the ones-tile MMA of the transposed orientation where that applies, with the column-sum kernel otherwise.
int pg_bf16_wgrad(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* x, int64_t x_gs, int ldx, const __nv_bfloat16* dy,
                  int64_t dy_gs, int lddy, float* dw, int64_t dw_gs, int lddw, float* db, int64_t db_gs, int G, int B, int in,
                  int out_dim, int zero_row_base, int accumulate) {
    auto padded = [](int m, int n) { return (double)pg_round_up(m, TM) * (double)(pg_cdiv(n, pick_bn(n)) * (pick_bn(n) + 48)); };
    // the transposed orientation gets the bias gradient for free (ones-tile MMA); the direct one pays a second pass over
    // dY, worth ~133 padded tile elements per output column at the measured rates of the two kernels
    bool direct = padded(in, out_dim) + (db ? 133.0 * out_dim : 0.0) < padded(out_dim, in);
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
        PG_TRY(launch<EPI_WGRAD_D>(ctx, st, p, Operand{x, x_gs, ldx}, Operand{dy, dy_gs, lddy}, "dense_wgrad_bf16", bytes));
        if (db) PG_TRY(pg_bf16_colsum(ctx, st, dy, dy_gs, lddy, db, db_gs, G, B, out_dim, accumulate));
        return PGMVAE_OK;
    }
    p.M = out_dim; p.N = in;
    p.ones = db != nullptr && pick_bn(in) <= ONES_COL - 16 + 16 && getenv("PGMVAE_WGRAD_NO_ONES") == nullptr;
    p.db = db; p.db_gs = db_gs;
    PG_TRY(launch<EPI_WGRAD_T>(ctx, st, p, Operand{dy, dy_gs, lddy}, Operand{x, x_gs, ldx}, "dense_wgrad_bf16", bytes));
    if (db && !p.ones) PG_TRY(pg_bf16_colsum(ctx, st, dy, dy_gs, lddy, db, db_gs, G, B, out_dim, accumulate));
    return PGMVAE_OK;
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

int pg_y_to_bits(pgmvae_ctx* ctx, cudaStream_t st, const uint8_t* y, int ldy, uint32_t* bits, int ldbits, int B, int V) {
    if (B <= 0) return PGMVAE_OK;
    const long long n = (long long)B * ldbits;
    PG_KERNEL(ctx, st, "y_to_bits", (double)B * V + 4.0 * n, 0.0);
    to_bits_kernel<uint8_t><<<(unsigned)pg_cdiv(n, 256), 256, 0, st>>>(y, ldy, bits, ldbits, B, V);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}
int pg_f32_to_bits(pgmvae_ctx* ctx, cudaStream_t st, const float* y, int ldy, uint32_t* bits, int ldbits, int B, int V) {
    if (B <= 0) return PGMVAE_OK;
    const long long n = (long long)B * ldbits;
    PG_KERNEL(ctx, st, "y_to_bits", 4.0 * B * V + 4.0 * n, 0.0);
    to_bits_kernel<float><<<(unsigned)pg_cdiv(n, 256), 256, 0, st>>>(y, ldy, bits, ldbits, B, V);
    PG_LAUNCHED(ctx);
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
