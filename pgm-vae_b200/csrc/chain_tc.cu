// Chain kernels: a whole stack of packed per-variable dense layers in ONE launch, the
// activations of a 128-sample tile never leaving the SM.
//
//   train               forward, then the dgrad stages of the same rows (19 stages): fd0..fd4 -> VQ (assignment,
//                       straight-through, commitment loss, EMA statistics) -> fd5..fd9 -> sigmoid + MSE/MAE +
//                       d(loss)/d(pre-activation) -> dgrad fd9..fd1 with the commitment gradient injected at the
//                       VQ boundary      reference core/model.py:39-55, core/quantizer.py:120-162, run.py:61-62
//   forward / backward  the two halves as separate launches (forward-only calls, PGMVAE_CHAIN_SPLIT=1)
//   encode              fd0..fd4 -> VQ assignment (-> PLL histogram)   core/model.py:48, :58-82
//
// Layout of one CTA (64 + 128 * nch threads, persistent over (variable, nch x 128-row tile) items, nch <= 3
// independent chains in flight):
//   warp 0   TMA producer: streams the weight k-blocks of every stage of every item through a
//            shared-memory ring; weights do not depend on activations, so it runs ahead of the
//            compute across stages AND items; a slot is released when every chain has consumed it
//   warp 1   owns the TMEM allocation (otherwise idle)
//   warps 2.. epilogue, four per chain, one sample row per thread: tcgen05.ld the pre-activations, bias +
//            activation (or the VQ / loss / act' step), store the row to HBM for the kernels that
//            need it later (wgrad), and tcgen05.st it back IN PLACE as the next A operand.  The chain's
//            first warp also ISSUES the chain's MMAs (tcgen05.mma kind::tf32, M = 128, A operand read from
//            TENSOR MEMORY, B = weights from the ring) once the four warps have met at the chain's named barrier.
//   Two TMEM regions ping-pong: stage j reads region (j & 1), writes region (~j & 1).
//
// What leaves the SM per layer is one write of the activations (256-bit stores: a thread owns a row); nothing is
// read back from HBM except the weights (L2-resident) -- against one read + one write per layer and ~47 launches
// per step for the layer-by-layer kernels in dense_tc.cu.  Chains are used when the network is narrow
// enough for TMEM (every padded width <= 256 and both regions within 512 columns), the codebook
// fits shared memory and D <= 32; otherwise the per-layer kernels run.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "ops.cuh"
#include "tc_common.cuh"

namespace {

constexpr int CH_TM = 128;
constexpr int CH_RING = 8;              // weight k-block stages: the producer runs ~4 layers ahead of the MMAs
constexpr int CH_MAXCH = 3;             // independent chains (128-row tiles of one variable) in flight per CTA
constexpr int CH_THREADS = 64 + 128 * CH_MAXCH;

struct ChainMaps { CUtensorMap m[PG_CHAIN_MAX_STAGES]; };

struct ChainP {
    int mode;                            // PG_CHAIN_*
    int nst;
    PgChainStage st[PG_CHAIN_MAX_STAGES];
    int G, g0, B, V, Vp, D, Dp, K, vq_stage;
    int tiles_m, tiles_real, tmem_cols, nbias, nch, sub_cols, tab_floats, tail_ok;
    uint32_t stage_bytes;
    // layer-0 operand / targets
    const float* a0; long long a0_gs; int lda0, a0_cols;    // fwd: yf [B][Vp] (shared); bwd: dpre of the top layer
    const float* yf; int ldyf;
    const uint8_t* y8; int ldy8;
    // VQ
    const float* E; long long e_gs;
    float* q; float* stq; long long zq_gs; int ldzq;
    int32_t* idx; long long idx_gs;
    float* stat_c; float* stat_w;
    double* acc;
    float gscale, cscale;
    unsigned long long* n1; unsigned long long* n0;
    const float* z; const float* qv;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }   // the four warps of one chain

// Asks for the 128-byte lines that hold [ptr, ptr + bytes) (bytes <= 384, a multiple of 32) to be brought into
// L2: the backward chain requests the rows it will read two stages ahead, so its register loads are L2 hits.
__device__ __forceinline__ void l2_prefetch(const float* ptr, int bytes) {
    const char* c = reinterpret_cast<const char*>(ptr);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c));
    if (bytes > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + 128));
    if (bytes > 256) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + 256));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + bytes - 32));     // the row's last sector may start a new line
}

// One thread moves (up to) 32 consecutive floats of its row with 256-bit accesses: every row starts on a 32-byte
// boundary (row strides are multiples of 8 floats), so a store fills whole 32-byte sectors -- half the memory
// instructions and half the L1 wavefronts of 128-bit accesses (the rows of a warp lie in 32 different lines).
__device__ __forceinline__ void store_chunk(float* dst, const float (&v)[32], int nv) {
#pragma unroll
    for (int j = 0; j < 32; j += 8)
        if (j < nv)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                         "f"(v[j + 2]), "f"(v[j + 3]), "f"(v[j + 4]), "f"(v[j + 5]), "f"(v[j + 6]), "f"(v[j + 7])
                         : "memory");
}
__device__ __forceinline__ void load_chunk(const float* src, float (&v)[32], int nv) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        if (j < nv) {
            asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=f"(v[j]), "=f"(v[j + 1]), "=f"(v[j + 2]), "=f"(v[j + 3]), "=f"(v[j + 4]), "=f"(v[j + 5]),
                           "=f"(v[j + 6]), "=f"(v[j + 7])
                         : "l"(src + j));
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j + i] = 0.f;
        }
    }
}

// Activations of the tensor-core chains: ex2.approx based (2 ulp); the pre-activations they act on
// already carry the 2^-11 relative error of the tf32 operands.  The exact-fp32 path (dense_simt.cu)
// keeps expf, and so does the EXACT instantiation (PGMVAE_CHAIN_EXACT=1), which reproduces the
// layer-by-layer tensor-core kernels bit for bit up to summation order.
__device__ __forceinline__ float ex2_ftz(float x) {       // bare MUFU.EX2: no denormal rescue around it
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool EXACT>
__device__ __forceinline__ float ch_selu(float x) {      // branch-free: 32 independent elements interleave
    if (EXACT) {
        const float e = PG_SELU_SCALE_ALPHA * (expf(fminf(x, 0.f)) - 1.0f);
        return x < 0.f ? e : PG_SELU_SCALE * x;
    }
    // scale * max(x, 0) + scale_alpha * (e^min(x, 0) - 1): the second term is exactly 0 for x >= 0
    const float neg = fmaf(PG_SELU_SCALE_ALPHA, ex2_ftz(fminf(x, 0.f) * 1.4426950408889634f), -PG_SELU_SCALE_ALPHA);
    return fmaf(PG_SELU_SCALE, fmaxf(x, 0.f), neg);
}
template <bool EXACT>
__device__ __forceinline__ float ch_sigmoid(float x) {
    return EXACT ? 1.0f / (1.0f + expf(-x)) : rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f));
}

// Exact fp32 arg-min of one row against the codebook in shared memory ([Kp][4*NV4] floats, Kp a
// multiple of 4, |e|^2 = +inf for the padding codes).  Four codes in flight (independent fmaf chains),
// each chain in the sequential order over d of the fp32 kernels: distances are bit-identical to
// vq_assign_kernel, lowest index on ties.
template <int NV4>
__device__ __forceinline__ int vq_row_argmin(const float (&v)[32], float zz, const float* __restrict__ sE,
                                             const float* __restrict__ sEE, int Kp) {
    float best = INFINITY;
    int bi = 0;
    for (int k0 = 0; k0 < Kp; k0 += 4) {
        float dot[4] = {0.f, 0.f, 0.f, 0.f};
        const float* er = sE + k0 * (4 * NV4);
#pragma unroll
        for (int d4 = 0; d4 < NV4; ++d4) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 e4 = *reinterpret_cast<const float4*>(er + kk * (4 * NV4) + d4 * 4);
                dot[kk] = fmaf(v[d4 * 4 + 0], e4.x, dot[kk]);
                dot[kk] = fmaf(v[d4 * 4 + 1], e4.y, dot[kk]);
                dot[kk] = fmaf(v[d4 * 4 + 2], e4.z, dot[kk]);
                dot[kk] = fmaf(v[d4 * 4 + 3], e4.w, dot[kk]);
            }
        }
        const float4 ee4 = *reinterpret_cast<const float4*>(sEE + k0);
        const float d0 = (zz - 2.0f * dot[0]) + ee4.x, d1 = (zz - 2.0f * dot[1]) + ee4.y;
        const float d2 = (zz - 2.0f * dot[2]) + ee4.z, d3 = (zz - 2.0f * dot[3]) + ee4.w;
        if (d0 < best) { best = d0; bi = k0; }
        if (d1 < best) { best = d1; bi = k0 + 1; }
        if (d2 < best) { best = d2; bi = k0 + 2; }
        if (d3 < best) { best = d3; bi = k0 + 3; }
    }
    return bi;
}

// The VQ step on one row held in registers (v[0 .. 4*NV4) = the padded latent; the rest of v is cleared):
// assignment, then either the PLL histogram (encode) or q = E[idx], the straight-through output st = z + (q - z)
// left in v, the commitment / codebook loss and the EMA statistics (core/quantizer.py:134-156).
template <int MODE, int NV4>
__device__ __forceinline__ void vq_stage(float (&v)[32], const ChainP& p, const float* __restrict__ sE,
                                         const float* __restrict__ sEE, uint32_t* sHist, int Kp, int g, int row, int rowc,
                                         bool valid, float& vq) {
    constexpr int DP = 4 * NV4;
#pragma unroll
    for (int d = DP; d < 32; ++d) v[d] = 0.f;
    float zz = 0.f;
#pragma unroll
    for (int d = 0; d < DP; ++d)
        if (d < p.D) zz = fmaf(v[d], v[d], zz);
    const int bi = vq_row_argmin<NV4>(v, zz, sE, sEE, Kp);
    if (valid && p.idx) p.idx[(long long)g * p.idx_gs + rowc] = bi;
    if (MODE == PG_CHAIN_ENCODE) {
        if (valid && p.n1) {
            const int bit = p.y8[(long long)row * p.ldy8 + p.g0 + g] != 0;
            atomicAdd(&sHist[(bit ? 0 : p.K) + bi], 1u);
        }
        return;
    }
    const float* er = sE + bi * DP;
    float qv[DP];
#pragma unroll
    for (int d = 0; d < DP; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(er + d);
        qv[d] = t.x; qv[d + 1] = t.y; qv[d + 2] = t.z; qv[d + 3] = t.w;
    }
    const long long zo = (long long)g * p.zq_gs + (long long)rowc * p.ldzq;
    if (valid) {
#pragma unroll
        for (int d = 0; d < DP; d += 8)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p.q + zo + d), "f"(qv[d]),
                         "f"(qv[d + 1]), "f"(qv[d + 2]), "f"(qv[d + 3]), "f"(qv[d + 4]), "f"(qv[d + 5]), "f"(qv[d + 6]),
                         "f"(qv[d + 7])
                         : "memory");
        if (p.stat_w) {
            float* dw = p.stat_w + ((long long)g * p.K + bi) * DP;
#pragma unroll
            for (int d = 0; d < DP; d += 4) red_add_v4(dw + d, v[d], v[d + 1], v[d + 2], v[d + 3]);
            atomicAdd(p.stat_c + (long long)g * p.K + bi, 1.0f);
        }
    }
#pragma unroll
    for (int d = 0; d < DP; ++d) {
        const float diff = qv[d] - v[d];
        if (valid && d < p.D) vq = fmaf(diff, diff, vq);
        v[d] = v[d] + diff;
    }
    if (valid) store_chunk(p.stq + zo, v, DP);
}

template <bool EXACT, int MODE>
__global__ void __launch_bounds__(CH_THREADS, 1)    // 14 warps are allocated as 16: 128 registers per thread
chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a shared-space pointer
    uint8_t* sB = smem;                                                          // [CH_RING][stage_bytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)CH_RING * p.stage_bytes);
    uint64_t* b_full = bars;                   // [CH_RING]
    uint64_t* b_empty = bars + CH_RING;        // [CH_RING]   one arrival per chain
    // (bars + 2 * CH_RING .. : CH_MAXCH spare slots)
    uint64_t* d_full = bars + 2 * CH_RING + CH_MAXCH;   // [CH_MAXCH]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CH_RING + 2 * CH_MAXCH);
    float* tables = reinterpret_cast<float*>(bars + 2 * CH_RING + 2 * CH_MAXCH + 2);     // per chain: sE | sEE | sHist | sBias
    const int Kp = (p.K + 3) & ~3;                                               // codes padded to a multiple of 4

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = p.G * p.tiles_m;         // item = (variable, group of nch consecutive 128-row tiles)
    // Consecutive items per CTA: mostly the same variable, so its codebook and biases stay in shared memory.
    // items = q * grid + rem: when the rem left-over items are few they are not handed out whole (which makes the
    // kernel as long as q + 1 items on a few CTAs while the others idle) but as single tiles, one per CTA, that run
    // in chain 0 alone at the end (a lone chain has the SM to itself and takes about half the time of a full item).
    int item_beg, item_end, tail_item = -1, tail_sel = 0;
    {
        const int grid = (int)gridDim.x, q = items / grid, rem = items - q * grid, b = (int)blockIdx.x;
        if (p.tail_ok && p.nch > 1 && q > 0 && rem > 0 && rem * p.nch <= grid) {
            item_beg = b * q;
            item_end = item_beg + q;
            if (b < rem * p.nch) {
                tail_item = q * grid + b / p.nch;
                tail_sel = b % p.nch;
                if ((tail_item % p.tiles_m) * p.nch + tail_sel >= p.tiles_real) tail_item = -1;     // padding tile
            }
        } else {
            item_beg = (int)((long long)b * items / grid);
            item_end = (int)((long long)(b + 1) * items / grid);
        }
    }
    const int n_main = item_end - item_beg, n_slots = n_main + (tail_item >= 0 ? 1 : 0);

    if (warp == 0 && lane == 0) {
        for (int j = 0; j < p.nst; ++j) tc::tma_prefetch_desc(&maps.m[j]);
        for (int s = 0; s < CH_RING; ++s) {
            tc::mbar_init(&b_full[s], 1);
            tc::mbar_init(&b_empty[s], p.nch);   // a weight k-block is released once every chain has consumed it
        }
        for (int c = 0; c < CH_MAXCH; ++c) {
            tc::mbar_init(&d_full[c], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        tc::tmem_relinquish();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: weight k-blocks of every stage of every item
        uint32_t s = 0, ph = 0;
        for (int sl = 0; sl < n_slots; ++sl) {
            const int item = sl < n_main ? item_beg + sl : tail_item;
            const int g = item / p.tiles_m;
            for (int j = 0; j < p.nst; ++j) {
                const PgChainStage& S = p.st[j];
                for (int kb = 0; kb < S.kblocks; ++kb) {
                    tc::mbar_wait(&b_empty[s], ph ^ 1);
                    if (tc::elect_one()) {
                        uint8_t* dst = sB + (size_t)s * p.stage_bytes;
                        tc::mbar_arrive_expect_tx(&b_full[s], S.kb_bytes);
                        if (S.b_mn) {
                            for (int pn = 0; pn < (S.N >> 5); ++pn)                  // panels [32 k][32 n]
                                tc::tma_load_3d(dst + pn * 4096, &maps.m[j], &b_full[s], pn * 32, kb * 32, g);
                        } else {
                            tc::tma_load_3d(dst, &maps.m[j], &b_full[s], kb * 32, 0, g);   // [N n][32 k]
                        }
                    }
                    __syncwarp();
                    if (++s == CH_RING) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp >= 2 && warp < 2 + 4 * p.nch) {      // (warp 1 only owns the TMEM allocation)
        // ===================== epilogue: one sample row per thread, four warps per chain
        const int q4 = warp & 3;                              // TMEM lane quarter == warp % 4
        const int ch = (warp - 2) >> 2;                       // chain of this warp
        const int r = q4 * 32 + lane;
        const int et = ((warp - 2) & 3) * 32 + lane;          // 0..127 among the threads of the chain
        const int bar_id = 1 + ch;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + ch * p.sub_cols;
        uint64_t* d_ful = &d_full[ch];
        float* sE = tables + (size_t)ch * p.tab_floats;       // [Kp][Dp]
        float* sEE = sE + (size_t)Kp * p.Dp;                  // [Kp]
        uint32_t* sHist = reinterpret_cast<uint32_t*>(sEE + Kp);   // [2][K]  (count mode)
        float* sBias = reinterpret_cast<float*>(sHist + 2 * Kp);  // biases of all stages, this variable
        uint32_t dph = 0;
        int cur_g = -1;
        // MMA issue.  There is no separate issuer warp: once the four warps of a chain have left the operand of a
        // stage in TMEM they meet at the chain's named barrier and the chain's first warp issues that stage's
        // tcgen05.mma itself (one elected lane) -- no polling warp that competes for issue slots, no mbarrier
        // round trip between the last tcgen05.st and the first MMA.  All chains consume the same k-block
        // sequence of the weight ring; a slot is released when every chain's MMAs have read it.
        // the issuers of the three chains sit on three different schedulers (1, 2, 3; scheduler 0 hosts the TMA
        // producer): with all of them on one, that scheduler's epilogue warps lagged and their siblings waited
        const bool issuer = q4 == 1 + ch;
        const uint64_t dB_mn = tc::make_smem_desc(tc::smem_u32(sB), 4096, 512, 1);
        const uint64_t dB_k = tc::make_smem_desc(tc::smem_u32(sB), 16, 1024, 2);
        const uint32_t stage_stride = p.stage_bytes >> 4;
        uint32_t wseq = 0;                                    // k-blocks consumed by this chain so far
        auto issue_stage = [&](int j) {
            tc::tmem_st_wait();
            tc::fence_before_thread_sync();
            epi_bar(bar_id);                                  // the operand rows of all 128 threads are in TMEM
            if (!issuer) return;
            tc::fence_after_thread_sync();
            const PgChainStage& S = p.st[j];
            const uint32_t idesc = tc::make_idesc(2, CH_TM, S.N, 0, S.b_mn);
            const uint32_t bstep = (S.b_mn ? 1024u : 32u) >> 4;
            const uint32_t tb = tmem_base + ch * p.sub_cols;
            for (int kb = 0; kb < S.kblocks; ++kb, ++wseq) {
                const uint32_t slot = wseq % CH_RING, gen = wseq / CH_RING;
                tc::mbar_wait(&b_full[slot], gen & 1);
                tc::fence_after_thread_sync();
                if (tc::elect_one()) {
                    const uint64_t dB = (S.b_mn ? dB_mn : dB_k) + (uint64_t)(slot * stage_stride);
                    const int nk = min(4, S.ksteps - kb * 4);
                    for (int k4 = 0; k4 < nk; ++k4)
                        tc::mma_tf32_ts(tb + S.d_col, tb + S.a_col + kb * 32 + k4 * 8, dB + (uint64_t)(k4 * bstep), idesc,
                                        (kb | k4) ? 1u : 0u);
                    tc::mma_commit(&b_empty[slot]);
                    if (kb == S.kblocks - 1) tc::mma_commit(d_ful);
                }
                __syncwarp();
            }
        };
        double acc_sq = 0.0, acc_ab = 0.0, acc_vq = 0.0;
        for (int sl = 0; sl < n_slots; ++sl) {
            const bool tail = sl >= n_main;
            const int item = tail ? tail_item : item_beg + sl;
            if (tail && ch != 0) {
                // this chain sits the single-tile item out; its issuer still hands every weight k-block back
                if (issuer) {
                    for (int j = 0; j < p.nst; ++j)
                        for (int kb = 0; kb < p.st[j].kblocks; ++kb, ++wseq) {
                            const uint32_t slot = wseq % CH_RING, gen = wseq / CH_RING;
                            tc::mbar_wait(&b_full[slot], gen & 1);
                            if (lane == 0) tc::mbar_arrive(&b_empty[slot]);
                            __syncwarp();
                        }
                }
                continue;
            }
            const int g = item / p.tiles_m, mt = item - g * p.tiles_m;
            const int row = (mt * p.nch + (tail ? tail_sel : ch)) * CH_TM + r;
            const bool valid = row < p.B;
            const int rowc = valid ? row : 0;
            // ---- codebook of this variable -> shared memory, |e|^2 in the order every fp32 path uses
            if (g != cur_g) {
                epi_bar(bar_id);                                    // everyone has left the previous variable's tables
                for (int j = 0; j < p.nst; ++j) {
                    const PgChainStage& S = p.st[j];
                    if (S.bias)
                        for (int i = et; i < S.pout; i += 128) sBias[S.bias_off + i] = S.bias[(long long)g * S.bias_gs + i];
                }
                if (p.vq_stage >= 0) {
                    const float* eg = p.E + (long long)g * p.e_gs;
                    for (int i = et * 4; i < p.K * p.Dp; i += 128 * 4)
                        *reinterpret_cast<float4*>(sE + i) = *reinterpret_cast<const float4*>(eg + i);
                    for (int i = p.K * p.Dp + et; i < Kp * p.Dp; i += 128) sE[i] = 0.f;
                    epi_bar(bar_id);
                    for (int k = et; k < Kp; k += 128) {
                        float s2 = 0.f;
                        for (int d = 0; d < p.D; ++d) s2 = fmaf(sE[k * p.Dp + d], sE[k * p.Dp + d], s2);
                        sEE[k] = k < p.K ? s2 : INFINITY;
                        if (MODE == PG_CHAIN_ENCODE && p.n1 && k < p.K) { sHist[k] = 0; sHist[p.K + k] = 0; }
                    }
                }
                epi_bar(bar_id);
                cur_g = g;
            }
            // ---- operand of the first stage: this thread's row -> TMEM
            {
                const float* ar = p.a0 + (long long)g * p.a0_gs + (long long)rowc * p.lda0;
                const uint32_t a_addr = lane_addr + p.st[0].a_col;
                for (int c = 0; c < p.a0_cols; c += 64) {     // sixteen 16-byte loads in flight per thread
                    float t0[32], t1[32];
                    load_chunk(ar + c, t0, valid ? min(32, p.a0_cols - c) : 0);
                    load_chunk(ar + c + 32, t1, valid ? min(32, p.a0_cols - c - 32) : 0);
                    tc::tmem_st_32x32(a_addr + c, t0);        // the region is a multiple of 32 columns wide
                    if (c + 32 < p.a0_cols) tc::tmem_st_32x32(a_addr + c + 32, t1);
                }
                issue_stage(0);
            }
            float sq = 0.f, ab = 0.f, vq = 0.f;
            auto prefetch_stage = [&](int jn) {        // backward: the rows stage jn reads from HBM -> L2
                if (MODE != PG_CHAIN_BWD || jn >= p.nst || !valid) return;
                const PgChainStage& N = p.st[jn];
                l2_prefetch(N.aux + (long long)g * N.aux_gs + (long long)rowc * N.ldaux, N.pout * 4);
                if (N.add_commit) {
                    const long long zo = (long long)g * p.zq_gs + (long long)rowc * p.ldzq;
                    l2_prefetch(p.z + zo, N.pout * 4);
                    l2_prefetch(p.qv + zo, N.pout * 4);
                }
            };
            prefetch_stage(0);
            prefetch_stage(1);
            for (int j = 0; j < p.nst; ++j) {
                const PgChainStage& S = p.st[j];
                prefetch_stage(j + 2);
                // operands that come from HBM are requested before waiting for the MMAs of this stage, and the next
                // chunk's before the current chunk is processed
                float hv[32];
                const float* hsrc = nullptr;
                // TRAIN = the forward stages followed by the dgrad stages of the same rows in one pass
                const bool dgrad = MODE == PG_CHAIN_BWD || (MODE == PG_CHAIN_TRAIN && S.kind == PG_CHAIN_EPI_DGRAD);
                constexpr bool HAS_LOSS = MODE == PG_CHAIN_FWD || MODE == PG_CHAIN_TRAIN;
                if (dgrad) hsrc = S.aux + (long long)g * S.aux_gs + (long long)rowc * S.ldaux;
                else if (HAS_LOSS && S.kind == PG_CHAIN_EPI_SIGMOID_MSE) hsrc = p.yf + (long long)rowc * p.ldyf;
                if (hsrc) load_chunk(hsrc, hv, min(32, S.pout));
                if (dgrad && S.add_commit) {          // commitment gradient (single chunk: Dp <= 32)
                    float zv[32], qv[32];
                    const long long zo = (long long)g * p.zq_gs + (long long)rowc * p.ldzq;
                    load_chunk(p.z + zo, zv, min(32, S.pout));
                    load_chunk(p.qv + zo, qv, min(32, S.pout));
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        hv[jj] = pg_dselu_from_out(hv[jj]);           // hv = 0 beyond pout
                        zv[jj] -= qv[jj];
                    }
                    tc::mbar_wait(d_ful, dph);
                    dph ^= 1;
                    tc::fence_after_thread_sync();
                    float v[32];
                    tc::tmem_ld_32x32(lane_addr + S.d_col, v);
                    tc::tmem_ld_wait(v);
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) v[jj] = fmaf(p.cscale, zv[jj], v[jj]) * hv[jj];
                    if (valid && S.outp)
                        store_chunk(S.outp + (long long)g * S.out_gs + (long long)rowc * S.ldo, v, min(32, S.pout));
                    if (j + 1 != p.nst) {
                        tc::tmem_st_32x32(lane_addr + S.d_col, v);
                        issue_stage(j + 1);
                    }
                    continue;
                }
                if (dgrad) {     // dX = (dY W^T) * act'(h): two chunks per round, both requested up front
                    float h1[32];
                    load_chunk(hsrc + 32, h1, min(32, S.pout - 32));
                    tc::mbar_wait(d_ful, dph);
                    dph ^= 1;
                    tc::fence_after_thread_sync();
                    float* orow = S.outp ? S.outp + (long long)g * S.out_gs + (long long)rowc * S.ldo : nullptr;
                    const bool wr = valid && orow != nullptr, fwd_st = j + 1 != p.nst;
                    for (int c = 0; c < S.pout; c += 64) {
                        if (c > 0) {
                            load_chunk(hsrc + c, hv, min(32, S.pout - c));
                            load_chunk(hsrc + c + 32, h1, min(32, S.pout - c - 32));
                        }
                        float v[32];
                        tc::tmem_ld_32x32(lane_addr + S.d_col + c, v);
                        tc::tmem_ld_wait(v);
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) v[jj] *= pg_dselu_from_out(hv[jj]);   // hv = 0 beyond pout
                        if (wr) store_chunk(orow + c, v, min(32, S.pout - c));
                        if (fwd_st) tc::tmem_st_32x32(lane_addr + S.d_col + c, v);
                        if (c + 32 < S.pout) {
                            tc::tmem_ld_32x32(lane_addr + S.d_col + c + 32, v);
                            tc::tmem_ld_wait(v);
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj) v[jj] *= pg_dselu_from_out(h1[jj]);
                            if (wr) store_chunk(orow + c + 32, v, min(32, S.pout - c - 32));
                            if (fwd_st) tc::tmem_st_32x32(lane_addr + S.d_col + c + 32, v);
                        }
                    }
                    if (fwd_st) issue_stage(j + 1);
                    continue;
                }
                tc::mbar_wait(d_ful, dph);
                dph ^= 1;
                tc::fence_after_thread_sync();
                const bool last = j + 1 == p.nst;
                const float* bias = sBias + S.bias_off;       // this variable's biases (shared memory)
                for (int c = 0; c < S.pout; c += 32) {
                    const int nv = min(32, S.pout - c);
                    float v[32];
                    tc::tmem_ld_32x32(lane_addr + S.d_col + c, v);
                    tc::tmem_ld_wait(v);
                    float* orow = S.outp ? S.outp + (long long)g * S.out_gs + (long long)rowc * S.ldo + c : nullptr;
                    if (MODE != PG_CHAIN_BWD && S.kind == PG_CHAIN_EPI_SELU) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj)
                            v[jj] = ch_selu<EXACT>(v[jj] + bias[c + jj]);        // columns >= pout are never consumed
                        if (valid && orow) store_chunk(orow, v, nv);
                        if (j == p.vq_stage) {
                            switch (p.Dp >> 3) {      // padded latent width: a multiple of 8, at most 32
                                case 1: vq_stage<MODE, 2>(v, p, sE, sEE, sHist, Kp, g, row, rowc, valid, vq); break;
                                case 2: vq_stage<MODE, 4>(v, p, sE, sEE, sHist, Kp, g, row, rowc, valid, vq); break;
                                case 3: vq_stage<MODE, 6>(v, p, sE, sEE, sHist, Kp, g, row, rowc, valid, vq); break;
                                default: vq_stage<MODE, 8>(v, p, sE, sEE, sHist, Kp, g, row, rowc, valid, vq); break;
                            }
                        }
                    } else if (HAS_LOSS && S.kind == PG_CHAIN_EPI_SIGMOID_MSE) {
                        if (c > 0) load_chunk(hsrc + c, hv, nv);
                        // columns that count: inside the data, not the net's own variable, row inside the batch
                        const int self = p.g0 + g - c;
                        uint32_t on = nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u);
                        if (p.V - c < 32) on &= p.V - c > 0 ? ((1u << (p.V - c)) - 1u) : 0u;
                        if (self >= 0 && self < 32) on &= ~(1u << self);
                        if (!valid) on = 0u;
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const float o = ch_sigmoid<EXACT>(v[jj] + bias[c + jj]);
                            const float d = (on & (1u << jj)) ? o - hv[jj] : 0.f;
                            sq = fmaf(d, d, sq);
                            ab += fabsf(d);
                            v[jj] = (p.gscale * d) * fmaf(-o, o, o);
                        }
                        if (valid && orow) store_chunk(orow, v, nv);
                    }
                    if (!last) tc::tmem_st_32x32(lane_addr + S.d_col + c, v);
                }
                if (!last) issue_stage(j + 1);
            }
            acc_sq += (double)sq; acc_ab += (double)ab; acc_vq += (double)vq;
            // ---- PLL histogram of this item -> global counters
            if (MODE == PG_CHAIN_ENCODE && p.n1) {
                epi_bar(bar_id);
                for (int k = et; k < 2 * p.K; k += 128) {
                    const uint32_t c = sHist[k];
                    if (c) {
                        sHist[k] = 0;
                        unsigned long long* dst = k < p.K ? p.n1 + (long long)g * p.K + k
                                                          : p.n0 + (long long)g * p.K + (k - p.K);
                        atomicAdd(dst, (unsigned long long)c);
                    }
                }
                epi_bar(bar_id);
            }
        }
        if (p.acc && (MODE == PG_CHAIN_FWD || MODE == PG_CHAIN_TRAIN)) {
            acc_sq = pg_warp_sum_d(acc_sq); acc_ab = pg_warp_sum_d(acc_ab); acc_vq = pg_warp_sum_d(acc_vq);
            if (lane == 0) {
                atomicAdd(p.acc + 0, acc_sq);
                atomicAdd(p.acc + 1, acc_ab);
                atomicAdd(p.acc + 2, acc_vq);
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 1) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

}  // namespace

// Builds the launch from the stage table prepared by model.cu (which owns the layer geometry).
int pg_chain_launch(pgmvae_ctx* ctx, cudaStream_t st, const PgChainArgs& a) {
    if (a.G <= 0 || a.B <= 0) return PGMVAE_OK;
    ChainP p{};
    ChainMaps maps{};
    p.mode = a.mode; p.nst = a.nst;
    p.G = a.G; p.g0 = a.g0; p.B = a.B; p.V = a.V; p.Vp = a.Vp; p.D = a.D; p.Dp = a.Dp; p.K = a.K;
    p.vq_stage = a.vq_stage;
    if (a.vq_stage < 0) p.K = 0;          // no codebook tables in shared memory
    int regw[2] = {a.a0_cols, 0};
    uint32_t stage_bytes = 0;
    double flops = 0.0, bytes = 0.0;
    for (int j = 0; j < a.nst; ++j) {
        PgChainStage S = a.st[j];
        S.N = pg_round_up(S.pout, 32);
        S.ksteps = S.K / 8;
        S.kblocks = (int)pg_cdiv(S.K, 32);
        S.kb_bytes = S.b_mn ? (uint32_t)(S.N / 32) * 4096u : (uint32_t)S.N * 128u;
        stage_bytes = std::max(stage_bytes, S.kb_bytes);
        regw[(j + 1) & 1] = std::max(regw[(j + 1) & 1], S.N);
        regw[j & 1] = std::max(regw[j & 1], S.K);
        S.bias_off = p.nbias;
        if (S.bias) p.nbias += S.pout;
        p.st[j] = S;
        if (S.b_mn)    // W[K = in][N = out], N contiguous: boxes [32 k][32 n], 32-byte-atom swizzle
            PG_TRY(tc::make_map(&maps.m[j], S.w, 4, (uint64_t)S.n_valid, (uint64_t)S.k_valid, (uint64_t)a.G,
                                (uint64_t)S.ldw, (uint64_t)S.w_gs, 32, 32, true));
        else           // W[N = in][K = out], K contiguous: boxes [N n][32 k]
            PG_TRY(tc::make_map(&maps.m[j], S.w, 4, (uint64_t)S.k_valid, (uint64_t)S.n_valid, (uint64_t)a.G,
                                (uint64_t)S.ldw, (uint64_t)S.w_gs, 32, (uint32_t)S.N));
        flops += 2.0 * a.G * (double)a.B * S.k_valid * S.n_valid;
        bytes += 4.0 * a.G * ((double)S.k_valid * S.n_valid + (S.outp ? (double)a.B * S.n_valid : 0.0) +
                              (S.aux && a.mode == PG_CHAIN_BWD ? (double)a.B * S.n_valid : 0.0));
    }
    const int w0 = pg_round_up(regw[0], 32), w1 = pg_round_up(regw[1], 32);
    for (int j = 0; j < a.nst; ++j) {
        p.st[j].a_col = (j & 1) ? w0 : 0;
        p.st[j].d_col = (j & 1) ? 0 : w0;
    }
    p.sub_cols = w0 + w1;
    p.tiles_real = (int)pg_cdiv(a.B, CH_TM);
    p.nch = std::max(1, std::min({CH_MAXCH, 512 / std::max(1, p.sub_cols), p.tiles_real}));
    if (const char* ev = getenv("PGMVAE_CHAINS")) p.nch = std::max(1, std::min(p.nch, atoi(ev)));
    p.tiles_m = (int)pg_cdiv(p.tiles_real, p.nch);
    p.tail_ok = getenv("PGMVAE_CHAIN_NO_TAIL") == nullptr;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.nch * p.sub_cols) p.tmem_cols <<= 1;
    p.stage_bytes = stage_bytes;
    p.a0 = a.a0; p.a0_gs = a.a0_gs; p.lda0 = a.lda0; p.a0_cols = a.a0_cols;
    p.yf = a.yf; p.ldyf = a.ldyf; p.y8 = a.y8; p.ldy8 = a.ldy8;
    p.E = a.E; p.e_gs = a.e_gs; p.q = a.q; p.stq = a.stq; p.zq_gs = a.zq_gs; p.ldzq = a.ldzq;
    p.idx = a.idx; p.idx_gs = a.idx_gs; p.stat_c = a.stat_c; p.stat_w = a.stat_w; p.acc = a.acc;
    p.gscale = a.gscale; p.cscale = a.cscale; p.n1 = a.n1; p.n0 = a.n0; p.z = a.z; p.qv = a.qv;
    bytes += 4.0 * (a.a0_gs ? (double)a.G : 1.0) * a.B * a.a0_cols;
    if (a.vq_stage >= 0) { flops += 2.0 * a.G * (double)a.B * a.D * a.K; bytes += 4.0 * a.G * (double)a.K * a.D; }

    {   // rows move with 256-bit accesses: 32-byte aligned bases, row strides multiples of 8 floats
        auto ok = [](const void* q, long long gs, int ld) { return !q || (!((uintptr_t)q & 31) && gs % 8 == 0 && ld % 8 == 0); };
        bool al = ok(a.a0, a.a0_gs, a.lda0) && ok(a.yf, 0, a.ldyf) && ok(a.q, a.zq_gs, a.ldzq) && ok(a.stq, a.zq_gs, a.ldzq) &&
                  ok(a.z, a.zq_gs, a.ldzq) && ok(a.qv, a.zq_gs, a.ldzq);
        for (int j = 0; j < a.nst; ++j)
            al = al && ok(a.st[j].outp, a.st[j].out_gs, a.st[j].ldo) && ok(a.st[j].aux, a.st[j].aux_gs, a.st[j].ldaux);
        if (!al) {
            pgmvae_set_error("chain kernel: activation rows must be 32-byte aligned");
            return PGMVAE_EINVAL;
        }
    }
    const size_t Kp = (size_t)((a.K + 3) & ~3);
    p.tab_floats = (int)(((a.vq_stage >= 0 ? Kp * a.Dp + 3 * Kp : 0) + p.nbias + 40 + 3) & ~(size_t)3);
    const size_t smem = 1024 + (size_t)CH_RING * stage_bytes + (size_t)p.nch * p.tab_floats * 4 + 512;
    if (smem > ctx->smem_optin || w0 + w1 > 512) {
        pgmvae_set_error("chain kernel: configuration does not fit (smem %zu, tmem columns %d)", smem, w0 + w1);
        return PGMVAE_EINVAL;
    }
    const bool exact = getenv("PGMVAE_CHAIN_EXACT") != nullptr;
    using KernelT = void (*)(const ChainMaps, const ChainP);
    static const KernelT kernels[2][4] = {
        {chain_kernel<false, PG_CHAIN_FWD>, chain_kernel<false, PG_CHAIN_ENCODE>, chain_kernel<false, PG_CHAIN_BWD>,
         chain_kernel<false, PG_CHAIN_TRAIN>},
        {chain_kernel<true, PG_CHAIN_FWD>, chain_kernel<true, PG_CHAIN_ENCODE>, chain_kernel<true, PG_CHAIN_BWD>,
         chain_kernel<true, PG_CHAIN_TRAIN>}};
    if (a.mode < 0 || a.mode > 3) return PGMVAE_EINVAL;
    const KernelT kern = kernels[exact][a.mode];
    static size_t configured[16][2][4] = {};          // per device: the attribute is set per device
    if (smem > configured[ctx->device & 15][exact][a.mode]) {
        PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15][exact][a.mode] = smem;
    }
    const int items = a.G * p.tiles_m;
    const int grid = std::min(items, ctx->sm_count);        // one CTA per SM, nch chains in flight each
    static const char* const names[4] = {"chain_fwd_tc", "chain_encode_tc", "chain_bwd_tc", "chain_train_tc"};
    PG_KERNEL(ctx, st, names[a.mode], bytes, flops);
    const int threads = 64 + 128 * p.nch;
    kern<<<grid, threads, smem, st>>>(maps, p);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

// Can this network run as chains?  (model.cu asks once at model creation)
bool pg_chain_supported(const int* pin, const int* pout, int nlayers, int Vp, int Dp, int K, size_t smem_optin) {
    int reg[2] = {Vp, 0};
    size_t stage = 0;
    for (int l = 0; l < nlayers; ++l) {
        if (pout[l] > 256 || pin[l] > 256) return false;
        const int N = pg_round_up(pout[l], 32);
        reg[(l + 1) & 1] = std::max(reg[(l + 1) & 1], N);
        reg[l & 1] = std::max(reg[l & 1], pin[l]);
        stage = std::max(stage, (size_t)(N / 32) * 4096);
        stage = std::max(stage, (size_t)pg_round_up(pin[l], 32) * 128);
    }
    if (pg_round_up(reg[0], 32) + pg_round_up(reg[1], 32) > 512) return false;
    if (Dp > 32) return false;
    size_t nbias = 0;
    for (int l = 0; l < nlayers; ++l) nbias += pout[l];
    const size_t smem = 1024 + CH_RING * stage + CH_MAXCH * ((size_t)(K + 3) * Dp + 3 * (K + 3) + nbias + 44) * 4 + 512;
    return smem <= smem_optin;
}
