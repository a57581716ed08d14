// Kernel (a), tensor-core flavour: the packed per-variable dense layer as a grouped GEMM on
// tcgen05 (kind::tf32 directly on the fp32 tensors as they lie in HBM, fp32 accumulation in
// TMEM), operands staged by TMA with the 128-byte swizzle, fused epilogues.  Replaces
// tf.matmul + bias + activation of FatDense.call (reference core/dense.py:99-111) and the
// gradient GEMMs of its autodiff (run.py:62).
//
//   forward : C[B,out] = A[B,in]   (K-major)  x  W[in,out]  (N contiguous -> MN-major B)
//   dgrad   : C[B,in]  = dY[B,out] (K-major)  x  W[in,out] read as [n=in][k=out] (K-major B)
//   wgrad   : C[in,out]= X[B,in] read as [k=b][m=in] (MN-major A) x dY[B,out] (MN-major B),
//             reduction over the batch split across CTAs, fp32 vector reductions into dW;
//             wgrad_multi_kernel runs the wgrad CTAs of ALL layers of a step as one launch
//
// One CTA computes one 128 x BN output tile of one variable: warp 0 = TMA producer over the
// k-blocks (32 fp32 = one swizzle row per k-block, `stages`-deep ring), warp 1 = MMA issuer,
// then all four warps run the epilogue (tcgen05.ld 32x32b: one output row per thread).
// Tiles are small and many, so latency is hidden by 2-3 resident CTAs per SM rather than by a
// persistent tile loop.  Out-of-bounds rows / columns / k are zero-filled by TMA, so ragged
// batches and the awkward layer widths (15, 14, 13, 50, 1555 ...) need no host-side padding
// beyond 16-byte row strides.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "ops.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;
constexpr int KBLK = 32;                          // fp32 elements per k-block (128 bytes)
constexpr int A_STAGE_BYTES = TM * 128;           // 16 KB
constexpr int MAX_STAGES = 8;

enum { EPI_BIAS_ACT = 0, EPI_SIGMOID_MSE = 1, EPI_DGRAD = 2, EPI_WGRAD = 3 };

struct DenseTcP {
    int M, N, K;                     // per-variable problem: C[M,N] = A[M,K] * B[K,N]
    int BN, kblocks, stages, tmem_cols;
    int a_stage_bytes, apan;         // bytes of one A stage; MN-major A: 32-row panels actually loaded (<= 4)
    int a_mn, b_mn;                  // operand is MN-major (panels of 128 bytes along M / N)
    int a_shared;                    // A is one matrix shared by all variables (the raw data y)
    int vecC;                        // rows of C / aux are 16-byte aligned
    int S, kb_per_split;             // split of the k-blocks over CTAs (wgrad)
    float* C; long long c_gs; int ldc;
    const float* bias; long long bias_gs; int act;
    const float* aux; long long aux_gs; int ldaux;          // SIGMOID_MSE: y (shared); DGRAD: h_in
    float* C2; double* acc; float gscale; int g0;
    const float* z; const float* q; long long zq_gs; int ldzq; float cscale;
    int zero_row_base;
    float* db; long long db_gs;      // wgrad: bias gradient (column sums of dY) from a ones-row MMA
};

// one thread moves (up to) 32 consecutive floats of its row; 128-bit accesses when the row is aligned
__device__ __forceinline__ void store_row(float* dst, const float (&v)[32], int nv, int vec) {
    if (vec && nv == 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nv) dst[j] = v[j];
    }
}
__device__ __forceinline__ void load_row(const float* src, float (&v)[32], int nv, int vec) {
    if (vec && nv == 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 t = *reinterpret_cast<const float4*>(src + j);
            v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = j < nv ? src[j] : 0.f;
    }
}

// The CTA program: (bx, by, bz) = (N tile, M tile, variable x batch split) of the problem p.
template <int EPI>
__device__ __forceinline__ void dense_tc_body(const CUtensorMap& mapA, const CUtensorMap& mapB, const DenseTcP& p,
                                              const int bx, const int by, const int bz) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a shared-space pointer
    const int b_stage_bytes = p.BN * 128;
    // [stages][a_stage_bytes]: an MN-major A tile only holds the panels that contain valid rows; the MMA
    // (M = 128) reads on into the next stage, which only feeds accumulator rows nobody looks at
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)p.stages * p.a_stage_bytes + (A_STAGE_BYTES - p.a_stage_bytes);   // [stages][BN*128]
    // wgrad: constant A tile whose row 0 is all ones (K-major, one 128-byte row): a second MMA per k-step
    // then leaves the column sums of dY (the bias gradient) in row 0 of a second accumulator
    uint8_t* sOnes = sB + (size_t)p.stages * b_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + (EPI == EPI_WGRAD ? A_STAGE_BYTES : 0));
    uint64_t* full = bars;                       // [MAX_STAGES]
    uint64_t* empty = bars + MAX_STAGES;         // [MAX_STAGES]
    uint64_t* acc_full = bars + 2 * MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);
    __shared__ double red[2][4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = bz / p.S, split = bz - g * p.S;
    const int m0 = by * TM, n0 = bx * p.BN;
    const int kb_beg = split * p.kb_per_split;
    const int kb_end = min(p.kblocks, kb_beg + p.kb_per_split);
    const int nkb = kb_end - kb_beg;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&mapA);
        tc::tma_prefetch_desc(&mapB);
        for (int s = 0; s < MAX_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (EPI == EPI_WGRAD && p.db) {
        float4* o4 = reinterpret_cast<float4*>(sOnes);
        for (int i = threadIdx.x; i < A_STAGE_BYTES / 16; i += 128)
            o4[i] = i < 8 ? make_float4(1.f, 1.f, 1.f, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // visible to the tensor core
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        tc::tmem_relinquish();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (nkb > 0) {
        if (warp == 0) {
            // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
            const uint32_t stage_tx = (uint32_t)(p.a_stage_bytes + b_stage_bytes);
            uint32_t s = 0, ph = 0;
            const int ga = p.a_shared ? 0 : g;
            for (int kb = kb_beg; kb < kb_end; ++kb) {
                tc::mbar_wait(&empty[s], ph ^ 1);
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&full[s], stage_tx);
                    uint8_t* a_dst = sA + (size_t)s * p.a_stage_bytes;
                    uint8_t* b_dst = sB + (size_t)s * b_stage_bytes;
                    if (!p.a_mn) {
                        tc::tma_load_3d(a_dst, &mapA, &full[s], kb * KBLK, m0, ga);                 // [128 m][32 k]
                    } else {
                        for (int pn = 0; pn < p.apan; ++pn)                                         // panels [32 k][32 m]
                            tc::tma_load_3d(a_dst + pn * 4096, &mapA, &full[s], m0 + pn * 32, kb * KBLK, ga);
                    }
                    if (!p.b_mn) {
                        tc::tma_load_3d(b_dst, &mapB, &full[s], kb * KBLK, n0, g);                 // [BN n][32 k]
                    } else {
                        for (int pn = 0; pn < p.BN / 32; ++pn)                                      // panels [32 k][32 n]
                            tc::tma_load_3d(b_dst + pn * 4096, &mapB, &full[s], n0 + pn * 32, kb * KBLK, g);
                    }
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
            const uint32_t idesc = tc::make_idesc(2, TM, p.BN, p.a_mn, p.b_mn);
            // K-major: rows of 128 B along k, 8-row atoms 1024 B apart; one MMA (k = 8) advances 32 B.
            // MN-major (32-byte-atom swizzle): panels of [32 k][128 B along m/n], 4096 B apart (LBO),
            //           4-k atoms 512 B apart (SBO); one MMA (k = 8) advances two atoms = 1024 B.
            const uint64_t dA0 = p.a_mn ? tc::make_smem_desc(tc::smem_u32(sA), 4096, 512, 1)
                                        : tc::make_smem_desc(tc::smem_u32(sA), 16, 1024, 2);
            const uint64_t dB0 = p.b_mn ? tc::make_smem_desc(tc::smem_u32(sB), 4096, 512, 1)
                                        : tc::make_smem_desc(tc::smem_u32(sB), 16, 1024, 2);
            const uint32_t a_step = (p.a_mn ? 1024u : 32u) >> 4, b_step = (p.b_mn ? 1024u : 32u) >> 4;
            const uint32_t a_stage = (uint32_t)p.a_stage_bytes >> 4, b_stage = (uint32_t)b_stage_bytes >> 4;
            uint32_t s = 0, ph = 0;
            for (int kb = kb_beg; kb < kb_end; ++kb) {
                tc::mbar_wait(&full[s], ph);
                tc::fence_after_thread_sync();
                if (tc::elect_one()) {
                    const uint64_t dA = dA0 + (uint64_t)(s * a_stage), dB = dB0 + (uint64_t)(s * b_stage);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        tc::mma_tf32(tmem_base, dA + (uint64_t)(k4 * a_step), dB + (uint64_t)(k4 * b_step), idesc,
                                     (kb > kb_beg || k4 > 0) ? 1u : 0u);
                    if (EPI == EPI_WGRAD && p.db && by == 0) {
                        const uint32_t idesc1 = tc::make_idesc(2, TM, p.BN, 0, p.b_mn);
                        const uint64_t dOnes = tc::make_smem_desc(tc::smem_u32(sOnes), 16, 1024, 2);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc::mma_tf32(tmem_base + p.BN, dOnes + (uint64_t)(k4 * 2), dB + (uint64_t)(k4 * b_step), idesc1,
                                         (kb > kb_beg || k4 > 0) ? 1u : 0u);
                    }
                    tc::mma_commit(&empty[s]);
                    if (kb == kb_end - 1) tc::mma_commit(acc_full);
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
        }
        __syncwarp();
        tc::mbar_wait(acc_full, 0);
        tc::fence_after_thread_sync();
    }

    // ===================== epilogue: thread = output row =====================
    const int row = m0 + warp * 32 + lane;
    const bool rvalid = row < p.M;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float sq = 0.f, ab = 0.f;
    for (int c = 0; c < p.BN; c += 32) {
        if (n0 + c >= p.N) break;                                   // uniform across the CTA
        float v[32];
        if (nkb > 0) {
            tc::tmem_ld_32x32(taddr + c, v);
            tc::tmem_ld_wait(v);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        const int nb = n0 + c;
        if (!rvalid) continue;
        const int nv = min(32, p.N - nb);                           // valid columns of this chunk
        const long long off = (long long)g * p.c_gs + (long long)row * p.ldc + nb;
        if (EPI == EPI_BIAS_ACT) {
            float bv[32];
            if (p.bias) load_row(p.bias + (long long)g * p.bias_gs + nb, bv, nv, p.vecC && !(p.bias_gs & 3));
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) bv[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float x = v[j] + bv[j];
                v[j] = p.act == PGMVAE_ACT_SELU ? pg_selu(x) : (p.act == PGMVAE_ACT_SIGMOID ? pg_sigmoid(x) : x);
            }
            store_row(p.C + off, v, nv, p.vecC);
        } else if (EPI == EPI_SIGMOID_MSE) {
            float yv[32], o[32], bv[32];
            if (p.bias) load_row(p.bias + (long long)g * p.bias_gs + nb, bv, nv, p.vecC && !(p.bias_gs & 3));
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) bv[j] = 0.f;
            }
            load_row(p.aux + (long long)row * p.ldaux + nb, yv, nv, p.vecC);
            const int self = p.g0 + g;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                o[j] = pg_sigmoid(v[j] + bv[j]);
                float dpre = 0.f;
                if (j < nv && nb + j != self) {
                    const float d = o[j] - yv[j];
                    sq = fmaf(d, d, sq);
                    ab += fabsf(d);
                    dpre = p.gscale * d * o[j] * (1.0f - o[j]);
                }
                v[j] = dpre;
            }
            store_row(p.C + off, v, nv, p.vecC);
            if (p.C2) store_row(p.C2 + off, o, nv, p.vecC);
        } else if (EPI == EPI_DGRAD) {
            if (p.z) {
                float zv[32], qv[32];
                const long long zo = (long long)g * p.zq_gs + (long long)row * p.ldzq + nb;
                load_row(p.z + zo, zv, nv, p.vecC);
                load_row(p.q + zo, qv, nv, p.vecC);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaf(p.cscale, zv[j] - qv[j], v[j]);
            }
            if (p.aux) {
                float hv[32];
                load_row(p.aux + (long long)g * p.aux_gs + (long long)row * p.ldaux + nb, hv, nv, p.vecC);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (p.act == PGMVAE_ACT_SELU) v[j] *= pg_dselu_from_out(hv[j]);
                    else if (p.act == PGMVAE_ACT_SIGMOID) v[j] *= hv[j] * (1.0f - hv[j]);
                }
            }
            store_row(p.C + off, v, nv, p.vecC);
        } else {  // EPI_WGRAD: rows are weight rows (input features), reduced over the batch splits
            if (p.zero_row_base >= 0 && row == p.zero_row_base + g) continue;
            float* dst = p.C + off;
            if (p.vecC) {                       // 16-byte aligned rows: four partial sums per reduction
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (j + 4 <= nv) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                                     "f"(v[j + 2]), "f"(v[j + 3])
                                     : "memory");
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (j + i < nv) atomicAdd(dst + j + i, v[j + i]);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nv) atomicAdd(dst + j, v[j]);
            }
        }
    }
    if (EPI == EPI_WGRAD && p.db && nkb > 0 && by == 0 && warp == 0) {      // lane 0 owns TMEM lane 0
        // row 0 of the second accumulator: sum_b dY[b, n0 .. n0 + BN)
        for (int c = 0; c < p.BN && n0 + c < p.N; c += 32) {
            float v[32];
            tc::tmem_ld_32x32(tmem_base + p.BN + c, v);
            tc::tmem_ld_wait(v);
            float* dst = p.db + (long long)g * p.db_gs + n0 + c;
            const int nv = min(32, p.N - (n0 + c));
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < nv) atomicAdd(dst + j, v[j]);
            }
        }
    }
    if (EPI == EPI_SIGMOID_MSE) {
        const double dsq = pg_warp_sum_d((double)sq), dab = pg_warp_sum_d((double)ab);
        if (lane == 0) { red[0][warp] = dsq; red[1][warp] = dab; }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (EPI == EPI_SIGMOID_MSE && threadIdx.x == 0) {
        atomicAdd(p.acc, red[0][0] + red[0][1] + red[0][2] + red[0][3]);
        atomicAdd(p.acc + 1, red[1][0] + red[1][1] + red[1][2] + red[1][3]);
    }
    if (warp == 2) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

template <int EPI>
__global__ void __launch_bounds__(128) dense_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                       const __grid_constant__ CUtensorMap mapB,
                                                       const __grid_constant__ DenseTcP p) {
    dense_tc_body<EPI>(mapA, mapB, p, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z);
}

// All weight-gradient GEMMs of a training step in ONE launch: ten launches of ~25 us each spend a third of their
// time ramping up and draining (each is a single wave of CTAs); one launch of all their CTAs streams through.
// CTAs are numbered problem by problem, the longest-running problems first.
constexpr int WG_MAX = 10;
struct WgradMultiP {
    CUtensorMap mA[WG_MAX], mB[WG_MAX];
    DenseTcP p[WG_MAX];
    int start[WG_MAX + 1];           // first CTA of every problem
    int nx[WG_MAX], ny[WG_MAX];      // grid extents of every problem (z = the rest)
    int n;
};
__global__ void __launch_bounds__(128) wgrad_multi_kernel(const __grid_constant__ WgradMultiP P) {
    int l = 0;
    while (l + 1 < P.n && (int)blockIdx.x >= P.start[l + 1]) ++l;
    int r = (int)blockIdx.x - P.start[l];
    const int bx = r % P.nx[l];
    r /= P.nx[l];
    const int by = r % P.ny[l];
    dense_tc_body<EPI_WGRAD>(P.mA[l], P.mB[l], P.p[l], bx, by, r / P.ny[l]);
}

inline bool tma_ok(const float* p, int64_t gs, int ld) { return !((uintptr_t)p & 15) && ld % 4 == 0 && gs % 4 == 0; }

int pick_bn(int N) {
    int bn = pg_round_up(N, 32);
    return bn > 128 ? 128 : bn;
}

// stage geometry of one problem; returns the dynamic shared memory its CTAs need
template <int EPI>
size_t configure_tc(DenseTcP& p) {
    p.tmem_cols = 32;
    while (p.tmem_cols < (EPI == EPI_WGRAD ? 2 : 1) * p.BN) p.tmem_cols <<= 1;
    // MN-major A (wgrad) with a single M tile: load only the 32-row panels that hold valid rows
    p.apan = TM / 32;
    if (p.a_mn && p.M <= TM) p.apan = (int)pg_cdiv(p.M, 32);
    p.a_stage_bytes = p.a_mn ? p.apan * 4096 : A_STAGE_BYTES;
    const size_t stage = (size_t)p.a_stage_bytes + (size_t)p.BN * 128;
    int kb_cta = p.kb_per_split < p.kblocks ? p.kb_per_split : p.kblocks;
    if (kb_cta < 1) kb_cta = 1;
    // wgrad streams long k ranges: as many stages as keep two CTAs per SM (~100 KB each), at most 8
    // (PGMVAE_WGRAD_SMEM_KB: under data parallelism a smaller budget leaves shared memory on every SM for the
    //  NCCL CTAs that run beside the wgrad kernels)
    int budget_kb = 100;
    if (const char* ev = getenv("PGMVAE_WGRAD_SMEM_KB")) budget_kb = atoi(ev) >= 48 ? atoi(ev) : 100;
    int max_stages = EPI == EPI_WGRAD ? (int)((budget_kb * 1024 - 2 * A_STAGE_BYTES) / stage) : 3;
    if (max_stages > MAX_STAGES) max_stages = MAX_STAGES;
    if (max_stages < 3) max_stages = 3;
    p.stages = kb_cta < max_stages ? kb_cta : max_stages;
    return 1024 + p.stages * stage + (A_STAGE_BYTES - p.a_stage_bytes) + (EPI == EPI_WGRAD ? A_STAGE_BYTES : 0) + 256;
}

template <int EPI>
int launch_tc(pgmvae_ctx* ctx, cudaStream_t st, DenseTcP& p, const CUtensorMap& mapA, const CUtensorMap& mapB, int G,
              const char* name, double bytes) {
    const size_t smem = configure_tc<EPI>(p);
    static size_t configured[16] = {};          // per device: the attribute is set per device
    if (smem > configured[ctx->device & 15]) {
        PG_CUDA(cudaFuncSetAttribute(dense_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15] = smem;
    }
    dim3 grid((unsigned)pg_cdiv(p.N, p.BN), (unsigned)pg_cdiv(p.M, TM), (unsigned)(G * p.S));
    if (grid.y > 65535u || grid.z > 65535u) {
        pgmvae_set_error("dense (tensor core): grid too large (%u,%u,%u)", grid.x, grid.y, grid.z);
        return PGMVAE_EINVAL;
    }
    PG_KERNEL(ctx, st, name, bytes, 2.0 * G * (double)p.M * p.N * p.K);
    dense_tc_kernel<EPI><<<grid, 128, smem, st>>>(mapA, mapB, p);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

}  // namespace

bool pg_dense_tc_supported(const float* a, int64_t a_gs, int lda, const float* b, int64_t b_gs, int ldb) {
    return tma_ok(a, a_gs, lda) && tma_ok(b, b_gs, ldb);
}

int pg_dense_fwd_tc(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w,
                    int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, float* out, int64_t out_gs, int ldo, int G,
                    int B, int in, int out_dim, int act) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    DenseTcP p{};
    p.M = B; p.N = out_dim; p.K = in; p.BN = pick_bn(out_dim);
    p.kblocks = (int)pg_cdiv(in, KBLK); p.S = 1; p.kb_per_split = p.kblocks;
    p.a_mn = 0; p.b_mn = 1;
    p.C = out; p.c_gs = out_gs; p.ldc = ldo; p.bias = bias; p.bias_gs = bias_gs; p.act = act;
    p.a_shared = x_gs == 0; p.vecC = tma_ok(out, out_gs, ldo);
    CUtensorMap mA, mB;
    PG_TRY(tc::make_map(&mA, x, 4, (uint64_t)in, (uint64_t)B, (uint64_t)G, (uint64_t)ldx, (uint64_t)x_gs, 32, TM));
    PG_TRY(tc::make_map(&mB, w, 4, (uint64_t)out_dim, (uint64_t)in, (uint64_t)G, (uint64_t)ldw, (uint64_t)w_gs, 32, 32, true));
    const double xg = x_gs == 0 ? 1.0 : (double)G;
    return launch_tc<EPI_BIAS_ACT>(ctx, st, p, mA, mB, G, "dense_fwd_tc",
                                   4.0 * (xg * B * in + (double)G * in * out_dim + (double)G * out_dim +
                                          (double)G * B * out_dim));
}

int pg_dense_fwd_sigmoid_mse_tc(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w,
                                int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, const float* y, int ldy,
                                float* dpre, int64_t dpre_gs, int ldd, float* out_opt, double* acc2, int G, int g0, int B,
                                int in, int V, float grad_scale) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    DenseTcP p{};
    p.M = B; p.N = V; p.K = in; p.BN = pick_bn(V);
    p.kblocks = (int)pg_cdiv(in, KBLK); p.S = 1; p.kb_per_split = p.kblocks;
    p.a_mn = 0; p.b_mn = 1;
    p.C = dpre; p.c_gs = dpre_gs; p.ldc = ldd; p.C2 = out_opt; p.bias = bias; p.bias_gs = bias_gs;
    p.aux = y; p.ldaux = ldy; p.acc = acc2; p.gscale = grad_scale; p.g0 = g0;
    p.a_shared = x_gs == 0;
    p.vecC = tma_ok(dpre, dpre_gs, ldd) && tma_ok(y, 0, ldy) && (!out_opt || tma_ok(out_opt, dpre_gs, ldd));
    CUtensorMap mA, mB;
    PG_TRY(tc::make_map(&mA, x, 4, (uint64_t)in, (uint64_t)B, (uint64_t)G, (uint64_t)ldx, (uint64_t)x_gs, 32, TM));
    PG_TRY(tc::make_map(&mB, w, 4, (uint64_t)V, (uint64_t)in, (uint64_t)G, (uint64_t)ldw, (uint64_t)w_gs, 32, 32, true));
    return launch_tc<EPI_SIGMOID_MSE>(ctx, st, p, mA, mB, G, "dense_fwd_sigmoid_mse_tc",
                                      4.0 * ((double)G * B * in + (double)G * in * V + (double)G * V + (double)B * V +
                                             (double)G * B * V * (out_opt ? 2 : 1)));
}

int pg_dense_dgrad_tc(pgmvae_ctx* ctx, cudaStream_t st, const float* dy, int64_t dy_gs, int lddy, const float* w,
                      int64_t w_gs, int ldw, const float* h_in, int64_t h_gs, int ldh, const float* z, const float* q,
                      int64_t zq_gs, int ldzq, float cscale, float* dx, int64_t dx_gs, int lddx, int G, int B, int in,
                      int out_dim, int act_below) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    DenseTcP p{};
    p.M = B; p.N = in; p.K = out_dim; p.BN = pick_bn(in);
    p.kblocks = (int)pg_cdiv(out_dim, KBLK); p.S = 1; p.kb_per_split = p.kblocks;
    p.a_mn = 0; p.b_mn = 0;
    p.C = dx; p.c_gs = dx_gs; p.ldc = lddx; p.aux = h_in; p.aux_gs = h_gs; p.ldaux = ldh; p.act = act_below;
    p.z = z; p.q = q; p.zq_gs = zq_gs; p.ldzq = ldzq; p.cscale = cscale;
    p.a_shared = dy_gs == 0;
    p.vecC = tma_ok(dx, dx_gs, lddx) && (!h_in || tma_ok(h_in, h_gs, ldh)) && (!z || (tma_ok(z, zq_gs, ldzq) && tma_ok(q, zq_gs, ldzq)));
    CUtensorMap mA, mB;
    PG_TRY(tc::make_map(&mA, dy, 4, (uint64_t)out_dim, (uint64_t)B, (uint64_t)G, (uint64_t)lddy, (uint64_t)dy_gs, 32, TM));
    // W[in][out] read as rows n = in, contiguous k = out
    PG_TRY(tc::make_map(&mB, w, 4, (uint64_t)out_dim, (uint64_t)in, (uint64_t)G, (uint64_t)ldw, (uint64_t)w_gs, 32,
                        (uint32_t)p.BN));
    return launch_tc<EPI_DGRAD>(ctx, st, p, mA, mB, G, "dense_dgrad_tc",
                                4.0 * ((double)G * B * out_dim + (double)G * in * out_dim +
                                       (double)G * B * in * (h_in ? 2 : 1) + (z ? 2.0 * G * B * in : 0.0)));
}

namespace {
// problem description + tensor maps of one weight-gradient GEMM; S = batch splits (0: fill one wave of two CTAs per SM)
int setup_wgrad(pgmvae_ctx* ctx, const float* x, int64_t x_gs, int ldx, const float* dy, int64_t dy_gs, int lddy,
                float* dw, int64_t dw_gs, int lddw, float* db, int64_t db_gs, int G, int B, int in, int out_dim,
                int zero_row_base, int64_t ctas_target, DenseTcP& p, CUtensorMap& mA, CUtensorMap& mB, double& bytes) {
    p = DenseTcP{};
    p.M = in; p.N = out_dim; p.K = B; p.BN = pick_bn(out_dim);
    p.kblocks = (int)pg_cdiv(B, KBLK);
    p.a_mn = 1; p.b_mn = 1;
    p.C = dw; p.c_gs = dw_gs; p.ldc = lddw; p.zero_row_base = zero_row_base;
    p.db = db; p.db_gs = db_gs;
    p.a_shared = x_gs == 0;
    p.vecC = tma_ok(dw, dw_gs, lddw);
    // batch splits: at least 8 k-blocks (256 samples) per CTA
    const int64_t tiles = pg_cdiv(out_dim, p.BN) * pg_cdiv(in, TM) * (int64_t)G;
    int S = (int)(ctas_target / (tiles > 0 ? tiles : 1));
    if (const char* ev = getenv("PGMVAE_WGRAD_SPLITS")) S = atoi(ev) > 0 ? atoi(ev) : S;
    const int maxS = (int)pg_cdiv(p.kblocks, 8);
    if (S > maxS) S = maxS;
    if (S < 1) S = 1;
    while ((int64_t)G * S > 65535 && S > 1) --S;
    p.kb_per_split = (int)pg_cdiv(p.kblocks, S);
    p.S = (int)pg_cdiv(p.kblocks, p.kb_per_split);
    // x[B][in] read as rows k = b, contiguous m = in; boxes of [32 b][32 m]
    PG_TRY(tc::make_map(&mA, x, 4, (uint64_t)in, (uint64_t)B, (uint64_t)G, (uint64_t)ldx, (uint64_t)x_gs, 32, 32, true));
    PG_TRY(tc::make_map(&mB, dy, 4, (uint64_t)out_dim, (uint64_t)B, (uint64_t)G, (uint64_t)lddy, (uint64_t)dy_gs, 32, 32, true));
    const double xg = x_gs == 0 ? 1.0 : (double)G;
    bytes = 4.0 * (xg * B * in + (double)G * B * out_dim + (double)G * in * out_dim + (db ? (double)G * out_dim : 0.0));
    return PGMVAE_OK;
}
}  // namespace

int pg_dense_wgrad_tc(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* dy,
                      int64_t dy_gs, int lddy, float* dw, int64_t dw_gs, int lddw, float* db, int64_t db_gs, int G, int B,
                      int in, int out_dim, int zero_row_base) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    DenseTcP p;
    CUtensorMap mA, mB;
    double bytes;
    // fill ONE wave of two CTAs per SM (a partial second wave costs a whole CTA time)
    PG_TRY(setup_wgrad(ctx, x, x_gs, ldx, dy, dy_gs, lddy, dw, dw_gs, lddw, db, db_gs, G, B, in, out_dim, zero_row_base,
                       (int64_t)ctx->sm_count * 2, p, mA, mB, bytes));
    PG_TRY(launch_tc<EPI_WGRAD>(ctx, st, p, mA, mB, G, "dense_wgrad_tc", bytes));
    return PGMVAE_OK;
}

bool pg_dense_wgrad_multi_supported(const PgWgradProblem* pr, int n) {
    if (n < 1 || n > WG_MAX) return false;
    for (int i = 0; i < n; ++i)
        if (!tma_ok(pr[i].x, pr[i].x_gs, pr[i].ldx) || !tma_ok(pr[i].dy, pr[i].dy_gs, pr[i].lddy)) return false;
    return true;
}

int pg_dense_wgrad_multi_tc(pgmvae_ctx* ctx, cudaStream_t st, const PgWgradProblem* pr, int n) {
    if (n < 1 || n > WG_MAX) {
        pgmvae_set_error("wgrad (multi): %d problems (1..%d)", n, WG_MAX);
        return PGMVAE_EINVAL;
    }
    WgradMultiP P;                        // 4.7 KB of kernel parameters
    double bytes = 0.0, flops = 0.0, cost[WG_MAX];
    size_t smem = 0;
    // batch splits: the fewest that still give ~4 waves of CTAs over all problems together (a CTA spends a fixed
    // ~3 us setting up and reducing its tile: measured 0.182 / 0.189 / 0.202 / 0.236 ms at 2 / 3 / 4 / 8 splits, cfg2)
    int64_t all_tiles = 0;
    for (int i = 0; i < n; ++i)
        if (pr[i].G > 0 && pr[i].B > 0)
            all_tiles += pg_cdiv(pr[i].out, pick_bn(pr[i].out)) * pg_cdiv(pr[i].in, TM) * (int64_t)pr[i].G;
    const int64_t splits = std::max<int64_t>(1, pg_cdiv((int64_t)ctx->sm_count * 2 * 4, std::max<int64_t>(1, all_tiles)));
    int order[WG_MAX], Gs[WG_MAX], cnt = 0;
    DenseTcP tp[WG_MAX];
    CUtensorMap tA[WG_MAX], tB[WG_MAX];
    for (int i = 0; i < n; ++i) {
        const PgWgradProblem& q = pr[i];
        if (q.G <= 0 || q.B <= 0) continue;
        double b;
        PG_TRY(setup_wgrad(ctx, q.x, q.x_gs, q.ldx, q.dy, q.dy_gs, q.lddy, q.dw, q.dw_gs, q.lddw, q.db, q.db_gs, q.G, q.B,
                           q.in, q.out, q.zero_row_base,
                           splits * pg_cdiv(q.out, pick_bn(q.out)) * pg_cdiv(q.in, TM) * (int64_t)q.G, tp[cnt], tA[cnt],
                           tB[cnt], b));
        smem = std::max(smem, configure_tc<EPI_WGRAD>(tp[cnt]));
        bytes += b;
        flops += 2.0 * q.G * (double)q.B * q.in * q.out;
        cost[cnt] = (double)tp[cnt].kb_per_split * (tp[cnt].a_stage_bytes + tp[cnt].BN * 128);   // bytes one CTA streams
        order[cnt] = cnt;
        Gs[cnt] = q.G;
        ++cnt;
    }
    if (cnt == 0) return PGMVAE_OK;
    for (int i = 1; i < cnt; ++i)                     // longest CTAs first
        for (int j = i; j > 0 && cost[order[j]] > cost[order[j - 1]]; --j) std::swap(order[j], order[j - 1]);
    int64_t total = 0;
    for (int i = 0; i < cnt; ++i) {
        const int o = order[i];
        P.p[i] = tp[o]; P.mA[i] = tA[o]; P.mB[i] = tB[o];
        P.nx[i] = (int)pg_cdiv(tp[o].N, tp[o].BN);
        P.ny[i] = (int)pg_cdiv(tp[o].M, TM);
        P.start[i] = (int)total;
        total += (int64_t)P.nx[i] * P.ny[i] * tp[o].S * Gs[o];
    }
    P.start[cnt] = (int)total;
    P.n = cnt;
    if (total > 0x7fffffff) {
        pgmvae_set_error("wgrad (multi): grid too large");
        return PGMVAE_EINVAL;
    }
    static size_t configured[16] = {};          // per device: the attribute is set per device
    if (smem > configured[ctx->device & 15]) {
        PG_CUDA(cudaFuncSetAttribute(wgrad_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15] = smem;
    }
    PG_KERNEL(ctx, st, "dense_wgrad_multi_tc", bytes, flops);
    wgrad_multi_kernel<<<(unsigned)total, 128, smem, st>>>(P);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}
