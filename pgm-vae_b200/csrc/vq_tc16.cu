// Kernel (b), fp16 tensor-core flavour (PGMVAE_PREC_BF16): single-pass VQ assignment with the
// |e|^2 term folded into the contraction, the z operand resident in TENSOR MEMORY and an
// (optional) fused EMA scatter.  Reference: core/quantizer.py:44-47 / :135-138 (distances +
// argmin) and :144-146 (per-code counts and sums); the [V,B,K] distance tensor and the one-hot
// matrix are never materialised.
//
// Score instead of distance:  s_k = z.e_k - |e_k|^2 / 2   (argmax s == argmin d, d = |z|^2 - 2 s).
// The correction rides in the contraction: the codebook copy is [e_k (fp16), hi, lo, 0..] with
// hi + lo = -|e_k|^2/2 split over two fp16 columns, and z is extended by two 1.0 columns, so the
// epilogue is a bare maximum (FMNMX3: two scores per instruction).
//
// Per CTA (persistent over row tiles of SUB x 128 samples of one variable):
//   z      each epilogue thread owns one row: reads it from HBM (fp32), rounds it to fp16 and
//          writes it with tcgen05.st into TMEM, where it stays as the A operand of every MMA of
//          the row tile (no fp16 copy of z in HBM, no shared-memory traffic for A)
//   E      [BN codes, KD] fp16 tiles, TMA -> shared-memory ring (128-byte swizzle), B operand
//   acc    SUB x [128, BN] fp32 in TMEM, double buffered (MMA of tile t+1 overlaps epilogue t)
//   warp 0 TMA producer | warps 1..SUB MMA issuers, one per 128-row sub-tile (warp 2 also owns the TMEM
//   allocation) | warps 4.. epilogue, four per sub-tile.  A sub-tile is its own pipeline (own issuer, own
//   acc_full / acc_empty barriers): its warps never wait for the slowest warp of another sub-tile.
//
// Exactness.  fp16 products only PROPOSE candidates; fp32 decides.  eps bounds the error of an
// approximate score (2^-10 |z| max|e| for round-to-nearest fp16 operands).  In ONE pass each
// row keeps a running maximum m and records every 8-code group whose maximum is >= m - 2 eps at
// the time it is seen: the group maximum, its first code and its eight scores go to a ring in
// shared memory.  m only grows, so the final candidate set {k : s~_k >= m_final - 2 eps} is a
// subset of the recorded scores, and the true arg-min is in it (header of vq_tc.cu); which codes
// they are is sorted out once, after the last tile.  A row with a single candidate is decided;
// rows with several are re-scored with exactly the fp32 arithmetic of the CUDA-core kernel
// (lowest index on ties).  Rows whose ring overflowed with still-relevant entries (many identical
// dead codes) or whose fp16 image is not finite go to the exact full-scan kernel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ops.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;                       // rows per MMA (TMEM lanes)
constexpr int KB_BYTES = 128;                 // one k-block = one 128-byte swizzle row = 64 halves
constexpr int RING = 8;                       // recorded 8-code groups kept per row: (group max, first code) + the 8 scores
constexpr int MAX_STAGES = 6;
constexpr int A_COLS = 128;                   // TMEM columns reserved for the z operand

template <int SUB>
struct Cfg {
    static constexpr int BN = 64;                          // codes per tile: 128 + 2*SUB*BN <= 512 columns; two chunks
    static constexpr int NCH = BN / 32;                    // 32-column chunks per tile
    static constexpr int TMR = TM * SUB;                   // rows per CTA row tile
    static constexpr int EPI_WARPS = 4 * SUB;
    static constexpr int THREADS = 128 + 32 * EPI_WARPS;
};

struct Vq16P {
    int G, B, D, K;
    int KD, ksteps, kblocks, tiles_m, tiles_n, Kpad, stages, zvec, qvec;
    const float* z; long long z_gs; int ldz;
    const float* e; long long e_gs; int lde;
    const float* ee;        // [G][Kpad] exact fp32 squared norms
    const float* emax;      // [1] max_k |e_k| over all groups
    int32_t* idx; long long idx_gs;
    float* best; float* gap;
    float* cnt; long long c_gs;             // fused EMA statistics (optional)
    float* dw; long long dw_gs; int lddw;
    int* flag_count; int2* flag_list;
    float margin_scale, margin_abs;
    // fused quantise step (optional; core/quantizer.py:141-142,156): q = E[idx] (fp32), the straight-through output
    // z + (q - z) as the bf16 operand of the first decoder layer, and sum (q - z)^2 into *loss
    float* q; __nv_bfloat16* stb; long long q_gs; int ldq; double* loss;
};

// Per code: exact |e|^2 (sequential fmaf, the order every fp32 path uses), the fp16 row
// [e, hi, lo, 0..] with hi + lo = -|e|^2/2, and max |e|.  Padding codes k >= K score -60000.
__global__ void e_prep_kernel(const float* __restrict__ e, long long e_gs, int lde, float* __restrict__ ee,
                              float* __restrict__ emax, __half* __restrict__ e16, int G, int K, int Kpad, int D, int KD) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f;
    if (i < (long long)G * Kpad) {
        const int g = (int)(i / Kpad), k = (int)(i - (long long)g * Kpad);
        __half* dst = e16 + i * KD;
        if (k < K) {
            const float* row = e + (long long)g * e_gs + (long long)k * lde;
            for (int d = 0; d < D; ++d) {
                const float v = row[d];
                s = fmaf(v, v, s);
                dst[d] = __float2half_rn(v);
            }
            ee[i] = s;
            const float h = -0.5f * s;
            const __half hi = __float2half_rn(h);
            dst[D] = hi;
            dst[D + 1] = __float2half_rn(h - __half2float(hi));
            for (int d = D + 2; d < KD; ++d) dst[d] = __half(0.f);
        } else {
            ee[i] = INFINITY;
            for (int d = 0; d < KD; ++d) dst[d] = __half(d == D ? -60000.f : 0.f);
        }
    }
    float m = sqrtf(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(emax), __float_as_int(m));   // m >= 0
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// EMA statistics of one row: counts[k] += 1, dw[k,:] += z[row,:]   (core/quantizer.py:144-146)
__device__ __forceinline__ void scatter_row(const Vq16P& p, int g, const float* zr, int k) {
    float* dst = p.dw + (long long)g * p.dw_gs + (long long)k * p.lddw;
    if (p.zvec && !(p.lddw & 3) && !(p.D & 3)) {
        for (int d = 0; d < p.D; d += 4) {
            const float4 v = *reinterpret_cast<const float4*>(zr + d);
            red_add_v4(dst + d, v.x, v.y, v.z, v.w);
        }
    } else {
        for (int d = 0; d < p.D; ++d) atomicAdd(dst + d, zr[d]);
    }
    atomicAdd(p.cnt + (long long)g * p.c_gs + k, 1.0f);
}

// EMA statistics of the 32 rows a warp has just decided, WARP-AGGREGATED (north_star (c)): lanes that chose the same
// code are found with match.any, the whole warp sums the rows of one code -- lane l owns floats 2l, 2l+1 (+64, ...) of the
// vector -- and issues ONE set of reductions per distinct code instead of one per row.  With a trained codebook the 32
// rows mostly differ and the per-row 128-bit reductions (scatter_row) are cheaper, so the warp takes this path only
// when at most 8 distinct codes are present (right after initialisation nearly all rows of a variable share a code:
// 4096 rows x 16 reductions into the same 256 bytes serialise in L2).  code < 0: the lane has nothing to add.
__device__ __forceinline__ void scatter_warp(const Vq16P& p, int g, const float* zr, int code, int lane) {
    const unsigned full = 0xffffffffu;
    const unsigned active = __ballot_sync(full, code >= 0);
    if (!active) return;
    const unsigned peers = __match_any_sync(full, code);
    const bool lead = code >= 0 && (__ffs(peers) - 1) == lane;
    const int ngroups = __popc(__ballot_sync(full, lead));
    const bool pairs_ok = p.zvec && !(p.D & 1) && !(p.lddw & 1) && !(p.dw_gs & 1);
    if (ngroups > 8 || !pairs_ok) {
        if (code >= 0) scatter_row(p, g, zr, code);
        return;
    }
    unsigned todo = __ballot_sync(full, lead);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned grp = __shfl_sync(full, peers, src);
        const int k = __shfl_sync(full, code, src);
        float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};          // D <= 128: two float2 per lane
        unsigned mm = grp;
        while (mm) {
            const int r = __ffs(mm) - 1;
            mm &= mm - 1;
            const float* row = reinterpret_cast<const float*>(__shfl_sync(full, (unsigned long long)zr, r));
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int d = 2 * lane + 64 * j;
                if (d < p.D) {
                    const float2 v = *reinterpret_cast<const float2*>(row + d);
                    acc[j].x += v.x; acc[j].y += v.y;
                }
            }
        }
        float* dst = p.dw + (long long)g * p.dw_gs + (long long)k * p.lddw;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int d = 2 * lane + 64 * j;
            if (d < p.D)
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + d), "f"(acc[j].x), "f"(acc[j].y) : "memory");
        }
        if (lane == 0) atomicAdd(p.cnt + (long long)g * p.c_gs + k, (float)__popc(grp));
    }
}

// quantise one row with its decided code: q, straight-through output (bf16), returns sum_d (q - z)^2
__device__ __forceinline__ float quantize_row(const Vq16P& p, int g, int row, const float* zr, int k) {
    const float* er = p.e + (long long)g * p.e_gs + (long long)k * p.lde;
    const long long o = (long long)g * p.q_gs + (long long)row * p.ldq;
    float part = 0.f;
    if (p.qvec) {
        for (int d = 0; d < p.D; d += 8) {
            float ev[8], zv[8], sv[8];
            *reinterpret_cast<float4*>(ev) = __ldg(reinterpret_cast<const float4*>(er + d));
            *reinterpret_cast<float4*>(ev + 4) = __ldg(reinterpret_cast<const float4*>(er + d + 4));
            *reinterpret_cast<float4*>(zv) = *reinterpret_cast<const float4*>(zr + d);
            *reinterpret_cast<float4*>(zv + 4) = *reinterpret_cast<const float4*>(zr + d + 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float diff = ev[j] - zv[j];
                part = fmaf(diff, diff, part);
                sv[j] = zv[j] + diff;                    // inputs + stop_gradient(quantized - inputs)
            }
            if (p.q) {
                *reinterpret_cast<float4*>(p.q + o + d) = *reinterpret_cast<const float4*>(ev);
                *reinterpret_cast<float4*>(p.q + o + d + 4) = *reinterpret_cast<const float4*>(ev + 4);
            }
            if (p.stb) {
                __nv_bfloat162 h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(sv[2 * j], sv[2 * j + 1]);
                *reinterpret_cast<uint4*>(p.stb + o + d) = *reinterpret_cast<const uint4*>(h);
            }
        }
    } else {
        for (int d = 0; d < p.D; ++d) {
            const float ev = __ldg(er + d), zv = zr[d];
            const float diff = ev - zv;
            part = fmaf(diff, diff, part);
            if (p.q) p.q[o + d] = ev;
            if (p.stb) p.stb[o + d] = __float2bfloat16_rn(zv + diff);
        }
    }
    return part;
}

// (16 warps x 104 registers: one CTA of the data-parallel exchange kernel -- model.cu: p2p_shard_adam_kernel, 4 warps x
// 80 registers -- fits on the SM next to this one)
template <int SUB>
__global__ void __maxnreg__(SUB == 3 ? 104 : 128)
vq_assign_f16_kernel(const __grid_constant__ CUtensorMap mapE, const Vq16P p) {
    using C = Cfg<SUB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a shared-space pointer
    constexpr int B_KB_BYTES = C::BN * KB_BYTES;                       // one k-block of one code tile
    const int stage_bytes = p.kblocks * B_KB_BYTES;
    uint8_t* sB = smem;                                                // [stages][kblocks][BN * 128]
    // record ring, three row-contiguous planes (a warp's 32 rows are 16 / 8 bytes apart: conflict-free stores)
    float4* ringA = reinterpret_cast<float4*>(sB + (size_t)p.stages * stage_bytes);   // [RING][TMR] scores 0..3
    float4* ringB = ringA + RING * C::TMR;                                           // [RING][TMR] scores 4..7
    float2* ringH = reinterpret_cast<float2*>(ringB + RING * C::TMR);                // [RING][TMR] (group max, first code)
    uint64_t* bars = reinterpret_cast<uint64_t*>(ringH + RING * C::TMR);
    // every 128-row sub-tile is its own pipeline (own issuer warp, own barriers): its four epilogue warps only
    // ever wait for each other, not for the slowest of all 4 * SUB warps of the CTA
    uint64_t* a_full = bars + 0;                   // [3]      z of the sub-tile sits in TMEM
    uint64_t* acc_full = bars + 3;                 // [3][2]
    uint64_t* acc_empty = bars + 9;                // [3][2]
    uint64_t* b_full = bars + 15;                  // [MAX_STAGES]
    uint64_t* b_empty = bars + 15 + MAX_STAGES;    // [MAX_STAGES]   one arrival per sub-tile pipeline
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15 + 2 * MAX_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = p.G * p.tiles_m;

    if (warp == 0 && lane == 0) tc::tma_prefetch_desc(&mapE);
    if (warp == 1 && lane == 0) {
        for (int u = 0; u < 3; ++u) {
            tc::mbar_init(&a_full[u], 4);                    // one arrival per epilogue warp of the sub-tile
            for (int s = 0; s < 2; ++s) {
                tc::mbar_init(&acc_full[u * 2 + s], 1);
                tc::mbar_init(&acc_empty[u * 2 + s], 4);
            }
        }
        for (int s = 0; s < MAX_STAGES; ++s) {
            tc::mbar_init(&b_full[s], 1);
            tc::mbar_init(&b_empty[s], SUB);
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

    if (warp == 0) {
        // ===================== TMA producer: code tiles (warp-uniform loop, one elected lane issues)
        uint32_t s = 0, ph = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int g = item / p.tiles_m;
            for (int t = 0; t < p.tiles_n; ++t) {
                tc::mbar_wait(&b_empty[s], ph ^ 1);
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&b_full[s], (uint32_t)stage_bytes);
                    for (int kb = 0; kb < p.kblocks; ++kb)
                        tc::tma_load_3d(sB + (size_t)s * stage_bytes + (size_t)kb * B_KB_BYTES, &mapE, &b_full[s],
                                        kb * 64, t * C::BN, g);
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 1 && warp <= SUB) {
        // ===================== MMA issuers: warp 1 + u serves sub-tile u (warp-uniform loop, one elected lane issues)
        const int u = warp - 1;
        const uint32_t idesc = tc::make_idesc(0, TM, C::BN, 0, 0);
        const uint64_t descB0 = tc::make_smem_desc(tc::smem_u32(sB), 16, 1024);
        const uint32_t stage_stride = (uint32_t)stage_bytes >> 4;
        const uint32_t a_tmem = tmem_base + u * (uint32_t)(p.KD >> 1);     // this sub-tile's z operand
        uint32_t it = 0, item_n = 0, s = 0, ph = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_n) {
            tc::mbar_wait(&a_full[u], item_n & 1);           // z of this sub-tile sits in TMEM
            for (int t = 0; t < p.tiles_n; ++t, ++it) {
                const uint32_t ab = it & 1, aph = (it >> 1) & 1;
                tc::mbar_wait(&b_full[s], ph);
                tc::mbar_wait(&acc_empty[u * 2 + ab], aph ^ 1);
                tc::fence_after_thread_sync();
                if (tc::elect_one()) {
                    const uint64_t descB = descB0 + (uint64_t)(s * stage_stride);
                    const uint32_t d_tmem = tmem_base + A_COLS + (ab * SUB + u) * C::BN;
                    constexpr int KBB = C::BN * KB_BYTES;
                    if (p.ksteps == 5) {         // D = 64 (+ the two correction columns): straight-line, literal offsets
#pragma unroll
                        for (int ks = 0; ks < 5; ++ks)
                            tc::mma_f16_ts(d_tmem, a_tmem + ks * 8, descB + ((uint32_t)((ks >> 2) * KBB + (ks & 3) * 32) >> 4),
                                           idesc, ks > 0 ? 1u : 0u);
                    } else {
                        for (int ks = 0; ks < p.ksteps; ++ks) {
                            const uint32_t offB = (uint32_t)((ks >> 2) * KBB + (ks & 3) * 32) >> 4;
                            tc::mma_f16_ts(d_tmem, a_tmem + ks * 8, descB + offB, idesc, ks > 0 ? 1u : 0u);
                        }
                    }
                    tc::mma_commit(&b_empty[s]);             // the stage is free once every sub-tile's MMAs have read it
                    tc::mma_commit(&acc_full[u * 2 + ab]);   // scores ready for this sub-tile's epilogue warps
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: one row per thread =====================
        const int q = warp & 3;                        // TMEM lane quarter == warp % 4
        const int sub = (warp - 4) >> 2;               // 128-row sub-tile
        const int r = sub * TM + q * 32 + lane;
        const float emax = *p.emax;
        float4* myA = ringA + r;                       // slot i at my?[i * TMR]
        float4* myB = ringB + r;
        float2* myH = ringH + r;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t it = 0;
        double qloss = 0.0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int g = item / p.tiles_m, mt = item - g * p.tiles_m;
            const int row = mt * C::TMR + r;
            const bool valid = row < p.B;
            const float* zr = p.z + (long long)g * p.z_gs + (long long)(valid ? row : 0) * p.ldz;
            // ---- z row: fp32 -> fp16 (+ two 1.0 columns) -> TMEM; |z|^2 in the reference's order
            float zz = 0.f;
            {
                const uint32_t a_addr = lane_addr + sub * (uint32_t)(p.KD >> 1);
                for (int ks = 0; ks < p.ksteps; ++ks) {
                    const int d0 = ks * 16;
                    float f[16];
                    if (p.zvec && d0 + 16 <= p.D) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 v = *reinterpret_cast<const float4*>(zr + d0 + j);
                            f[j] = v.x; f[j + 1] = v.y; f[j + 2] = v.z; f[j + 3] = v.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int d = d0 + j;
                            f[j] = d < p.D ? zr[d] : ((d == p.D || d == p.D + 1) ? 1.0f : 0.f);
                        }
                    }
                    uint32_t u[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        if (d0 + j < p.D) zz = fmaf(f[j], f[j], zz);
                        if (d0 + j + 1 < p.D) zz = fmaf(f[j + 1], f[j + 1], zz);
                        const __half2 h = __floats2half2_rn(f[j], f[j + 1]);
                        u[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    tc::tmem_st_32x8(a_addr + ks * 8, u);
                }
                tc::tmem_st_wait();
                tc::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&a_full[sub]);
            }
            const float znorm = sqrtf(zz);
            // margins in distance units (d = |z|^2 - 2 s); the scores use half of them
            // operand rounding (relative to |z| |e|) + fp32 accumulation (relative to the magnitudes) + fp16
            // subnormals (|x| < 6e-5 carries an absolute error of 2^-25 per element)
            const float margin_d = p.margin_scale * znorm * emax + p.margin_abs * (znorm + emax) * (znorm + emax) +
                                   1.0e-6f * (znorm + emax);
            const float margin = 0.5f * margin_d;
            // the fp16 image must be finite and the codebook representable
            // (|z_d| < 65504; -|e|^2/2 must stay above the -60000 of the padding codes)
            const bool zbad = !(zz < 1.0e9f) || !(emax < 250.0f);
            float runmax = -INFINITY, thr = -INFINITY, evmax = -INFINITY;
            int cnt = 0;

            static_assert(C::NCH == 2, "the epilogue pipeline below is written for two 32-column chunks per tile");
            const uint32_t acc_addr = lane_addr + A_COLS + sub * C::BN;
            auto consume = [&](const float (&v)[32], int n) {
                // maxima of the four 8-code groups and of the chunk: 18 instructions for 32 scores
                float gm[4];
#pragma unroll
                for (int qq = 0; qq < 4; ++qq)
                    gm[qq] = tc::max3(tc::max3(v[8 * qq], v[8 * qq + 1], v[8 * qq + 2]),
                                      tc::max3(v[8 * qq + 3], v[8 * qq + 4], v[8 * qq + 5]),
                                      fmaxf(v[8 * qq + 6], v[8 * qq + 7]));
                const float cm = fmaxf(tc::max3(gm[0], gm[1], gm[2]), gm[3]);
                if (__any_sync(0xffffffffu, cm >= thr)) {
                    // some row of this warp has a score within the band of its running maximum: that row records
                    // every 8-code group whose maximum is in the band -- the group's eight scores go to the ring as
                    // they are (three stores), which code(s) matter is sorted out once, at the end of the row
                    runmax = fmaxf(runmax, cm);
                    thr = runmax - margin;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        if (gm[qq] >= thr) {
                            const int sl = (cnt & (RING - 1)) * C::TMR;
                            if (cnt >= RING) evmax = fmaxf(evmax, myH[sl].x);
                            myH[sl] = make_float2(gm[qq], __int_as_float(n * 32 + 8 * qq));
                            myA[sl] = make_float4(v[8 * qq], v[8 * qq + 1], v[8 * qq + 2], v[8 * qq + 3]);
                            myB[sl] = make_float4(v[8 * qq + 4], v[8 * qq + 5], v[8 * qq + 6], v[8 * qq + 7]);
                            ++cnt;
                        }
                    }
                }
            };
            // Straight-line pipeline over (tile, chunk): the load of the next chunk -- also across tile boundaries --
            // is in flight while the current one is reduced; an accumulator buffer goes back to the MMA warp as
            // soon as its second chunk sits in registers.  (A first version drove this with generic
            // issue/retire counters: ~40 of its ~128 instructions per chunk were loop bookkeeping.)
            float va[32], vb[32];
            {
                const uint32_t ab = it & 1;
                tc::mbar_wait(&acc_full[sub * 2 + ab], (it >> 1) & 1);
                tc::fence_after_thread_sync();
                tc::tmem_ld_32x32(acc_addr + ab * SUB * C::BN, va);
            }
            for (int t = 0; t < p.tiles_n; ++t) {
                const uint32_t git = it + t, ab = git & 1;
                tc::tmem_ld_wait(va);
                tc::tmem_ld_32x32(acc_addr + ab * SUB * C::BN + 32, vb);
                consume(va, 2 * t);
                tc::tmem_ld_wait(vb);
                tc::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc_empty[sub * 2 + ab]);          // both chunks of this buffer are in registers
                if (t + 1 < p.tiles_n) {
                    const uint32_t nb = (git + 1) & 1;
                    tc::mbar_wait(&acc_full[sub * 2 + nb], ((git + 1) >> 1) & 1);
                    tc::fence_after_thread_sync();
                    tc::tmem_ld_32x32(acc_addr + nb * SUB * C::BN, va);
                }
                consume(vb, 2 * t + 1);
            }
            it += p.tiles_n;

            int my_code = -1;                               // decided code of this lane's row (for the warp-aggregated scatter)
            if (valid) {
                const long long o = (long long)g * p.idx_gs + row;
                const float thr_f = runmax - margin;
                // NaN-safe: every comparison below is false for NaN, which routes the row to the full scan
                bool ok = !zbad && (runmax > -3.0e38f) && (runmax < 3.0e38f) && !(evmax >= thr_f) && cnt > 0;
                int nq = 0, k0 = 0;
                const int nring = cnt < RING ? cnt : RING;
                // candidates: codes of the recorded groups whose score lies in the band of the FINAL maximum
                // (bit j of cmask[i]: code j of ring entry i)
                uint32_t cmask[RING];
#pragma unroll
                for (int i = 0; i < RING; ++i) {
                    cmask[i] = 0u;
                    if (ok && i < nring && myH[i * C::TMR].x >= thr_f) {
                        const float4 a = myA[i * C::TMR], b = myB[i * C::TMR];
                        const int kb = __float_as_int(myH[i * C::TMR].y);
                        const float sc[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (sc[j] >= thr_f && kb + j < p.K) { cmask[i] |= 1u << j; k0 = kb + j; ++nq; }
                    }
                }
                if (ok) ok = nq >= 1;
                if (ok) {
                    int bi = k0;
                    if (nq > 1 || p.best || p.gap) {
                        // exact fp32 distances of the candidates: same arithmetic as the CUDA-core kernel
                        float best = INFINITY, second = INFINITY;
                        bi = 0x7fffffff;
                        const float* eg = p.e + (long long)g * p.e_gs;
#pragma unroll
                        for (int i = 0; i < RING; ++i) {
                            uint32_t mk = cmask[i];
                            const int kb = mk ? __float_as_int(myH[i * C::TMR].y) : 0;
                            while (mk) {
                                const int k = kb + __ffs(mk) - 1;
                                mk &= mk - 1;
                                const float* er = eg + (long long)k * p.lde;
                                float acc = 0.f;
                                for (int d = 0; d < p.D; ++d) acc = fmaf(zr[d], __ldg(er + d), acc);
                                const float dist = (zz - 2.0f * acc) + __ldg(p.ee + (long long)g * p.Kpad + k);
                                if (dist < best || (dist == best && k < bi)) { second = best; best = dist; bi = k; }
                                else if (dist < second) second = dist;
                            }
                        }
                        if (p.best) p.best[o] = best;
                        // exact gap when a runner-up lies inside the error band, otherwise a lower bound
                        if (p.gap) p.gap[o] = nq > 1 ? second - best : margin_d;
                    }
                    p.idx[o] = bi;
                    my_code = bi;
                    if (p.q || p.stb || p.loss) qloss += (double)quantize_row(p, g, row, zr, bi);
                } else {
                    p.idx[o] = 0;
                    const int slot = atomicAdd(p.flag_count, 1);
                    p.flag_list[slot] = make_int2(g, row);
                }
            }
            __syncwarp();
            if (p.dw) scatter_warp(p, g, zr, my_code, lane);
        }
        if (p.loss) {
            qloss = pg_warp_sum_d(qloss);
            if (lane == 0 && qloss != 0.0) atomicAdd(p.loss, qloss);
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 2) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc(tmem_base, 512u);
    }
}

// Exact fp32 full scan of the flagged rows: one warp per row, lanes over codes; identical
// arithmetic to vq_assign_kernel (sequential fmaf over d, (zz - 2 dot) + ee, lowest index on ties).
__global__ void __launch_bounds__(256) vq16_rescore_kernel(const Vq16P p) {
    extern __shared__ float zsm[];                 // [8 warps][D]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* zs = zsm + warp * p.D;
    const int n = *p.flag_count;
    for (int w = blockIdx.x * 8 + warp; w < n; w += gridDim.x * 8) {
        const int2 gr = p.flag_list[w];
        const float* zr = p.z + (long long)gr.x * p.z_gs + (long long)gr.y * p.ldz;
        __syncwarp();
        for (int d = lane; d < p.D; d += 32) zs[d] = zr[d];
        __syncwarp();
        float zz = 0.f;
        for (int d = 0; d < p.D; ++d) zz = fmaf(zs[d], zs[d], zz);
        float best = INFINITY, second = INFINITY;
        int bi = 0x7fffffff;
        const float* eg = p.e + (long long)gr.x * p.e_gs;
        for (int k = lane; k < p.K; k += 32) {
            const float* er = eg + (long long)k * p.lde;
            float acc = 0.f;
            for (int d = 0; d < p.D; ++d) acc = fmaf(zs[d], __ldg(er + d), acc);
            const float dist = (zz - 2.0f * acc) + __ldg(p.ee + (long long)gr.x * p.Kpad + k);
            if (dist < best) { second = best; best = dist; bi = k; }
            else if (dist < second) second = dist;
        }
        float gb = best; int gi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, gb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, gi, o);
            if (ob < gb || (ob == gb && oi < gi)) { gb = ob; gi = oi; }
        }
        if (gi == 0x7fffffff) gi = 0;              // every distance NaN: tf.argmin returns 0
        float cand = (bi == gi) ? second : best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = fminf(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0) {
            const long long o = (long long)gr.x * p.idx_gs + gr.y;
            p.idx[o] = gi;
            if (p.best) p.best[o] = gb;
            if (p.gap) p.gap[o] = cand - gb;
            if (p.dw) scatter_row(p, gr.x, zr, gi);
            if (p.q || p.stb || p.loss) {
                const float part = quantize_row(p, gr.x, gr.y, zr, gi);
                if (p.loss) atomicAdd(p.loss, (double)part);
            }
        }
    }
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int ensure_scratch(pgmvae_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return PGMVAE_OK;
    if (ctx->scratch) {
        PG_CUDA(cudaStreamSynchronize(ctx->stream));
        PG_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    PG_CUDA(cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return PGMVAE_OK;
}

template <int SUB>
int launch16(pgmvae_ctx* ctx, cudaStream_t st, const CUtensorMap& mapE, Vq16P& p) {
    using C = Cfg<SUB>;
    p.tiles_m = (int)pg_cdiv(p.B, C::TMR);
    const size_t fixed = 1024 + (size_t)RING * C::TMR * (2 * sizeof(float4) + sizeof(float2)) + 256;
    const size_t stage_bytes = (size_t)p.kblocks * C::BN * KB_BYTES;
    p.stages = MAX_STAGES;
    while (p.stages > 2 && fixed + p.stages * stage_bytes > ctx->smem_optin) --p.stages;
    const size_t smem = fixed + p.stages * stage_bytes;
    if (smem > ctx->smem_optin) {
        pgmvae_set_error("vq_assign (fp16 tensor core): shared memory %zu exceeds %zu", smem, ctx->smem_optin);
        return PGMVAE_EINVAL;
    }
    static size_t configured[16] = {};          // per device: the attribute is set per device
    if (smem > configured[ctx->device & 15]) {
        PG_CUDA(cudaFuncSetAttribute(vq_assign_f16_kernel<SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15] = smem;
    }
    const int items = p.G * p.tiles_m;
    const int grid = items < ctx->sm_count ? items : ctx->sm_count;
    vq_assign_f16_kernel<SUB><<<grid, C::THREADS, smem, st>>>(mapE, p);
    return PGMVAE_OK;
}

}  // namespace

bool pg_vq_assign_f16_supported(int D, int K) { return K >= 1 && D >= 1 && D + 2 <= 128; }

// assignment (+ optional fused EMA statistics when cnt/dw are given; they are accumulated into)
int pg_vq_assign_f16(pgmvae_ctx* ctx, cudaStream_t st, const float* z, int64_t z_gs, int ldz, const float* e,
                     int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt,
                     float* cnt_opt, int64_t c_gs, float* dw_opt, int64_t dw_gs, int lddw, int G, int B, int D, int K,
                     float* q_opt, __nv_bfloat16* stb_opt, int64_t q_gs, int ldq, double* loss_opt) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    Vq16P p{};
    p.q = q_opt; p.stb = stb_opt; p.q_gs = q_gs; p.ldq = ldq; p.loss = loss_opt;
    p.G = G; p.B = B; p.D = D; p.K = K;
    p.KD = pg_round_up(D + 2, 16);
    p.ksteps = p.KD / 16;
    p.kblocks = (int)pg_cdiv(p.KD, 64);
    // three 128-row sub-tiles (12 epilogue warps, 64-code tiles) when the z operand fits 128 TMEM columns
    // three 128-row sub-tiles (12 epilogue warps, 64-code tiles) when the z operand fits 128 TMEM columns
    int sub = B >= 3 * TM ? 3 : 2;
    if (const char* ev = getenv("PGMVAE_VQ_SUB")) sub = atoi(ev) == 3 ? 3 : 2;
    if (3 * (p.KD / 2) > A_COLS) sub = 2;
    const int BN = sub == 2 ? Cfg<2>::BN : Cfg<3>::BN;
    p.tiles_n = (int)pg_cdiv(K, BN);
    p.Kpad = p.tiles_n * BN;
    const size_t off_ee = 0, off_emax = align256((size_t)G * p.Kpad * 4), off_cnt = off_emax + 256,
                 off_list = off_cnt + 256, off_e16 = align256(off_list + (size_t)G * B * sizeof(int2)),
                 total = off_e16 + (size_t)G * p.Kpad * p.KD * 2;
    PG_TRY(ensure_scratch(ctx, total));
    ctx->vq_cnt_off = off_cnt;
    uint8_t* sc = (uint8_t*)ctx->scratch;
    float* ee = (float*)(sc + off_ee);
    float* emax = (float*)(sc + off_emax);
    int* cnt = (int*)(sc + off_cnt);
    int2* list = (int2*)(sc + off_list);
    __half* e16 = (__half*)(sc + off_e16);
    p.z = z; p.z_gs = z_gs; p.ldz = ldz; p.e = e; p.e_gs = e_gs; p.lde = lde; p.ee = ee; p.emax = emax;
    p.zvec = !((uintptr_t)z & 15) && ldz % 4 == 0 && z_gs % 4 == 0;
    p.qvec = p.zvec && D % 8 == 0 && !((uintptr_t)e & 15) && lde % 4 == 0 && e_gs % 4 == 0 && ldq % 8 == 0 && q_gs % 8 == 0 &&
             !((uintptr_t)q_opt & 15) && !((uintptr_t)stb_opt & 15);
    p.idx = idx; p.idx_gs = idx_gs; p.best = best_opt; p.gap = gap_opt;
    p.cnt = cnt_opt; p.c_gs = c_gs; p.dw = (cnt_opt && dw_opt) ? dw_opt : nullptr; p.dw_gs = dw_gs; p.lddw = lddw;
    p.flag_count = cnt; p.flag_list = list;
    // 2 eps in distance units: fp16 rounds to nearest (2^-11 per operand) -> |d~ - d| <= 2^-9 |z| |e|
    p.margin_scale = 0.00390625f;
    p.margin_abs = 2e-6f;

    PG_CUDA(cudaMemsetAsync(emax, 0, 512, st));       // emax and the flag counter
    PG_KERNEL(ctx, st, "vq_e_prep", (double)G * K * (4.0 * D + 4.0 + 2.0 * p.KD), 2.0 * G * K * D);
    e_prep_kernel<<<(unsigned)pg_cdiv((int64_t)G * p.Kpad, 128), 128, 0, st>>>(e, e_gs, lde, ee, emax, e16, G, K, p.Kpad, D,
                                                                              p.KD);
    PG_LAUNCHED(ctx);

    CUtensorMap mapE;
    PG_TRY(tc::make_map(&mapE, e16, 2, (uint64_t)p.KD, (uint64_t)p.Kpad, (uint64_t)G, (uint64_t)p.KD,
                        (uint64_t)p.Kpad * p.KD, 64, (uint32_t)BN));
    const double fused = p.dw ? 4.0 * G * (double)K * (D + 1.0) : 0.0;
    PG_KERNEL(ctx, st, p.dw ? "vq_assign_ema_tc_f16" : "vq_assign_tc_f16",
              4.0 * ((double)G * B * D + (double)G * K * D + (double)G * B) + fused, 2.0 * G * B * (double)D * K);
    if (sub == 3) PG_TRY(launch16<3>(ctx, st, mapE, p));
    else PG_TRY(launch16<2>(ctx, st, mapE, p));
    PG_LAUNCHED(ctx);

    PG_KERNEL(ctx, st, "vq_rescore_fp32", 0.0, 0.0);
    vq16_rescore_kernel<<<ctx->sm_count * 4, 256, 8 * D * sizeof(float), st>>>(p);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}
