// Kernel (b), tensor-core flavour: VQ assignment as a z . E^T contraction on tcgen05 with the
// |e|^2 correction and the arg-min fused into the TMEM epilogue (reference
// core/quantizer.py:44-47 / :135-138; the [V,B,K] distance tensor is never materialised).
//
// Per CTA (persistent over row tiles of 256 samples of one variable):
//   z tile  [2 x 128 rows, D]  TMA (128-byte swizzle) -> smem once per row tile
//   E tiles [BN codes, D]      TMA -> smem ring (3-4 stages); every tile feeds 2 MMA groups, so
//                              the L2 -> SM traffic per distance is half of a 128-row design
//   acc     2 x [128, BN] fp32 in TMEM, double buffered: MMA of code tile t+1 overlaps the
//                              epilogue of code tile t
//   warp 0 TMA producer | warp 1 MMA issuer (one thread) | warp 2 TMEM allocator |
//   warps 4-11 epilogue, one sample row per thread (tcgen05.ld 32x32b: lane == row)
// Operands: kind::tf32 directly on the fp32 data as it lies in HBM (no copy, no conversion pass).
// The fp16 flavour lives in vq_tc16.cu (single pass, z operand in tensor memory).
//
// Exactness.  Low-precision products only PROPOSE candidates; fp32 decides.  With
// eps = 2^-8 |z| max|e| (tf32, truncation) or 2^-9 |z| max|e| (fp16, round-to-nearest) bounding
// the error of every approximate distance d~, two passes over the code tiles compute
//   pass 1:  m~ = min_k d~_k        pass 2:  C = { k : d~_k <= m~ + 2 eps }
// and the true arg-min is always in C (d~_k* <= d_k* + eps <= d_j + eps <= d~_j + 2 eps for the
// approximate minimiser j).  The candidates (1-2 per row in practice) are re-scored with
// exactly the fp32 arithmetic of the CUDA-core kernel, lowest index on ties, so the indices
// are those of the fp32 path.  Rows with more candidates than the list holds (many identical
// dead codes) go to vq_rescore_kernel, an exact full scan.
#include <stdlib.h>

#include "common.cuh"
#include "ops.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;                       // rows per MMA (TMEM lanes)
constexpr int SUB = 2;                        // 128-row sub-tiles per CTA row tile
constexpr int TMR = TM * SUB;
constexpr int KB_BYTES = 128;                 // one k-block = one 128-byte swizzle row
constexpr int A_TILE_BYTES = TM * KB_BYTES;   // 16 KB per (sub-tile, k-block)
constexpr int EPI_WARPS = 4 * SUB;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int CAND_CAP = 8;                   // candidate codes kept per row
constexpr int MAX_STAGES = 4;

struct VqTcParams {
    int G, B, D, K;
    int kblocks, ksteps, BN, tiles_m, tiles_n, tmem_cols, Kpad, stages;
    const float* z; long long z_gs; int ldz;
    const float* e; long long e_gs; int lde;
    const float* ee;        // [G][Kpad] squared norms, +inf padding
    const float* emax;      // [1] max_k |e_k| over all groups
    int32_t* idx; long long idx_gs;
    float* best; float* gap;
    int* flag_count; int2* flag_list;
    float margin_scale, margin_abs;
    int dbg;     // PGMVAE_VQ_DBG timing experiments: bit0 = skip the reduction, bit1 = skip the TMEM loads
};

// ee[g][k] = |e_k|^2 for k < K, +inf for the padding codes k in [K, Kpad); emax = max_k |e_k|
__global__ void enorm_kernel(const float* __restrict__ e, long long e_gs, int lde, float* __restrict__ ee,
                             float* __restrict__ emax, int G, int K, int Kpad, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f;
    if (i < (long long)G * Kpad) {
        const int g = (int)(i / Kpad), k = (int)(i - (long long)g * Kpad);
        if (k < K) {
            const float* row = e + (long long)g * e_gs + (long long)k * lde;
            for (int d = 0; d < D; ++d) s = fmaf(row[d], row[d], s);
            ee[i] = s;
        } else {
            ee[i] = INFINITY;
        }
    }
    float m = sqrtf(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(emax), __float_as_int(m));   // m >= 0
}

// minimum of d' = ee - 2 z.e over one 32-column chunk (four independent chains for ILP)
__device__ __forceinline__ float chunk_min(const float (&v)[32], const float* __restrict__ see) {
    float m0 = INFINITY, m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 e4 = *reinterpret_cast<const float4*>(see + j);
        m0 = fminf(m0, fmaf(-2.0f, v[j + 0], e4.x));
        m1 = fminf(m1, fmaf(-2.0f, v[j + 1], e4.y));
        m2 = fminf(m2, fmaf(-2.0f, v[j + 2], e4.z));
        m3 = fminf(m3, fmaf(-2.0f, v[j + 3], e4.w));
    }
    return fminf(fminf(m0, m1), fminf(m2, m3));
}

__global__ void __launch_bounds__(128 + EPI_THREADS, 1)
vq_assign_tc_kernel(const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapE,
                    const VqTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a shared-space pointer
    constexpr int KB_ELEMS = 32;                      // fp32 elements per 128-byte k-block
    const int b_tile_bytes = p.BN * KB_BYTES;
    uint8_t* sA = smem;                                               // [SUB][kblocks][16 KB]
    uint8_t* sB = sA + (size_t)SUB * p.kblocks * A_TILE_BYTES;        // [stages][kblocks][BN*128]
    float* sEE = reinterpret_cast<float*>(sB + (size_t)p.stages * p.kblocks * b_tile_bytes);   // [Kpad]
    int* cand = reinterpret_cast<int*>(sEE + p.Kpad);                 // [TMR][CAND_CAP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cand + TMR * CAND_CAP);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* ee_full = bars + 2;
    uint64_t* e_done = bars + 3;
    uint64_t* acc_full = bars + 4;                 // [2]
    uint64_t* acc_empty = bars + 6;                // [2]
    uint64_t* b_full = bars + 8;                   // [MAX_STAGES]
    uint64_t* b_empty = bars + 8 + MAX_STAGES;     // [MAX_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * MAX_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = p.G * p.tiles_m;
    const int tiles_item = 2 * p.tiles_n;             // both passes

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&mapZ);
        tc::tma_prefetch_desc(&mapE);
    }
    if (warp == 1 && lane == 0) {
        tc::mbar_init(a_full, 1);
        tc::mbar_init(a_empty, 1);
        tc::mbar_init(ee_full, 1);
        tc::mbar_init(e_done, EPI_WARPS);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], EPI_WARPS);      // one arrival per epilogue warp
        }
        for (int s = 0; s < MAX_STAGES; ++s) {
            tc::mbar_init(&b_full[s], 1);
            tc::mbar_init(&b_empty[s], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        tc::tmem_relinquish();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        uint32_t item_n = 0, s = 0, ph = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_n) {
            const int g = item / p.tiles_m, mt = item - g * p.tiles_m;
            tc::mbar_wait(a_empty, (item_n & 1) ^ 1);
            if (tc::elect_one()) {
                tc::mbar_arrive_expect_tx(a_full, (uint32_t)(SUB * p.kblocks * A_TILE_BYTES));
                for (int sub = 0; sub < SUB; ++sub)
                    for (int kb = 0; kb < p.kblocks; ++kb)
                        tc::tma_load_3d(sA + (size_t)(sub * p.kblocks + kb) * A_TILE_BYTES, &mapZ, a_full,
                                        kb * KB_ELEMS, mt * TMR + sub * TM, g);
            }
            __syncwarp();
            // |e|^2 of this group's codes: reloaded once the epilogue has left the previous item
            tc::mbar_wait(e_done, (item_n & 1) ^ 1);
            if (tc::elect_one()) {
                tc::mbar_arrive_expect_tx(ee_full, (uint32_t)(p.Kpad * 4));
                tc::bulk_load_1d(sEE, p.ee + (long long)g * p.Kpad, (uint32_t)(p.Kpad * 4), ee_full);
            }
            __syncwarp();
            for (int tt = 0; tt < tiles_item; ++tt) {
                const int t = tt < p.tiles_n ? tt : tt - p.tiles_n;
                tc::mbar_wait(&b_empty[s], ph ^ 1);
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&b_full[s], (uint32_t)(p.kblocks * b_tile_bytes));
                    for (int kb = 0; kb < p.kblocks; ++kb)
                        tc::tma_load_3d(sB + ((size_t)s * p.kblocks + kb) * b_tile_bytes, &mapE, &b_full[s],
                                        kb * KB_ELEMS, t * p.BN, g);
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        const uint32_t idesc = tc::make_idesc(2, TM, p.BN, 0, 0);
        // Descriptors differ only in the 14-bit start-address field (units of 16 bytes).
        const uint64_t descA0 = tc::make_smem_desc(tc::smem_u32(sA), 16, 1024);
        const uint64_t descB0 = tc::make_smem_desc(tc::smem_u32(sB), 16, 1024);
        const uint32_t sub_stride = (uint32_t)(p.kblocks * A_TILE_BYTES) >> 4;
        const uint32_t stage_stride = (uint32_t)(p.kblocks * b_tile_bytes) >> 4;
        uint32_t it = 0, item_n = 0, s = 0, ph = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_n) {
            tc::mbar_wait(a_full, item_n & 1);
            for (int tt = 0; tt < tiles_item; ++tt, ++it) {
                const uint32_t ab = it & 1, aph = (it >> 1) & 1;
                tc::mbar_wait(&b_full[s], ph);
                tc::mbar_wait(&acc_empty[ab], aph ^ 1);
                tc::fence_after_thread_sync();
                if (tc::elect_one()) {
                    const uint64_t descB = descB0 + (uint64_t)(s * stage_stride);
#pragma unroll
                    for (int sub = 0; sub < SUB; ++sub) {
                        const uint32_t d_tmem = tmem_base + (ab * SUB + sub) * p.BN;
                        const uint64_t descA = descA0 + (uint64_t)(sub * sub_stride);
                        for (int ks = 0; ks < p.ksteps; ++ks) {
                            const uint32_t offA = (uint32_t)((ks >> 2) * A_TILE_BYTES + (ks & 3) * 32) >> 4;
                            const uint32_t offB = (uint32_t)((ks >> 2) * b_tile_bytes + (ks & 3) * 32) >> 4;
                            tc::mma_tf32(d_tmem, descA + offA, descB + offB, idesc, ks > 0 ? 1u : 0u);
                        }
                    }
                    tc::mma_commit(&b_empty[s]);      // smem stage free once these MMAs have read it
                    tc::mma_commit(&acc_full[ab]);    // accumulators ready for the epilogue
                    if (tt == tiles_item - 1) tc::mma_commit(a_empty);    // z tile may be overwritten
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
        int* mycand = cand + r * CAND_CAP;
        uint32_t it = 0, item_n = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_n) {
            const int g = item / p.tiles_m, mt = item - g * p.tiles_m;
            const int row = mt * TMR + r;
            const bool valid = row < p.B;
            const float* zr = p.z + (long long)g * p.z_gs + (long long)(valid ? row : 0) * p.ldz;
            float zz = 0.f;
            for (int d = 0; d < p.D; ++d) zz = fmaf(zr[d], zr[d], zz);
            const float znorm = sqrtf(zz);
            // 2 eps: operand rounding (relative to |z| |e|) + fp32 accumulation / final-sum rounding, which is
            // relative to the magnitudes involved -- NOT an absolute constant: at initialisation codes and
            // latents are ~1e-2 and an absolute 2e-5 made every code a candidate (all rows to the full scan)
            const float margin = p.margin_scale * znorm * emax + p.margin_abs * (znorm + emax) * (znorm + emax);
            float thr = INFINITY, m = INFINITY;
            int ncand = 0;
            tc::mbar_wait(ee_full, item_n & 1);
            // Flat stream of 32-column chunks over (pass, code tile): the TMEM load of chunk n+1 -- also across
            // tile boundaries -- is in flight while chunk n is reduced, and an accumulator buffer is handed
            // back to the MMA warp as soon as its last chunk sits in registers.
            const int cpt = p.BN >> 5;                          // chunks per tile (padding codes carry +inf)
            const int nchunks = 2 * p.tiles_n * cpt, npass1 = p.tiles_n * cpt;
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + sub * p.BN;
            int ld_tile = 0, ld_c = 0;                           // next chunk to load
            int rt_c = 0;  uint32_t rt_it = it;                  // next chunk to retire
            auto issue = [&](float (&buf)[32]) {
                const uint32_t git = it + ld_tile, ab = git & 1;
                if (ld_c == 0) {
                    tc::mbar_wait(&acc_full[ab], (git >> 1) & 1);
                    tc::fence_after_thread_sync();
                }
                if (!(p.dbg & 2)) tc::tmem_ld_32x32(lane_base + ab * SUB * p.BN + ld_c * 32, buf);
                if (++ld_c == cpt) { ld_c = 0; ++ld_tile; }
            };
            auto retire = [&]() {                                // after wait::ld: the chunk is in registers
                if (++rt_c == cpt) {
                    rt_c = 0;
                    tc::fence_before_thread_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&acc_empty[rt_it & 1]);
                    ++rt_it;
                }
            };
            int cs_tile = 0, cs_c = 0;                           // chunk being consumed
            auto consume = [&](const float (&v)[32], int n) {
                if (n == npass1) thr = m + margin;
                if (p.dbg & 1) { if (++cs_c == cpt) { cs_c = 0; ++cs_tile; } return; }
                const int t = cs_tile < p.tiles_n ? cs_tile : cs_tile - p.tiles_n;
                const float* see = sEE + t * p.BN + cs_c * 32;
                const float cm = chunk_min(v, see);
                if (n < npass1) {
                    m = fminf(m, cm);
                } else if (cm <= thr && !(p.dbg & 19)) {
                    const int kbase = t * p.BN + cs_c * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (fmaf(-2.0f, v[j], see[j]) <= thr) {
                            if (ncand < CAND_CAP) mycand[ncand] = kbase + j;
                            ++ncand;
                        }
                    }
                }
                if (++cs_c == cpt) { cs_c = 0; ++cs_tile; }
            };
            float va[32], vb[32];
            if (p.dbg & 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) { va[j] = (float)(j + lane) * 1e-3f; vb[j] = va[j] + 0.5f; }
            }
            issue(va);
            for (int n = 0; n < nchunks; n += 2) {               // nchunks is even
                tc::tmem_ld_wait(va);
                retire();
                issue(vb);
                consume(va, n);
                tc::tmem_ld_wait(vb);
                retire();
                if (n + 2 < nchunks) issue(va);
                consume(vb, n + 1);
            }
            it += 2 * p.tiles_n;
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(e_done);    // sEE may be replaced
            if (valid) {
                const long long o = (long long)g * p.idx_gs + row;
                if (ncand >= 1 && ncand <= CAND_CAP) {
                    // exact fp32 distances of the candidates: same arithmetic as the CUDA-core kernel
                    float best = INFINITY, second = INFINITY;
                    int bi = 0x7fffffff;
                    const float* eg = p.e + (long long)g * p.e_gs;
                    for (int i = 0; i < ncand; ++i) {
                        const int k = mycand[i];
                        const float* er = eg + (long long)k * p.lde;
                        float acc = 0.f;
                        for (int d = 0; d < p.D; ++d) acc = fmaf(zr[d], __ldg(er + d), acc);
                        const float dist = (zz - 2.0f * acc) + __ldg(p.ee + (long long)g * p.Kpad + k);
                        if (dist < best || (dist == best && k < bi)) { second = best; best = dist; bi = k; }
                        else if (dist < second) second = dist;
                    }
                    p.idx[o] = bi;
                    if (p.best) p.best[o] = best;
                    // exact gap when a runner-up lies inside the error band, otherwise a lower bound
                    if (p.gap) p.gap[o] = ncand > 1 ? second - best : margin;
                } else if (!(p.dbg & 19)) {
                    // more tied / near-tied codes than the list holds (e.g. many identical dead codes), or
                    // non-finite low-precision distances: hand the row to the exact full-scan kernel
                    p.idx[o] = 0;
                    const int slot = atomicAdd(p.flag_count, 1);
                    p.flag_list[slot] = make_int2(g, row);
                }
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 2) {
        tc::fence_after_thread_sync();
        tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// Exact fp32 re-scoring of the flagged rows: one warp per row, lanes over codes; identical
// arithmetic to vq_assign_kernel (sequential fmaf over d, (zz - 2 dot) + ee, lowest index on ties).
__global__ void __launch_bounds__(256) vq_rescore_kernel(const float* __restrict__ z, long long z_gs, int ldz,
                                                         const float* __restrict__ e, long long e_gs, int lde,
                                                         const float* __restrict__ ee, int32_t* __restrict__ idx,
                                                         long long idx_gs, float* __restrict__ best_out,
                                                         float* __restrict__ gap_out, const int* __restrict__ flag_count,
                                                         const int2* __restrict__ flag_list, int D, int K, int Kpad) {
    extern __shared__ float zsm[];                 // [8 warps][D]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* zs = zsm + warp * D;
    const int n = *flag_count;
    for (int w = blockIdx.x * 8 + warp; w < n; w += gridDim.x * 8) {
        const int2 gr = flag_list[w];
        const float* zr = z + (long long)gr.x * z_gs + (long long)gr.y * ldz;
        __syncwarp();
        for (int d = lane; d < D; d += 32) zs[d] = zr[d];
        __syncwarp();
        float zz = 0.f;
        for (int d = 0; d < D; ++d) zz = fmaf(zs[d], zs[d], zz);
        float best = INFINITY, second = INFINITY;
        int bi = 0x7fffffff;
        const float* eg = e + (long long)gr.x * e_gs;
        for (int k = lane; k < K; k += 32) {
            const float* er = eg + (long long)k * lde;
            float acc = 0.f;
            for (int d = 0; d < D; ++d) acc = fmaf(zs[d], __ldg(er + d), acc);
            const float dist = (zz - 2.0f * acc) + __ldg(ee + (long long)gr.x * Kpad + k);
            if (dist < best) { second = best; best = dist; bi = k; }
            else if (dist < second) second = dist;
        }
        // warp arg-min, lowest index on ties
        float gb = best; int gi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, gb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, gi, o);
            if (ob < gb || (ob == gb && oi < gi)) { gb = ob; gi = oi; }
        }
        float cand = (bi == gi) ? second : best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = fminf(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0) {
            const long long o = (long long)gr.x * idx_gs + gr.y;
            idx[o] = gi;
            if (best_out) best_out[o] = gb;
            if (gap_out) gap_out[o] = cand - gb;
        }
    }
}

}  // namespace

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// scratch owned by the context (grown on demand)
static int ensure_scratch(pgmvae_ctx* ctx, size_t bytes) {
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

bool pg_vq_assign_tc_supported(int prec, int D, int K, int ldz, int lde, const float* z, const float* e, int64_t z_gs,
                               int64_t e_gs) {
    if (prec != PGMVAE_PREC_TF32) return false;
    if (K < 1 || K > 8192) return false;                  // |e|^2 of one group is staged in shared memory
    if (D > 64) return false;                             // two 128-byte k-blocks of fp32
    return !(((uintptr_t)z & 15) || ((uintptr_t)e & 15) || ldz % 4 || lde % 4 || z_gs % 4 || e_gs % 4);
}

static int launch_vq_tc(pgmvae_ctx* ctx, cudaStream_t st, const CUtensorMap& mapZ, const CUtensorMap& mapE,
                        const VqTcParams& p, size_t smem, int grid) {
    static size_t configured[16] = {};          // per device: the attribute is set per device
    if (smem > configured[ctx->device & 15]) {
        PG_CUDA(cudaFuncSetAttribute(vq_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15] = smem;
    }
    vq_assign_tc_kernel<<<grid, 128 + EPI_THREADS, smem, st>>>(mapZ, mapE, p);
    return PGMVAE_OK;
}

int pg_vq_assign_tc(pgmvae_ctx* ctx, cudaStream_t st, int prec, const float* z, int64_t z_gs, int ldz, const float* e,
                    int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt, int G, int B,
                    int D, int K) {
    (void)prec;                                               // tf32 only; the fp16 path is pg_vq_assign_f16
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    VqTcParams p{};
    p.G = G; p.B = B; p.D = D; p.K = K;
    p.ksteps = (int)pg_cdiv((int64_t)D * 4, 32);              // one MMA consumes 32 bytes of K
    p.kblocks = (int)pg_cdiv((int64_t)D * 4, KB_BYTES);
    p.BN = 128;
    if (K < p.BN) p.BN = pg_round_up(K, 32);                  // UMMA N % 16 == 0; the epilogue reads 32-column chunks
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * SUB * p.BN) p.tmem_cols <<= 1;
    p.tiles_m = (int)pg_cdiv(B, TMR);
    p.tiles_n = (int)pg_cdiv(K, p.BN);
    p.Kpad = p.tiles_n * p.BN;
    const size_t off_ee = 0, off_emax = align256((size_t)G * p.Kpad * 4), off_cnt = off_emax + 256,
                 off_list = off_cnt + 256, total = align256(off_list + (size_t)G * B * sizeof(int2));
    PG_TRY(ensure_scratch(ctx, total));
    ctx->vq_cnt_off = off_cnt;
    uint8_t* sc = (uint8_t*)ctx->scratch;
    float* ee = (float*)(sc + off_ee);
    float* emax = (float*)(sc + off_emax);
    int* cnt = (int*)(sc + off_cnt);
    int2* list = (int2*)(sc + off_list);
    p.z = z; p.z_gs = z_gs; p.ldz = ldz; p.e = e; p.e_gs = e_gs; p.lde = lde; p.ee = ee; p.emax = emax;
    p.idx = idx; p.idx_gs = idx_gs; p.best = best_opt; p.gap = gap_opt;
    p.flag_count = cnt; p.flag_list = list;
    // 2 eps: tf32 operands are read by truncation (2^-10 per operand); see header
    p.margin_scale = 0.0078125f;
    p.margin_abs = 2e-6f;                              // ~32 ulp of the largest term
    p.dbg = getenv("PGMVAE_VQ_DBG") ? atoi(getenv("PGMVAE_VQ_DBG")) : 0;

    PG_CUDA(cudaMemsetAsync(emax, 0, 512, st));       // emax and the flag counter
    PG_KERNEL(ctx, st, "vq_enorm", 4.0 * G * K * (D + 1.0), 2.0 * G * K * D);
    enorm_kernel<<<(unsigned)pg_cdiv((int64_t)G * p.Kpad, 256), 256, 0, st>>>(e, e_gs, lde, ee, emax, G, K, p.Kpad, D);
    PG_LAUNCHED(ctx);

    CUtensorMap mapZ, mapE;
    PG_TRY(tc::make_map(&mapZ, z, 4, (uint64_t)D, (uint64_t)B, (uint64_t)G, (uint64_t)ldz, (uint64_t)z_gs, 32, TM));
    PG_TRY(tc::make_map(&mapE, e, 4, (uint64_t)D, (uint64_t)K, (uint64_t)G, (uint64_t)lde, (uint64_t)e_gs, 32,
                        (uint32_t)p.BN));
    // smem: A [SUB][kblocks] | B ring | ee | candidates | barriers
    const size_t fixed = 1024 + (size_t)SUB * p.kblocks * A_TILE_BYTES + (size_t)p.Kpad * 4 +
                         (size_t)TMR * CAND_CAP * 4 + 256;
    const size_t stage_bytes = (size_t)p.kblocks * p.BN * KB_BYTES;
    p.stages = MAX_STAGES;
    while (p.stages > 2 && fixed + p.stages * stage_bytes > ctx->smem_optin) --p.stages;
    const size_t smem = fixed + p.stages * stage_bytes;
    if (smem > ctx->smem_optin) {
        pgmvae_set_error("vq_assign (tensor core): shared memory %zu exceeds %zu", smem, ctx->smem_optin);
        return PGMVAE_EINVAL;
    }
    const int items = G * p.tiles_m;
    const int grid = items < ctx->sm_count ? items : ctx->sm_count;
    PG_KERNEL(ctx, st, "vq_assign_tc_tf32", 4.0 * ((double)G * B * D + (double)G * K * D + (double)G * B),
              2.0 * G * B * (double)D * K);
    PG_TRY(launch_vq_tc(ctx, st, mapZ, mapE, p, smem, grid));
    PG_LAUNCHED(ctx);

    PG_KERNEL(ctx, st, "vq_rescore_fp32", 0.0, 0.0);
    vq_rescore_kernel<<<ctx->sm_count * 4, 256, 8 * D * sizeof(float), st>>>(z, z_gs, ldz, e, e_gs, lde, ee, idx, idx_gs,
                                                                             best_opt, gap_opt, cnt, list, D, K, p.Kpad);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

// number of rows the last pg_vq_assign_tc call on this context handed to the exact full scan (synchronises)
int pg_vq_assign_tc_last_flagged(pgmvae_ctx* ctx, int G, int K, int* out) {
    (void)G; (void)K;
    if (!ctx->scratch) { *out = 0; return PGMVAE_OK; }
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    PG_CUDA(cudaMemcpy(out, (uint8_t*)ctx->scratch + ctx->vq_cnt_off, sizeof(int), cudaMemcpyDeviceToHost));
    return PGMVAE_OK;
}
