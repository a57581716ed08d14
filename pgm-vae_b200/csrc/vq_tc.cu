// Kernel (b), tensor-core flavour: VQ assignment as a z . E^T contraction on tcgen05 with the
// |e|^2 correction and the arg-min fused into the TMEM epilogue (reference
// core/quantizer.py:44-47 / :135-138; the [V,B,K] distance tensor is never materialised).
//
//   z tile  [128 rows, D]   TMA (128-byte swizzle) -> smem, loaded once per row tile
//   E tiles [BN codes, D]   TMA -> smem, 2-stage ring
//   acc     [128, BN] fp32  TMEM, 2 buffers (2*BN <= 512 columns): the MMA of code tile t+1
//                           overlaps the epilogue of code tile t
//   warp 0 = TMA producer, warp 1 = MMA issuer (one thread, kind::tf32 on the fp32 data as it
//   lies in HBM), warp 2 = TMEM allocator, warps 4-7 = epilogue (one row per thread).
//
// Exactness: tf32 products carry a relative error < 2^-9, so the arg-min of a row is trusted
// only if its runner-up is further away than the rigorous bound 2^-7 |z| max_k|e_k|.  Other
// rows are appended to a list and re-scored by vq_rescore_kernel with exactly the fp32
// arithmetic of the CUDA-core kernel (lowest index on ties), so that the indices are those of
// the fp32 path everywhere.
#include "common.cuh"
#include "ops.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;
constexpr int KB_BYTES = 128;                 // one k-block = 32 fp32 = one swizzle row
constexpr int A_TILE_BYTES = TM * KB_BYTES;   // 16 KB per k-block

struct VqTcParams {
    int G, B, D, K;
    int kblocks, ksteps, BN, tiles_m, tiles_n, tmem_cols;
    const float* z; long long z_gs; int ldz;
    const float* ee;        // [G][K] squared norms
    const float* emax;      // [1] max_k |e_k| over all groups
    int32_t* idx; long long idx_gs;
    float* best; float* gap;
    int* flag_count; int2* flag_list;
    float margin_scale, margin_abs;
};

__global__ void enorm_kernel(const float* __restrict__ e, long long e_gs, int lde, float* __restrict__ ee,
                             float* __restrict__ emax, int G, int K, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f;
    if (i < (long long)G * K) {
        const int g = (int)(i / K), k = (int)(i - (long long)g * K);
        const float* row = e + (long long)g * e_gs + (long long)k * lde;
        for (int d = 0; d < D; ++d) s = fmaf(row[d], row[d], s);
        ee[i] = s;
    }
    float m = sqrtf(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(emax), __float_as_int(m));   // m >= 0
}

__global__ void __launch_bounds__(256, 1)
vq_assign_tc_kernel(const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapE,
                    const VqTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_tile_bytes = p.BN * KB_BYTES;
    uint8_t* sA = smem;                                               // [kblocks][16 KB]
    uint8_t* sB = sA + (size_t)p.kblocks * A_TILE_BYTES;              // [2][kblocks][BN*128]
    float* sEE = reinterpret_cast<float*>(sB + (size_t)2 * p.kblocks * b_tile_bytes);   // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sEE + 2 * p.BN);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;       // [2]
    uint64_t* b_empty = bars + 4;      // [2]
    uint64_t* acc_full = bars + 6;     // [2]
    uint64_t* acc_empty = bars + 8;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = p.G * p.tiles_m;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&mapZ);
        tc::tma_prefetch_desc(&mapE);
    }
    if (warp == 1 && lane == 0) {
        tc::mbar_init(a_full, 1);
        tc::mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&b_full[s], 1);
            tc::mbar_init(&b_empty[s], 1);
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], 128);
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0, item_n = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_n) {
                const int g = item / p.tiles_m, mt = item - g * p.tiles_m;
                tc::mbar_wait(a_empty, (item_n & 1) ^ 1);
                tc::mbar_arrive_expect_tx(a_full, (uint32_t)(p.kblocks * A_TILE_BYTES));
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tc::tma_load_3d(sA + (size_t)kb * A_TILE_BYTES, &mapZ, a_full, kb * 32, mt * TM, g);
                for (int t = 0; t < p.tiles_n; ++t, ++it) {
                    const uint32_t s = it & 1, ph = (it >> 1) & 1;
                    tc::mbar_wait(&b_empty[s], ph ^ 1);
                    tc::mbar_arrive_expect_tx(&b_full[s], (uint32_t)(p.kblocks * b_tile_bytes));
                    for (int kb = 0; kb < p.kblocks; ++kb)
                        tc::tma_load_3d(sB + ((size_t)s * p.kblocks + kb) * b_tile_bytes, &mapE, &b_full[s], kb * 32,
                                        t * p.BN, g);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc(2, TM, p.BN, 0, 0);
            uint32_t it = 0, item_n = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_n) {
                tc::mbar_wait(a_full, item_n & 1);
                tc::fence_after_thread_sync();
                for (int t = 0; t < p.tiles_n; ++t, ++it) {
                    const uint32_t s = it & 1, ph = (it >> 1) & 1;
                    tc::mbar_wait(&b_full[s], ph);
                    tc::mbar_wait(&acc_empty[s], ph ^ 1);
                    tc::fence_after_thread_sync();
                    const uint32_t d_tmem = tmem_base + s * p.BN;
                    int kstep = 0;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        const uint32_t a_addr = tc::smem_u32(sA + (size_t)kb * A_TILE_BYTES);
                        const uint32_t b_addr = tc::smem_u32(sB + ((size_t)s * p.kblocks + kb) * b_tile_bytes);
                        for (int k4 = 0; k4 < 4 && kstep < p.ksteps; ++k4, ++kstep) {
                            const uint64_t da = tc::make_smem_desc(a_addr + k4 * 32, 16, 1024);
                            const uint64_t db = tc::make_smem_desc(b_addr + k4 * 32, 16, 1024);
                            tc::mma_tf32(d_tmem, da, db, idesc, kstep > 0 ? 1u : 0u);
                        }
                    }
                    tc::mma_commit(&b_empty[s]);      // smem stage free once these MMAs have read it
                    tc::mma_commit(&acc_full[s]);     // accumulator ready for the epilogue
                }
                tc::mma_commit(a_empty);              // z tile may be overwritten
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: one row per thread =====================
        const int q = warp - 4;                        // TMEM lane quarter == warp % 4
        const int r = q * 32 + lane;
        const float emax = *p.emax;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int g = item / p.tiles_m, mt = item - g * p.tiles_m;
            const int row = mt * TM + r;
            const bool valid = row < p.B;
            float zz = 0.f;
            if (valid) {
                const float* zr = p.z + (long long)g * p.z_gs + (long long)row * p.ldz;
                for (int d = 0; d < p.D; ++d) zz = fmaf(zr[d], zr[d], zz);
            }
            float best = INFINITY, second = INFINITY;
            int bi = 0;
            for (int t = 0; t < p.tiles_n; ++t, ++it) {
                const uint32_t s = it & 1, ph = (it >> 1) & 1;
                float* see = sEE + s * p.BN;
                for (int j = r; j < p.BN; j += 128) {
                    const int k = t * p.BN + j;
                    see[j] = k < p.K ? __ldg(p.ee + (long long)g * p.K + k) : INFINITY;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                tc::mbar_wait(&acc_full[s], ph);
                tc::fence_after_thread_sync();
                const int ncol = min(p.BN, p.K - t * p.BN);
                for (int c = 0; c < ncol; c += 32) {
                    float v[32];
                    tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + s * p.BN + c, v);
                    tc::tmem_ld_wait();
                    const int kbase = t * p.BN + c;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        // reference association: (|z|^2 - 2 z.e) + |e|^2 ; padding codes carry +inf
                        const float dist = (zz - 2.0f * v[j]) + see[c + j];
                        if (dist < best) { second = best; best = dist; bi = kbase + j; }
                        else if (dist < second) second = dist;
                    }
                }
                tc::fence_before_thread_sync();
                tc::mbar_arrive(&acc_empty[s]);
            }
            if (valid) {
                const long long o = (long long)g * p.idx_gs + row;
                p.idx[o] = bi;
                if (p.best) p.best[o] = best;
                if (p.gap) p.gap[o] = second - best;
                const float margin = p.margin_scale * sqrtf(zz) * emax + p.margin_abs;
                if (!(second - best > margin)) {
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
                                                         const int2* __restrict__ flag_list, int D, int K) {
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
            const float dist = (zz - 2.0f * acc) + __ldg(ee + (long long)gr.x * K + k);
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

// scratch owned by the context (grown on demand): ee [G*K] | emax | flag counter | flag list [G*B]
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

bool pg_vq_assign_tc_supported(int D, int K, int ldz, int lde, const float* z, const float* e, int64_t z_gs,
                               int64_t e_gs) {
    if (D > 128 || K < 1) return false;
    if (((uintptr_t)z & 15) || ((uintptr_t)e & 15) || ldz % 4 || lde % 4 || z_gs % 4 || e_gs % 4) return false;
    return true;
}

int pg_vq_assign_tc(pgmvae_ctx* ctx, cudaStream_t st, const float* z, int64_t z_gs, int ldz, const float* e,
                    int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt, int G, int B,
                    int D, int K) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    VqTcParams p{};
    p.G = G; p.B = B; p.D = D; p.K = K;
    p.ksteps = (int)pg_cdiv(D, 8);
    p.kblocks = (int)pg_cdiv(D, 32);
    p.BN = p.kblocks <= 2 ? 256 : 128;
    if (K < p.BN) p.BN = pg_round_up(K, 32);          // UMMA N: multiple of 16; the epilogue reads 32 columns at a time
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.BN) p.tmem_cols <<= 1;
    p.tiles_m = (int)pg_cdiv(B, TM);
    p.tiles_n = (int)pg_cdiv(K, p.BN);
    const size_t off_ee = 0, off_emax = align256((size_t)G * K * 4), off_cnt = off_emax + 256, off_list = off_cnt + 256;
    PG_TRY(ensure_scratch(ctx, off_list + (size_t)G * B * sizeof(int2)));
    uint8_t* sc = (uint8_t*)ctx->scratch;
    float* ee = (float*)(sc + off_ee);
    float* emax = (float*)(sc + off_emax);
    int* cnt = (int*)(sc + off_cnt);
    int2* list = (int2*)(sc + off_list);
    p.z = z; p.z_gs = z_gs; p.ldz = ldz; p.ee = ee; p.emax = emax;
    p.idx = idx; p.idx_gs = idx_gs; p.best = best_opt; p.gap = gap_opt;
    p.flag_count = cnt; p.flag_list = list;
    p.margin_scale = 0.0078125f;    // 2^-7: rigorous tf32 truncation bound, see header comment
    p.margin_abs = 2e-5f;

    PG_CUDA(cudaMemsetAsync(emax, 0, 512, st));       // emax and the flag counter
    PG_KERNEL(ctx, st, "vq_enorm", 4.0 * G * K * (D + 1.0), 2.0 * G * K * D);
    enorm_kernel<<<(unsigned)pg_cdiv((int64_t)G * K, 256), 256, 0, st>>>(e, e_gs, lde, ee, emax, G, K, D);
    PG_LAUNCHED(ctx);

    CUtensorMap mapZ, mapE;
    PG_TRY(tc::make_map_f32(&mapZ, z, (uint64_t)D, (uint64_t)B, (uint64_t)G, (uint64_t)ldz, (uint64_t)z_gs, 32, TM));
    PG_TRY(tc::make_map_f32(&mapE, e, (uint64_t)D, (uint64_t)K, (uint64_t)G, (uint64_t)lde, (uint64_t)e_gs, 32,
                            (uint32_t)p.BN));
    const size_t smem = 1024 + (size_t)p.kblocks * A_TILE_BYTES + (size_t)2 * p.kblocks * p.BN * KB_BYTES +
                        (size_t)2 * p.BN * 4 + 128;
    static size_t configured = 0;
    if (smem > configured) {
        PG_CUDA(cudaFuncSetAttribute(vq_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int items = G * p.tiles_m;
    const int grid = items < ctx->sm_count ? items : ctx->sm_count;
    PG_KERNEL(ctx, st, "vq_assign_tc", 4.0 * ((double)G * B * D + (double)G * K * D + (double)G * B),
              2.0 * G * B * (double)D * K);
    vq_assign_tc_kernel<<<grid, 256, smem, st>>>(mapZ, mapE, p);
    PG_LAUNCHED(ctx);

    PG_KERNEL(ctx, st, "vq_rescore_fp32", 0.0, 0.0);
    vq_rescore_kernel<<<ctx->sm_count * 4, 256, 8 * D * sizeof(float), st>>>(z, z_gs, ldz, e, e_gs, lde, ee, idx, idx_gs,
                                                                             best_opt, gap_opt, cnt, list, D, K);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

// number of rows the last pg_vq_assign_tc call on this context re-scored in fp32 (synchronises)
int pg_vq_assign_tc_last_flagged(pgmvae_ctx* ctx, int G, int K, int* out) {
    if (!ctx->scratch) { *out = 0; return PGMVAE_OK; }
    const size_t off_cnt = align256((size_t)G * K * 4) + 256;
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    PG_CUDA(cudaMemcpy(out, (uint8_t*)ctx->scratch + off_cnt, sizeof(int), cudaMemcpyDeviceToHost));
    return PGMVAE_OK;
}
