// Kernel (a), exact-fp32 flavour: grouped (one GEMM per variable) CUDA-core GEMM with the
// fused epilogues of the packed dense layer.  Replaces tf.matmul + bias + activation of
// FatDense.call (reference core/dense.py:99-111) and its autodiff (run.py:62).
//
// One template serves forward, dgrad and wgrad by operand layout:
//   forward : C[B,out] = A[B,in](k-contig)  * B[in,out](n-contig)    + bias, activation
//   dgrad   : C[B,in]  = A[B,out](k-contig) * B[in,out] read as [n,k] (k-contig), * act'(h)
//   wgrad   : C[in,out]= A[B,in] read as [k,m](m-contig) * B[B,out](n-contig), split over B
// 128x64x16 tiles, 256 threads, 8x4 register micro-tile, register-prefetch double buffering.
#include "common.cuh"
#include "ops.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int AS_LD = BM + 4, BS_LD = BN + 4;

enum { EPI_BIAS_ACT = 0, EPI_SIGMOID_MSE = 1, EPI_DGRAD = 2, EPI_WGRAD = 3 };

struct GemmP {
    const float* A; long long a_gs; int lda;
    const float* B; long long b_gs; int ldb;
    float* C; long long c_gs; int ldc;
    int M, N, K;
    int vecA, vecB, vecC;
    int S, kchunk;                    // split of the reduction dimension (wgrad)
    // EPI_BIAS_ACT / EPI_SIGMOID_MSE
    const float* bias; long long bias_gs; int act;
    // EPI_SIGMOID_MSE: aux = y (shared), EPI_DGRAD: aux = h_in
    const float* aux; long long aux_gs; int ldaux;
    float* C2; double* acc; float gscale; int g0;
    // EPI_DGRAD at the VQ boundary
    const float* z; const float* q; long long zq_gs; int ldzq; float cscale;
    // EPI_WGRAD
    float* db; long long db_gs; int ones_row; int zero_row_base;
    // weight indirection (reference core/dense.py:104-105, tf.gather(self.kernel, fts)): group g multiplies by the weights
    // (and adds the bias) of network gidx[g]; null = identity
    const int* gidx;
};

template <int ALAY, int BLAY, int EPI>
__global__ void __launch_bounds__(256) gemm_grouped_kernel(const GemmP p) {
    __shared__ __align__(16) float As[BK][AS_LD];
    __shared__ __align__(16) float Bs[BK][BS_LD];
    __shared__ double red[2][8];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int g = blockIdx.z / p.S, s = blockIdx.z - g * p.S;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = s * p.kchunk;
    const int kend = min(p.K, kbeg + p.kchunk);
    const float* __restrict__ A = p.A + (long long)g * p.a_gs;
    const int wg = p.gidx ? __ldg(p.gidx + g) : g;
    const float* __restrict__ Bm = p.B + (long long)wg * p.b_gs;
    const int Mext = p.M + ((EPI == EPI_WGRAD) ? p.ones_row : 0);

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb;

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * 256;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ALAY == 0) {
                const int row = f >> 2, kq = f & 3;
                const int gm = m0 + row, gk = k0 + kq * 4;
                if (gm < p.M && gk < kend) {
                    const float* ptr = A + (long long)gm * p.lda + gk;
                    if (p.vecA && gk + 3 < kend) {
                        v = __ldg(reinterpret_cast<const float4*>(ptr));
                    } else {
                        v.x = __ldg(ptr);
                        if (gk + 1 < kend) v.y = __ldg(ptr + 1);
                        if (gk + 2 < kend) v.z = __ldg(ptr + 2);
                        if (gk + 3 < kend) v.w = __ldg(ptr + 3);
                    }
                }
            } else {
                const int krow = f >> 5, mq = f & 31;
                const int gk = k0 + krow, gm = m0 + mq * 4;
                if (gk < kend && gm < Mext) {
                    const float* ptr = A + (long long)gk * p.lda + gm;
                    if (p.vecA && gm + 3 < p.M) {
                        v = __ldg(reinterpret_cast<const float4*>(ptr));
                    } else {
                        float t[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int mm = gm + j;
                            t[j] = mm < p.M ? __ldg(ptr + j) : ((mm == p.M && Mext > p.M) ? 1.0f : 0.f);
                        }
                        v = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
            }
            ra[i] = v;
        }
        {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (BLAY == 0) {
                const int krow = tid >> 4, nq = tid & 15;
                const int gk = k0 + krow, gn = n0 + nq * 4;
                if (gk < kend && gn < p.N) {
                    const float* ptr = Bm + (long long)gk * p.ldb + gn;
                    if (p.vecB && gn + 3 < p.N) {
                        v = __ldg(reinterpret_cast<const float4*>(ptr));
                    } else {
                        v.x = __ldg(ptr);
                        if (gn + 1 < p.N) v.y = __ldg(ptr + 1);
                        if (gn + 2 < p.N) v.z = __ldg(ptr + 2);
                        if (gn + 3 < p.N) v.w = __ldg(ptr + 3);
                    }
                }
            } else {
                const int nrow = tid >> 2, kq = tid & 3;
                const int gn = n0 + nrow, gk = k0 + kq * 4;
                if (gn < p.N && gk < kend) {
                    const float* ptr = Bm + (long long)gn * p.ldb + gk;
                    if (p.vecB && gk + 3 < kend) {
                        v = __ldg(reinterpret_cast<const float4*>(ptr));
                    } else {
                        v.x = __ldg(ptr);
                        if (gk + 1 < kend) v.y = __ldg(ptr + 1);
                        if (gk + 2 < kend) v.z = __ldg(ptr + 2);
                        if (gk + 3 < kend) v.w = __ldg(ptr + 3);
                    }
                }
            }
            rb = v;
        }
    };

    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * 256;
            if (ALAY == 0) {
                const int row = f >> 2, kq = f & 3;
                As[kq * 4 + 0][row] = ra[i].x;
                As[kq * 4 + 1][row] = ra[i].y;
                As[kq * 4 + 2][row] = ra[i].z;
                As[kq * 4 + 3][row] = ra[i].w;
            } else {
                const int krow = f >> 5, mq = f & 31;
                *reinterpret_cast<float4*>(&As[krow][mq * 4]) = ra[i];
            }
        }
        if (BLAY == 0) {
            const int krow = tid >> 4, nq = tid & 15;
            *reinterpret_cast<float4*>(&Bs[krow][nq * 4]) = rb;
        } else {
            const int nrow = tid >> 2, kq = tid & 3;
            Bs[kq * 4 + 0][nrow] = rb.x;
            Bs[kq * 4 + 1][nrow] = rb.y;
            Bs[kq * 4 + 2][nrow] = rb.z;
            Bs[kq * 4 + 3][nrow] = rb.w;
        }
    };

    if (kbeg < kend) {
        load_tiles(kbeg);
        for (int k0 = kbeg; k0 < kend; k0 += BK) {
            store_tiles();
            __syncthreads();
            if (k0 + BK < kend) load_tiles(k0 + BK);
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // ------------------------------------------------------------------ epilogues
    const int gn0 = n0 + tx * 4;
    if (EPI == EPI_BIAS_ACT) {
        float* __restrict__ C = p.C + (long long)g * p.c_gs;
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (gn0 + j < p.N) bv[j] = __ldg(p.bias + (long long)wg * p.bias_gs + gn0 + j);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gm = m0 + ty * 8 + i;
            if (gm >= p.M || gn0 >= p.N) continue;
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][j] + bv[j];
                o[j] = p.act == PGMVAE_ACT_SELU ? pg_selu(x) : (p.act == PGMVAE_ACT_SIGMOID ? pg_sigmoid(x) : x);
            }
            float* dst = C + (long long)gm * p.ldc + gn0;
            if (p.vecC && gn0 + 3 < p.N) {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gn0 + j < p.N) dst[j] = o[j];
            }
        }
    } else if (EPI == EPI_SIGMOID_MSE) {
        float* __restrict__ C = p.C + (long long)g * p.c_gs;
        float* __restrict__ C2 = p.C2 ? p.C2 + (long long)g * p.c_gs : nullptr;
        const int self = p.g0 + g;
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (gn0 + j < p.N) bv[j] = __ldg(p.bias + (long long)wg * p.bias_gs + gn0 + j);
        }
        float sq = 0.f, ab = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gm = m0 + ty * 8 + i;
            if (gm >= p.M || gn0 >= p.N) continue;
            float d4[4], o4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = gn0 + j;
                d4[j] = 0.f;
                o4[j] = 0.f;
                if (gn < p.N) {
                    const float o = pg_sigmoid(acc[i][j] + bv[j]);
                    o4[j] = o;
                    if (gn != self) {
                        const float t = __ldg(p.aux + (long long)gm * p.ldaux + gn);
                        const float d = o - t;
                        sq = fmaf(d, d, sq);
                        ab += fabsf(d);
                        d4[j] = p.gscale * d * o * (1.0f - o);
                    }
                }
            }
            float* dst = C + (long long)gm * p.ldc + gn0;
            if (p.vecC && gn0 + 3 < p.N) {
                *reinterpret_cast<float4*>(dst) = make_float4(d4[0], d4[1], d4[2], d4[3]);
                if (C2) *reinterpret_cast<float4*>(C2 + (long long)gm * p.ldc + gn0) = make_float4(o4[0], o4[1], o4[2], o4[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gn0 + j < p.N) {
                        dst[j] = d4[j];
                        if (C2) C2[(long long)gm * p.ldc + gn0 + j] = o4[j];
                    }
            }
        }
        double dsq = pg_warp_sum_d((double)sq), dab = pg_warp_sum_d((double)ab);
        const int w = tid >> 5;
        if ((tid & 31) == 0) { red[0][w] = dsq; red[1][w] = dab; }
        __syncthreads();
        if (tid == 0) {
            double a = 0, b = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
            atomicAdd(p.acc, a);
            atomicAdd(p.acc + 1, b);
        }
    } else if (EPI == EPI_DGRAD) {
        float* __restrict__ C = p.C + (long long)g * p.c_gs;
        const float* __restrict__ H = p.aux ? p.aux + (long long)g * p.aux_gs : nullptr;
        const float* __restrict__ Z = p.z ? p.z + (long long)g * p.zq_gs : nullptr;
        const float* __restrict__ Q = p.q ? p.q + (long long)g * p.zq_gs : nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gm = m0 + ty * 8 + i;
            if (gm >= p.M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = gn0 + j;
                if (gn >= p.N) continue;
                float v = acc[i][j];
                if (Z) v = fmaf(p.cscale, Z[(long long)gm * p.ldzq + gn] - Q[(long long)gm * p.ldzq + gn], v);
                if (H) {
                    const float h = H[(long long)gm * p.ldaux + gn];
                    if (p.act == PGMVAE_ACT_SELU) v *= pg_dselu_from_out(h);
                    else if (p.act == PGMVAE_ACT_SIGMOID) v *= h * (1.0f - h);
                }
                C[(long long)gm * p.ldc + gn] = v;
            }
        }
    } else {  // EPI_WGRAD
        float* __restrict__ C = p.C + (long long)g * p.c_gs;
        const int zrow = p.zero_row_base >= 0 ? p.zero_row_base + g : -1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gm = m0 + ty * 8 + i;
            if (gm >= Mext || gm == zrow) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = gn0 + j;
                if (gn >= p.N) continue;
                if (gm < p.M) atomicAdd(C + (long long)gm * p.ldc + gn, acc[i][j]);
                else atomicAdd(p.db + (long long)g * p.db_gs + gn, acc[i][j]);
            }
        }
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int ALAY, int BLAY, int EPI>
int launch(pgmvae_ctx* ctx, cudaStream_t st, GemmP& p, int G, int Mtiles_dim, const char* name, double bytes) {
    if (G <= 0 || p.M <= 0 || p.N <= 0) return PGMVAE_OK;
    dim3 grid((unsigned)pg_cdiv(p.N, BN), (unsigned)pg_cdiv(Mtiles_dim, BM), (unsigned)(G * p.S));
    if (grid.y > 65535u || grid.z > 65535u) {
        pgmvae_set_error("dense: grid too large (%u,%u,%u)", grid.x, grid.y, grid.z);
        return PGMVAE_EINVAL;
    }
    PG_KERNEL(ctx, st, name, bytes, 2.0 * G * (double)p.M * p.N * p.K);
    gemm_grouped_kernel<ALAY, BLAY, EPI><<<grid, 256, 0, st>>>(p);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

}  // namespace

int pg_dense_fwd_fp32(pgmvae_ctx* ctx, cudaStream_t stream, const float* x, int64_t x_gs, int ldx,
                          const float* w, int64_t w_gs, int ldw, const float* bias, int64_t bias_gs,
                          float* out, int64_t out_gs, int ldo, int G, int B, int in, int out_dim, int act, const int* gidx) {
    GemmP p{};
    p.gidx = gidx;
    p.A = x; p.a_gs = x_gs; p.lda = ldx;
    p.B = w; p.b_gs = w_gs; p.ldb = ldw;
    p.C = out; p.c_gs = out_gs; p.ldc = ldo;
    p.M = B; p.N = out_dim; p.K = in;
    p.vecA = aligned16(x) && ldx % 4 == 0 && x_gs % 4 == 0;
    p.vecB = aligned16(w) && ldw % 4 == 0 && w_gs % 4 == 0;
    p.vecC = aligned16(out) && ldo % 4 == 0 && out_gs % 4 == 0;
    p.S = 1; p.kchunk = pg_round_up(in > 0 ? in : 1, BK);
    p.bias = bias; p.bias_gs = bias_gs; p.act = act;
    const double xg = x_gs == 0 ? 1.0 : (double)G;
    return launch<0, 0, EPI_BIAS_ACT>(ctx, stream, p, G, B, "dense_fwd_fp32",
                                      4.0 * (xg * B * in + (double)G * in * out_dim + (double)G * out_dim +
                                             (double)G * B * out_dim));
}

int pg_dense_fwd_sigmoid_mse_fp32(pgmvae_ctx* ctx, cudaStream_t stream, const float* x, int64_t x_gs, int ldx,
                                      const float* w, int64_t w_gs, int ldw, const float* bias, int64_t bias_gs,
                                      const float* y, int ldy, float* dpre, int64_t dpre_gs, int ldd,
                                      float* out_opt, double* acc2, int G, int g0, int B, int in, int V,
                                      float grad_scale) {
    GemmP p{};
    p.A = x; p.a_gs = x_gs; p.lda = ldx;
    p.B = w; p.b_gs = w_gs; p.ldb = ldw;
    p.C = dpre; p.c_gs = dpre_gs; p.ldc = ldd;
    p.C2 = out_opt;
    p.M = B; p.N = V; p.K = in;
    p.vecA = aligned16(x) && ldx % 4 == 0 && x_gs % 4 == 0;
    p.vecB = aligned16(w) && ldw % 4 == 0 && w_gs % 4 == 0;
    p.vecC = aligned16(dpre) && ldd % 4 == 0 && dpre_gs % 4 == 0 && (!out_opt || aligned16(out_opt));
    p.S = 1; p.kchunk = pg_round_up(in > 0 ? in : 1, BK);
    p.bias = bias; p.bias_gs = bias_gs; p.act = PGMVAE_ACT_SIGMOID;
    p.aux = y; p.aux_gs = 0; p.ldaux = ldy;
    p.acc = acc2; p.gscale = grad_scale; p.g0 = g0;
    return launch<0, 0, EPI_SIGMOID_MSE>(ctx, stream, p, G, B, "dense_fwd_sigmoid_mse_fp32",
                                         4.0 * ((double)G * B * in + (double)G * in * V + (double)G * V + (double)B * V +
                                                (double)G * B * V * (out_opt ? 2 : 1)));
}

int pg_dense_dgrad_fp32(pgmvae_ctx* ctx, cudaStream_t stream, const float* dy, int64_t dy_gs, int lddy,
                            const float* w, int64_t w_gs, int ldw, const float* h_in, int64_t h_gs, int ldh,
                            const float* z, const float* q, int64_t zq_gs, int ldzq, float cscale,
                            float* dx, int64_t dx_gs, int lddx, int G, int B, int in, int out_dim,
                            int act_below) {
    GemmP p{};
    p.A = dy; p.a_gs = dy_gs; p.lda = lddy;
    p.B = w; p.b_gs = w_gs; p.ldb = ldw;          // read as [n = in][k = out], k contiguous
    p.C = dx; p.c_gs = dx_gs; p.ldc = lddx;
    p.M = B; p.N = in; p.K = out_dim;
    p.vecA = aligned16(dy) && lddy % 4 == 0 && dy_gs % 4 == 0;
    p.vecB = aligned16(w) && ldw % 4 == 0 && w_gs % 4 == 0;
    p.vecC = 0;
    p.S = 1; p.kchunk = pg_round_up(out_dim > 0 ? out_dim : 1, BK);
    p.aux = h_in; p.aux_gs = h_gs; p.ldaux = ldh; p.act = act_below;
    p.z = z; p.q = q; p.zq_gs = zq_gs; p.ldzq = ldzq; p.cscale = cscale;
    return launch<0, 1, EPI_DGRAD>(ctx, stream, p, G, B, "dense_dgrad_fp32",
                                   4.0 * ((double)G * B * out_dim + (double)G * in * out_dim +
                                          (double)G * B * in * (h_in ? 2 : 1) + (z ? 2.0 * G * B * in : 0.0)));
}

int pg_dense_wgrad_fp32(pgmvae_ctx* ctx, cudaStream_t stream, const float* x, int64_t x_gs, int ldx,
                            const float* dy, int64_t dy_gs, int lddy, float* dw, int64_t dw_gs, int lddw,
                            float* db, int64_t db_gs, int G, int B, int in, int out_dim, int zero_row_base) {
    GemmP p{};
    p.A = x; p.a_gs = x_gs; p.lda = ldx;          // read as [k = b][m = in], m contiguous
    p.B = dy; p.b_gs = dy_gs; p.ldb = lddy;
    p.C = dw; p.c_gs = dw_gs; p.ldc = lddw;
    p.M = in; p.N = out_dim; p.K = B;
    p.vecA = aligned16(x) && ldx % 4 == 0 && x_gs % 4 == 0;
    p.vecB = aligned16(dy) && lddy % 4 == 0 && dy_gs % 4 == 0;
    p.vecC = 0;
    p.db = db; p.db_gs = db_gs; p.ones_row = db ? 1 : 0; p.zero_row_base = zero_row_base;
    // split the batch so that the grid fills the machine (~4 CTAs per SM)
    const int64_t tiles = pg_cdiv(out_dim, BN) * pg_cdiv(in + p.ones_row, BM) * (int64_t)G;
    int S = (int)pg_cdiv((int64_t)ctx->sm_count * 4, tiles > 0 ? tiles : 1);
    const int maxS = (int)pg_cdiv(B, 256);
    if (S > maxS) S = maxS;
    if (S < 1) S = 1;
    while ((int64_t)G * S > 65535 && S > 1) --S;
    p.kchunk = pg_round_up((int)pg_cdiv(B, S), BK);
    p.S = (int)pg_cdiv(B, p.kchunk);
    const double xg = x_gs == 0 ? 1.0 : (double)G;
    return launch<1, 0, EPI_WGRAD>(ctx, stream, p, G, in + p.ones_row, "dense_wgrad_fp32",
                                   4.0 * (xg * B * in + (double)G * B * out_dim + (double)G * (in + 1) * out_dim));
}

