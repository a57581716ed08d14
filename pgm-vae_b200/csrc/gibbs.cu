// Sub-net ("fts") path and the block-Gibbs sampler on the device (reference core/model.py:98-148).
//
// The reference evaluates a SUBSET of the V networks through the `fts` branch of every layer
// (tf.gather(self.kernel, fts), core/dense.py:104-105; core/quantizer.py:134) and drives it from a Python loop of
// num_smp * p1 sampler steps, each of which rebuilds its inputs with tf.map_fn.  Here the weights are read IN PLACE
// from the model's parameter buffer through a per-group network index (no copy, no host round trip), the sampler
// state lives in HBM, and a step is 8 asynchronous launches:
//   network index of every block -> 5 packed dense layers (exact fp32 kernels with the weight indirection)
//   -> VQ assignment against the codebooks of the selected networks -> p(y = 1 | code) lookup + Bernoulli draw +
//   state / counter update.
// Leave-one-out inputs need no gather either: layers 0 / 9 are stored expanded over all V data columns with the
// weight row of a network's own variable fixed at zero, so block b simply feeds its whole state row.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ops.cuh"

// (model.cu) read-only view of what the sampler needs of a model
struct PgModelView {
    pgmvae_ctx* ctx;
    int V, Vp, D, Dp, K;
    const float* params;
    const float* E;
    struct { int in, out, pin, pout; size_t w_off, b_off; } L[5];
};
int pg_model_view(pgmvae_model* m, PgModelView* v);

namespace {

// y = marker + mod(i, vol)   (core/model.py:133)
__global__ void gibbs_fts_kernel(int* __restrict__ fts, int blocks, int p1, int dim, long long i) {
    const int b = threadIdx.x;
    if (b >= blocks) return;
    const int vol = b == blocks - 1 ? dim - p1 * (blocks - 1) : p1;
    fts[b] = b * p1 + (int)(i % vol);
}

__device__ __forceinline__ float u01(unsigned long long seed, unsigned long long ctr) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (ctr + 1);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    return (float)(x >> 40) * (1.0f / 16777216.0f);
}

// prb = dist[fts, code] (float32, core/model.py:107-108); gibbs = uniform < prb; state[b, :, y_b] = gibbs; counter (:138-141)
__global__ void gibbs_update_kernel(float* __restrict__ state, float* __restrict__ cnt, const int* __restrict__ fts,
                                    const int32_t* __restrict__ idx, const float* __restrict__ distf,
                                    const float* __restrict__ uni, unsigned long long seed, long long i, int blocks, int B,
                                    int Vp, int V, int K, int count) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (s >= B) return;
    const int v = fts[b];
    const float p = distf[(long long)v * K + idx[(long long)b * B + s]];
    const long long ctr = (i * blocks + b) * B + s;
    const float u = uni ? uni[ctr] : u01(seed, (unsigned long long)ctr);
    const float g = u < p ? 1.0f : 0.0f;
    state[((long long)b * B + s) * Vp + v] = g;
    if (count) cnt[(long long)s * V + v] += g;            // blocks own disjoint variables: no two threads share a cell
}

__global__ void gibbs_init_kernel(float* __restrict__ state, const uint8_t* __restrict__ x, int blocks, int B, int Vp, int V) {
    const long long n = (long long)blocks * B * Vp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Vp);
        const int s = (int)((i / Vp) % B);
        state[i] = c < V && x[(long long)s * V + c] ? 1.0f : 0.0f;
    }
}

// sum over (sample, variable) of x log(cmll + 1e-5) + (1 - x) log(1 - cmll + 1e-5), cmll = cnt / den  (core/model.py:146-149)
__global__ void __launch_bounds__(256) gibbs_reduce_kernel(const float* __restrict__ cnt, const uint8_t* __restrict__ x, int B, int V,
                                                           int last_from, float valid, float valid_end, double* out) {
    __shared__ double red[8];
    double acc = 0.0;
    const long long n = (long long)B * V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % V);
        const float c = cnt[i] / (v >= last_from ? valid_end : valid);
        acc += x[i] ? (double)logf(c + 1e-5f) : (double)logf(1.0f - c + 1e-5f);
    }
    acc = pg_warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int i = 0; i < 8; ++i) a += red[i];
        atomicAdd(out, a);
    }
}

struct Scratch {
    pgmvae_ctx* ctx;
    std::vector<void*> ptrs;
    int alloc(void** p, size_t bytes) {
        PG_TRY(pgmvae_malloc(ctx, bytes ? bytes : 16, p));
        ptrs.push_back(*p);
        return PGMVAE_OK;
    }
    ~Scratch() {
        for (void* p : ptrs) pgmvae_free(ctx, p);
    }
};

// encoder fd0..fd4 of the networks fts[f] on inputs x [F][B][Vp] (expanded over all data columns) + assignment
int fts_encode(const PgModelView& mv, cudaStream_t st, const float* x, const int* fts_dev, int F, int B, float* const* H,
               int32_t* idx) {
    for (int l = 0; l < 5; ++l) {
        const auto& L = mv.L[l];
        const float* in = l == 0 ? x : H[l - 1];
        const int ldx = l == 0 ? mv.Vp : mv.L[l - 1].pout;
        PG_TRY(pg_dense_fwd_fp32(mv.ctx, st, in, (int64_t)B * ldx, ldx, mv.params + L.w_off, (int64_t)L.pin * L.pout, L.pout,
                                 mv.params + L.b_off, L.pout, H[l], (int64_t)B * L.pout, L.pout, F, B, L.in, L.out,
                                 PGMVAE_ACT_SELU, fts_dev));
    }
    return pg_vq_assign_fp32(mv.ctx, st, H[4], (int64_t)B * mv.Dp, mv.Dp, mv.E, (int64_t)mv.K * mv.Dp, mv.Dp, idx, B, nullptr,
                             nullptr, F, B, mv.D, mv.K, fts_dev);
}

}  // namespace

extern "C" {

/* Codes of the selected networks (VqVAE.call(..., code_only=True, fts=...), core/model.py:41-48): x_exp [F][B][V] float32
 * on the HOST, expanded over all V data columns (column fts[f] of block f is ignored: its weight row is zero);
 * idx_host [F][B] int32.  Weights are read in place on the device. */
int pgmvae_model_fts_encode(pgmvae_model* m, const float* x_exp_host, const int32_t* fts_host, int F, int B, int32_t* idx_host) {
    PG_CHECK_ARG(m && x_exp_host && fts_host && idx_host && F >= 1 && B >= 1);
    PgModelView mv;
    PG_TRY(pg_model_view(m, &mv));
    for (int f = 0; f < F; ++f) PG_CHECK_ARG(fts_host[f] >= 0 && fts_host[f] < mv.V);
    cudaStream_t st = mv.ctx->stream;
    PG_CUDA(cudaSetDevice(mv.ctx->device));
    Scratch sc{mv.ctx};
    float *x = nullptr, *H[5];
    int* fts = nullptr;
    int32_t* idx = nullptr;
    PG_TRY(sc.alloc((void**)&x, (size_t)F * B * mv.Vp * 4));
    PG_TRY(sc.alloc((void**)&fts, (size_t)F * 4));
    PG_TRY(sc.alloc((void**)&idx, (size_t)F * B * 4));
    for (int l = 0; l < 5; ++l) PG_TRY(sc.alloc((void**)&H[l], (size_t)F * B * mv.L[l].pout * 4));
    PG_CUDA(cudaMemsetAsync(x, 0, (size_t)F * B * mv.Vp * 4, st));
    PG_CUDA(cudaMemcpy2DAsync(x, (size_t)mv.Vp * 4, x_exp_host, (size_t)mv.V * 4, (size_t)mv.V * 4, (size_t)F * B,
                              cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(fts, fts_host, (size_t)F * 4, cudaMemcpyHostToDevice, st));
    PG_TRY(fts_encode(mv, st, x, fts, F, B, H, idx));
    PG_CUDA(cudaMemcpyAsync(idx_host, idx, (size_t)F * B * 4, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    return PGMVAE_OK;
}

/* Conditional marginal log-likelihood by block Gibbs sampling, entirely on the device (core/model.py:110-148; intended call
 * run.py:74: p1 = n_var // 10, num_smp = 3000, burn_in = 150).  x [B][V] uint8 (host), dist [V][K] float64 (host, the CPT).
 * uniform_host: optional U[0,1) draws [num_smp * p1][blocks][B] float32 (parity runs inject them: the reference's tf.random
 * stream is not reproducible); NULL = counter-based generator seeded with `seed`. */
int pgmvae_model_gibbs_cmll(pgmvae_model* m, const uint8_t* x_host, int B, int p1, int num_smp, int burn_in,
                            const double* dist_host, uint64_t seed, const float* uniform_host, double* cmll_out) {
    PG_CHECK_ARG(m && x_host && dist_host && cmll_out && B >= 1 && p1 >= 1 && num_smp >= 1 && burn_in >= 0 && burn_in < num_smp);
    PgModelView mv;
    PG_TRY(pg_model_view(m, &mv));
    PG_CHECK_ARG(p1 <= mv.V);
    cudaStream_t st = mv.ctx->stream;
    PG_CUDA(cudaSetDevice(mv.ctx->device));
    const int V = mv.V, K = mv.K;
    const int blocks = (V + p1 - 1) / p1;                                       // :123
    const int vol_last = V - p1 * (blocks - 1);                                 // :124
    PG_CHECK_ARG(blocks <= 1024);
    const long long steps = (long long)num_smp * p1;
    Scratch sc{mv.ctx};
    float *state = nullptr, *cnt = nullptr, *distf = nullptr, *uni = nullptr, *H[5];
    uint8_t* x = nullptr;
    int* fts = nullptr;
    int32_t* idx = nullptr;
    double* out = nullptr;
    PG_TRY(sc.alloc((void**)&state, (size_t)blocks * B * mv.Vp * 4));
    PG_TRY(sc.alloc((void**)&cnt, (size_t)B * V * 4));
    PG_TRY(sc.alloc((void**)&distf, (size_t)V * K * 4));
    PG_TRY(sc.alloc((void**)&x, (size_t)B * V));
    PG_TRY(sc.alloc((void**)&fts, (size_t)blocks * 4));
    PG_TRY(sc.alloc((void**)&idx, (size_t)blocks * B * 4));
    PG_TRY(sc.alloc((void**)&out, 8));
    for (int l = 0; l < 5; ++l) PG_TRY(sc.alloc((void**)&H[l], (size_t)blocks * B * mv.L[l].pout * 4));
    if (uniform_host) {
        PG_TRY(sc.alloc((void**)&uni, (size_t)steps * blocks * B * 4));
        PG_CUDA(cudaMemcpyAsync(uni, uniform_host, (size_t)steps * blocks * B * 4, cudaMemcpyHostToDevice, st));
    }
    std::vector<float> df((size_t)V * K);
    for (size_t i = 0; i < df.size(); ++i) df[i] = (float)dist_host[i];          // tf.cast(dist, float32)  (:107)
    PG_CUDA(cudaMemcpyAsync(distf, df.data(), df.size() * 4, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(x, x_host, (size_t)B * V, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemsetAsync(cnt, 0, (size_t)B * V * 4, st));
    PG_CUDA(cudaMemsetAsync(out, 0, 8, st));
    gibbs_init_kernel<<<(unsigned)std::min<long long>(pg_cdiv((long long)blocks * B * mv.Vp, 256), 4096), 256, 0, st>>>(state, x, blocks, B,
                                                                                                                        mv.Vp, V);
    PG_LAUNCHED(mv.ctx);
    dim3 ugrid((unsigned)pg_cdiv(B, 128), (unsigned)blocks);
    for (long long i = 0; i < steps; ++i) {                                      // :132
        gibbs_fts_kernel<<<1, 1024, 0, st>>>(fts, blocks, p1, V, i);
        PG_LAUNCHED(mv.ctx);
        PG_TRY(fts_encode(mv, st, state, fts, blocks, B, H, idx));              // get_probability (:137 -> :99-108)
        gibbs_update_kernel<<<ugrid, 128, 0, st>>>(state, cnt, fts, idx, distf, uni, seed, i, blocks, B, mv.Vp, V, K,
                                                   i > (long long)burn_in * p1 ? 1 : 0);
        PG_LAUNCHED(mv.ctx);
        if ((i & 255) == 255) PG_CUDA(cudaStreamSynchronize(st));               // bound the launch queue
    }
    const float valid = (float)(num_smp - burn_in);                             // :146
    const float valid_end = floorf(valid * (float)p1 / (float)vol_last);        // :147  (float32 floor division)
    gibbs_reduce_kernel<<<(unsigned)std::min<long long>(pg_cdiv((long long)B * V, 256), 1024), 256, 0, st>>>(cnt, x, B, V, V - vol_last, valid,
                                                                                                            valid_end, out);
    PG_LAUNCHED(mv.ctx);
    double sum = 0.0;
    PG_CUDA(cudaMemcpyAsync(&sum, out, 8, cudaMemcpyDeviceToHost, st));
    PG_CUDA(cudaStreamSynchronize(st));
    *cmll_out = sum / B;                                                        // :149
    return PGMVAE_OK;
}

}  // extern "C"
