// Operator-level bf16 entry points: the fp32 tensors of the C-ABI dense operators (pgmvae_dense_*,
// include/pgmvae.h) are copied to bf16 scratch operands and run through the persistent tcgen05
// kernels of dense_bf16.cu.  The model (model.cu) keeps its operands in bf16 and calls those
// kernels directly; these wrappers exist so that every bf16 GEMM flavour can be checked against the
// oracle through the same operator ABI as the fp32 / tf32 flavours (reference core/dense.py:99-111).
#include <algorithm>

#include "common.cuh"
#include "ops.cuh"

namespace {

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, long long s_gs, int lds, float* __restrict__ dst,
                                   long long d_gs, int ldd, int rows, int cols) {
    const int g = blockIdx.y;
    const long long n = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
        dst[(long long)g * d_gs + (long long)r * ldd + c] = __bfloat162float(src[(long long)g * s_gs + (long long)r * lds + c]);
    }
}

struct Carver {
    uint8_t* base; size_t off = 0;
    template <class T> T* take(size_t n) {
        T* p = reinterpret_cast<T*>(base + off);
        off += (n * sizeof(T) + 255) & ~(size_t)255;
        return p;
    }
};

int ensure(pgmvae_ctx* ctx, size_t bytes) {
    if (ctx->scratch_b_bytes >= bytes) return PGMVAE_OK;
    if (ctx->scratch_b) {
        PG_CUDA(cudaDeviceSynchronize());
        PG_CUDA(cudaFree(ctx->scratch_b));
        ctx->scratch_b = nullptr;
        ctx->scratch_b_bytes = 0;
    }
    PG_CUDA(cudaMalloc(&ctx->scratch_b, bytes));
    ctx->scratch_b_bytes = bytes;
    return PGMVAE_OK;
}

inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }
inline int P8(int n) { return pg_round_up(n, 8); }

}  // namespace

int pg_dense_fwd_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w, int64_t w_gs,
                      int ldw, const float* bias, int64_t bias_gs, float* out, int64_t out_gs, int ldo, int G, int B, int in,
                      int out_dim, int act) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    // the weights as the model keeps them: a bf16 copy in the stored [in][out] orientation, read MN-major
    const int pin = P8(in), po = P8(out_dim), xg = x_gs == 0 ? 1 : G;
    const size_t nx = (size_t)xg * B * pin, nw = (size_t)G * in * po;
    PG_TRY(ensure(ctx, pad256(nx * 2) + pad256(nw * 2)));
    Carver c{(uint8_t*)ctx->scratch_b};
    __nv_bfloat16* xb = c.take<__nv_bfloat16>(nx);
    __nv_bfloat16* wc = c.take<__nv_bfloat16>(nw);
    PG_TRY(pg_f32_to_bf16(ctx, st, x, x_gs, ldx, xb, (int64_t)B * pin, pin, xg, B, in));
    PG_TRY(pg_bf16_shadow(ctx, st, w, w_gs, ldw, in, out_dim, nullptr, 0, 0, wc, (int64_t)in * po, po, G));
    return pg_bf16_fwd(ctx, st, xb, x_gs == 0 ? 0 : (int64_t)B * pin, pin, wc, (int64_t)in * po, po, bias, bias_gs, nullptr,
                       0, 0, out, out_gs, ldo, G, B, in, out_dim, act, 1);
}

int pg_dense_fwd_sigmoid_mse_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w,
                                  int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, const float* y, int ldy,
                                  float* dpre, int64_t dpre_gs, int ldd, float* out_opt, double* acc2, int G, int g0, int B,
                                  int in, int V, float grad_scale) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    const int pin = P8(in), pv = P8(V), xg = x_gs == 0 ? 1 : G;
    const int ldbits = pv / 32 + 2;
    const size_t nx = (size_t)xg * B * pin, nw = (size_t)G * in * pv, ny = (size_t)B * ldbits * 2, nd = (size_t)G * B * pv;
    PG_TRY(ensure(ctx, pad256(nx * 2) + pad256(nw * 2) + pad256(ny * 2) + pad256(nd * 2)));
    Carver c{(uint8_t*)ctx->scratch_b};
    __nv_bfloat16* xb = c.take<__nv_bfloat16>(nx);
    __nv_bfloat16* wt = c.take<__nv_bfloat16>(nw);
    uint32_t* yb = reinterpret_cast<uint32_t*>(c.take<__nv_bfloat16>(ny));
    __nv_bfloat16* db = c.take<__nv_bfloat16>(nd);
    PG_TRY(pg_f32_to_bf16(ctx, st, x, x_gs, ldx, xb, (int64_t)B * pin, pin, xg, B, in));
    PG_TRY(pg_bf16_shadow(ctx, st, w, w_gs, ldw, in, V, nullptr, 0, 0, wt, (int64_t)in * pv, pv, G));
    PG_TRY(pg_f32_to_bits(ctx, st, y, ldy, yb, ldbits, B, V));
    PG_TRY(pg_bf16_fwd_sigmoid_mse(ctx, st, xb, x_gs == 0 ? 0 : (int64_t)B * pin, pin, wt, (int64_t)in * pv, pv, bias, bias_gs, yb,
                                   ldbits, db, (int64_t)B * pv, pv, out_opt, dpre_gs, ldd, acc2, G, g0, B, in, V, grad_scale, 1));
    dim3 grid((unsigned)std::min<int64_t>(pg_cdiv((int64_t)B * V, 1024), 2048), (unsigned)G);
    bf16_to_f32_kernel<<<grid, 256, 0, st>>>(db, (long long)B * pv, pv, dpre, dpre_gs, ldd, B, V);
    PG_LAUNCHED(ctx);
    return PGMVAE_OK;
}

int pg_dense_dgrad_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* dy, int64_t dy_gs, int lddy, const float* w, int64_t w_gs,
                        int ldw, const float* h_in, int64_t h_gs, int ldh, const float* z, const float* q, int64_t zq_gs,
                        int ldzq, float cscale, float* dx, int64_t dx_gs, int lddx, int G, int B, int in, int out_dim,
                        int act_below) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    const int po = P8(out_dim), dg = dy_gs == 0 ? 1 : G;
    const size_t nd = (size_t)dg * B * po, nw = (size_t)G * in * po;
    PG_TRY(ensure(ctx, pad256(nd * 2) + pad256(nw * 2)));
    Carver c{(uint8_t*)ctx->scratch_b};
    __nv_bfloat16* dyb = c.take<__nv_bfloat16>(nd);
    __nv_bfloat16* wc = c.take<__nv_bfloat16>(nw);
    PG_TRY(pg_f32_to_bf16(ctx, st, dy, dy_gs, lddy, dyb, (int64_t)B * po, po, dg, B, out_dim));
    PG_TRY(pg_bf16_shadow(ctx, st, w, w_gs, ldw, in, out_dim, nullptr, 0, 0, wc, (int64_t)in * po, po, G));
    return pg_bf16_dgrad(ctx, st, dyb, dy_gs == 0 ? 0 : (int64_t)B * po, po, wc, (int64_t)in * po, po, nullptr, 0, 0, h_in, h_gs, ldh,
                         z, q, zq_gs, ldzq, cscale, nullptr, 0, 0, dx, dx_gs, lddx, G, B, in, out_dim, act_below);
}

int pg_dense_wgrad_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* dy, int64_t dy_gs,
                        int lddy, float* dw, int64_t dw_gs, int lddw, float* db, int64_t db_gs, int G, int B, int in,
                        int out_dim, int zero_row_base) {
    if (G <= 0 || B <= 0) return PGMVAE_OK;
    const int pin = P8(in), po = P8(out_dim), xg = x_gs == 0 ? 1 : G;
    const size_t nx = (size_t)xg * B * pin, nd = (size_t)G * B * po;
    PG_TRY(ensure(ctx, pad256(nx * 2) + pad256(nd * 2)));
    Carver c{(uint8_t*)ctx->scratch_b};
    __nv_bfloat16* xb = c.take<__nv_bfloat16>(nx);
    __nv_bfloat16* dyb = c.take<__nv_bfloat16>(nd);
    PG_TRY(pg_f32_to_bf16(ctx, st, x, x_gs, ldx, xb, (int64_t)B * pin, pin, xg, B, in));
    PG_TRY(pg_f32_to_bf16(ctx, st, dy, dy_gs, lddy, dyb, (int64_t)B * po, po, G, B, out_dim));
    // the operator ABI accumulates into dw / db (+=)
    return pg_bf16_wgrad(ctx, st, xb, x_gs == 0 ? 0 : (int64_t)B * pin, pin, dyb, (int64_t)B * po, po, dw, dw_gs, lddw, db, db_gs, G,
                         B, in, out_dim, zero_row_base, 1);
}
