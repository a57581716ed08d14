// Shared definitions of libpgmvae.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/pgmvae.h"

// per-launch record of the optional kernel profiler (pgmvae_ctx_profile_begin/end)
struct pg_prof_rec {
    const char* name;
    cudaEvent_t e0, e1;
    cudaStream_t st;
    double bytes, flops;
};

struct pgmvae_ctx {
    bool profiling = false;
    bool prof_open = false;
    std::vector<pg_prof_rec> prof;
    int device = 0;
    int sm_count = 148;           // SMs the library's persistent kernels may fill (device SMs minus the reserved ones)
    int sm_total = 148;
    int precision = PGMVAE_PREC_FP32;
    bool coresident = false;      // the persistent GEMMs leave room for a co-resident exchange CTA (dense_bf16.cu: SLIM)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    size_t smem_optin = 0;
    void* scratch = nullptr;      // grown on demand by the tensor-core kernels
    size_t scratch_bytes = 0;
    size_t vq_cnt_off = 0;
    void* scratch_b = nullptr;    // bf16 operand copies of the operator-level bf16 entry points (dense_bf16_ops.cu)
    size_t scratch_b_bytes = 0;
};

void pgmvae_set_error(const char* fmt, ...);

#define PG_CHECK_ARG(cond)                                                          \
    do {                                                                            \
        if (!(cond)) {                                                              \
            pgmvae_set_error("%s: invalid argument: %s", __func__, #cond);          \
            return PGMVAE_EINVAL;                                                   \
        }                                                                           \
    } while (0)

#define PG_CUDA(call)                                                               \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            pgmvae_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call,             \
                             cudaGetErrorString(e__));                              \
            return PGMVAE_ECUDA;                                                    \
        }                                                                           \
    } while (0)

#define PG_TRY(call)                                                                \
    do {                                                                            \
        int r__ = (call);                                                           \
        if (r__ != PGMVAE_OK) return r__;                                           \
    } while (0)

// Before a kernel launch: name it and state its ALGORITHMIC bytes / flops (the figures the
// roofline is computed from).  With profiling on, the launch is bracketed by CUDA events.
void pg_prof_begin(pgmvae_ctx* ctx, cudaStream_t st, const char* name, double bytes, double flops);
void pg_prof_end(pgmvae_ctx* ctx);
#define PG_KERNEL(ctx, st, name, bytes, flops)                                      \
    do {                                                                            \
        if ((ctx)->profiling) pg_prof_begin((ctx), (st), (name), (double)(bytes), (double)(flops)); \
    } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define PG_LAUNCHED(ctx)                                                            \
    do {                                                                            \
        (ctx)->launches++;                                                          \
        if ((ctx)->profiling) pg_prof_end(ctx);                                     \
        PG_CUDA(cudaGetLastError());                                                \
    } while (0)

static inline cudaStream_t pg_stream(pgmvae_ctx* ctx, void* s) {
    return s ? (cudaStream_t)s : ctx->stream;
}

static inline int pg_round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int64_t pg_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// TensorFlow's fused selu constants (functor::Selu / SeluGrad)
#define PG_SELU_SCALE 1.0507009873554804934193349852946f
#define PG_SELU_SCALE_ALPHA 1.7580993408473768599402175208123f

__device__ __forceinline__ float pg_selu(float x) {
    return x < 0.f ? PG_SELU_SCALE_ALPHA * (expf(x) - 1.0f) : PG_SELU_SCALE * x;
}
// derivative expressed on the activation OUTPUT h (TF SeluGrad)
__device__ __forceinline__ float pg_dselu_from_out(float h) {
    return h < 0.f ? h + PG_SELU_SCALE_ALPHA : PG_SELU_SCALE;
}
__device__ __forceinline__ float pg_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float pg_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double pg_warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
