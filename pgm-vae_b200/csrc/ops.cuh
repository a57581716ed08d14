// Internal (non-exported) entry points shared between translation units.
#pragma once
#include "common.cuh"

// exact-fp32 CUDA-core kernels (dense_simt.cu, vq.cu)
int pg_dense_fwd_fp32(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w,
                      int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, float* out, int64_t out_gs, int ldo,
                      int G, int B, int in, int out_dim, int act);
int pg_dense_fwd_sigmoid_mse_fp32(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx,
                                  const float* w, int64_t w_gs, int ldw, const float* bias, int64_t bias_gs,
                                  const float* y, int ldy, float* dpre, int64_t dpre_gs, int ldd, float* out_opt,
                                  double* acc2, int G, int g0, int B, int in, int V, float grad_scale);
int pg_dense_dgrad_fp32(pgmvae_ctx* ctx, cudaStream_t st, const float* dy, int64_t dy_gs, int lddy, const float* w,
                        int64_t w_gs, int ldw, const float* h_in, int64_t h_gs, int ldh, const float* z,
                        const float* q, int64_t zq_gs, int ldzq, float cscale, float* dx, int64_t dx_gs, int lddx,
                        int G, int B, int in, int out_dim, int act_below);
int pg_dense_wgrad_fp32(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* dy,
                        int64_t dy_gs, int lddy, float* dw, int64_t dw_gs, int lddw, float* db, int64_t db_gs, int G,
                        int B, int in, int out_dim, int zero_row_base);
int pg_vq_assign_fp32(pgmvae_ctx* ctx, cudaStream_t st, const float* z, int64_t z_gs, int ldz, const float* e,
                      int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt, int G,
                      int B, int D, int K);

// tcgen05 tensor-core kernels (vq_tc.cu, dense_tc.cu)
bool pg_vq_assign_tc_supported(int prec, int D, int K, int ldz, int lde, const float* z, const float* e, int64_t z_gs,
                               int64_t e_gs);
int pg_vq_assign_tc(pgmvae_ctx* ctx, cudaStream_t st, int prec, const float* z, int64_t z_gs, int ldz, const float* e,
                    int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt, int G, int B,
                    int D, int K);
bool pg_vq_assign_f16_supported(int D, int K);
int pg_vq_assign_f16(pgmvae_ctx* ctx, cudaStream_t st, const float* z, int64_t z_gs, int ldz, const float* e,
                     int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt,
                     float* cnt_opt, int64_t c_gs, float* dw_opt, int64_t dw_gs, int lddw, int G, int B, int D, int K);
int pg_vq_assign_tc_last_flagged(pgmvae_ctx* ctx, int G, int K, int* out);
bool pg_dense_tc_supported(const float* a, int64_t a_gs, int lda, const float* b, int64_t b_gs, int ldb);
int pg_dense_fwd_tc(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w,
                    int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, float* out, int64_t out_gs, int ldo, int G,
                    int B, int in, int out_dim, int act);
int pg_dense_fwd_sigmoid_mse_tc(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w,
                                int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, const float* y, int ldy,
                                float* dpre, int64_t dpre_gs, int ldd, float* out_opt, double* acc2, int G, int g0, int B,
                                int in, int V, float grad_scale);
int pg_dense_dgrad_tc(pgmvae_ctx* ctx, cudaStream_t st, const float* dy, int64_t dy_gs, int lddy, const float* w,
                      int64_t w_gs, int ldw, const float* h_in, int64_t h_gs, int ldh, const float* z, const float* q,
                      int64_t zq_gs, int ldzq, float cscale, float* dx, int64_t dx_gs, int lddx, int G, int B, int in,
                      int out_dim, int act_below);
int pg_dense_wgrad_tc(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* dy,
                      int64_t dy_gs, int lddy, float* dw, int64_t dw_gs, int lddw, float* db, int64_t db_gs, int G, int B,
                      int in, int out_dim, int zero_row_base);
