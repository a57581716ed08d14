// Public operator entry points of kernels (a) and (b): argument checks and the choice of
// arithmetic (ctx precision): exact fp32 on CUDA cores, or tcgen05 tensor cores.
#include "common.cuh"
#include "ops.cuh"

extern "C" {

int pgmvae_dense_fwd(pgmvae_ctx* ctx, void* stream, const float* x, int64_t x_gs, int ldx, const float* w,
                     int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, float* out, int64_t out_gs, int ldo,
                     int G, int B, int in, int out_dim, int act) {
    PG_CHECK_ARG(ctx && x && w && out);
    PG_CHECK_ARG(G >= 0 && B >= 0 && in > 0 && out_dim > 0);
    PG_CHECK_ARG(ldx >= in && ldw >= out_dim && ldo >= out_dim);
    PG_CHECK_ARG(act >= PGMVAE_ACT_NONE && act <= PGMVAE_ACT_SIGMOID);
    if (ctx->precision == PGMVAE_PREC_BF16)
        return pg_dense_fwd_bf16(ctx, pg_stream(ctx, stream), x, x_gs, ldx, w, w_gs, ldw, bias, bias_gs, out, out_gs, ldo, G, B,
                                 in, out_dim, act);
    if (ctx->precision != PGMVAE_PREC_FP32 && pg_dense_tc_supported(x, x_gs, ldx, w, w_gs, ldw))
        return pg_dense_fwd_tc(ctx, pg_stream(ctx, stream), x, x_gs, ldx, w, w_gs, ldw, bias, bias_gs, out, out_gs, ldo,
                               G, B, in, out_dim, act);
    return pg_dense_fwd_fp32(ctx, pg_stream(ctx, stream), x, x_gs, ldx, w, w_gs, ldw, bias, bias_gs, out, out_gs, ldo,
                             G, B, in, out_dim, act);
}

int pgmvae_dense_fwd_sigmoid_mse(pgmvae_ctx* ctx, void* stream, const float* x, int64_t x_gs, int ldx, const float* w,
                                 int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, const float* y, int ldy,
                                 float* dpre, int64_t dpre_gs, int ldd, float* out_opt, double* acc2, int G, int g0,
                                 int B, int in, int V, float grad_scale) {
    PG_CHECK_ARG(ctx && x && w && y && dpre && acc2);
    PG_CHECK_ARG(G >= 0 && B >= 0 && in > 0 && V > 0 && g0 >= 0 && g0 + G <= V);
    PG_CHECK_ARG(ldx >= in && ldw >= V && ldd >= V && ldy >= V);
    if (ctx->precision == PGMVAE_PREC_BF16)
        return pg_dense_fwd_sigmoid_mse_bf16(ctx, pg_stream(ctx, stream), x, x_gs, ldx, w, w_gs, ldw, bias, bias_gs, y, ldy, dpre,
                                             dpre_gs, ldd, out_opt, acc2, G, g0, B, in, V, grad_scale);
    if (ctx->precision != PGMVAE_PREC_FP32 && pg_dense_tc_supported(x, x_gs, ldx, w, w_gs, ldw))
        return pg_dense_fwd_sigmoid_mse_tc(ctx, pg_stream(ctx, stream), x, x_gs, ldx, w, w_gs, ldw, bias, bias_gs, y, ldy,
                                           dpre, dpre_gs, ldd, out_opt, acc2, G, g0, B, in, V, grad_scale);
    return pg_dense_fwd_sigmoid_mse_fp32(ctx, pg_stream(ctx, stream), x, x_gs, ldx, w, w_gs, ldw, bias, bias_gs, y,
                                         ldy, dpre, dpre_gs, ldd, out_opt, acc2, G, g0, B, in, V, grad_scale);
}

int pgmvae_dense_dgrad(pgmvae_ctx* ctx, void* stream, const float* dy, int64_t dy_gs, int lddy, const float* w,
                       int64_t w_gs, int ldw, const float* h_in, int64_t h_gs, int ldh, const float* z,
                       const float* q, int64_t zq_gs, int ldzq, float cscale, float* dx, int64_t dx_gs, int lddx,
                       int G, int B, int in, int out_dim, int act_below) {
    PG_CHECK_ARG(ctx && dy && w && dx);
    PG_CHECK_ARG(G >= 0 && B >= 0 && in > 0 && out_dim > 0);
    PG_CHECK_ARG(lddy >= out_dim && ldw >= out_dim && lddx >= in);
    PG_CHECK_ARG((z == nullptr) == (q == nullptr));
    PG_CHECK_ARG(!h_in || ldh >= in);
    if (ctx->precision == PGMVAE_PREC_BF16)
        return pg_dense_dgrad_bf16(ctx, pg_stream(ctx, stream), dy, dy_gs, lddy, w, w_gs, ldw, h_in, h_gs, ldh, z, q, zq_gs, ldzq,
                                   cscale, dx, dx_gs, lddx, G, B, in, out_dim, act_below);
    if (ctx->precision != PGMVAE_PREC_FP32 && pg_dense_tc_supported(dy, dy_gs, lddy, w, w_gs, ldw))
        return pg_dense_dgrad_tc(ctx, pg_stream(ctx, stream), dy, dy_gs, lddy, w, w_gs, ldw, h_in, h_gs, ldh, z, q, zq_gs,
                                 ldzq, cscale, dx, dx_gs, lddx, G, B, in, out_dim, act_below);
    return pg_dense_dgrad_fp32(ctx, pg_stream(ctx, stream), dy, dy_gs, lddy, w, w_gs, ldw, h_in, h_gs, ldh, z, q,
                               zq_gs, ldzq, cscale, dx, dx_gs, lddx, G, B, in, out_dim, act_below);
}

int pgmvae_dense_wgrad(pgmvae_ctx* ctx, void* stream, const float* x, int64_t x_gs, int ldx, const float* dy,
                       int64_t dy_gs, int lddy, float* dw, int64_t dw_gs, int lddw, float* db, int64_t db_gs, int G,
                       int B, int in, int out_dim, int zero_row_base) {
    PG_CHECK_ARG(ctx && x && dy && dw);
    PG_CHECK_ARG(G >= 0 && B >= 0 && in > 0 && out_dim > 0);
    PG_CHECK_ARG(ldx >= in && lddy >= out_dim && lddw >= out_dim);
    if (ctx->precision == PGMVAE_PREC_BF16)
        return pg_dense_wgrad_bf16(ctx, pg_stream(ctx, stream), x, x_gs, ldx, dy, dy_gs, lddy, dw, dw_gs, lddw, db, db_gs, G, B,
                                   in, out_dim, zero_row_base);
    if (ctx->precision != PGMVAE_PREC_FP32 && pg_dense_tc_supported(x, x_gs, ldx, dy, dy_gs, lddy))
        return pg_dense_wgrad_tc(ctx, pg_stream(ctx, stream), x, x_gs, ldx, dy, dy_gs, lddy, dw, dw_gs, lddw, db, db_gs, G,
                                 B, in, out_dim, zero_row_base);
    return pg_dense_wgrad_fp32(ctx, pg_stream(ctx, stream), x, x_gs, ldx, dy, dy_gs, lddy, dw, dw_gs, lddw, db, db_gs,
                               G, B, in, out_dim, zero_row_base);
}

int pgmvae_vq_assign(pgmvae_ctx* ctx, void* stream, const float* z, int64_t z_gs, int ldz, const float* e,
                     int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt, int G,
                     int B, int D, int K) {
    PG_CHECK_ARG(ctx && z && e && idx);
    PG_CHECK_ARG(G >= 0 && B >= 0 && D > 0 && K > 0 && ldz >= D && lde >= D);
    if (ctx->precision == PGMVAE_PREC_BF16 && pg_vq_assign_f16_supported(D, K))
        return pg_vq_assign_f16(ctx, pg_stream(ctx, stream), z, z_gs, ldz, e, e_gs, lde, idx, idx_gs, best_opt, gap_opt,
                                nullptr, 0, nullptr, 0, 0, G, B, D, K);
    if (ctx->precision == PGMVAE_PREC_TF32 &&
        pg_vq_assign_tc_supported(ctx->precision, D, K, ldz, lde, z, e, z_gs, e_gs))
        return pg_vq_assign_tc(ctx, pg_stream(ctx, stream), ctx->precision, z, z_gs, ldz, e, e_gs, lde, idx, idx_gs,
                               best_opt, gap_opt, G, B, D, K);
    return pg_vq_assign_fp32(ctx, pg_stream(ctx, stream), z, z_gs, ldz, e, e_gs, lde, idx, idx_gs, best_opt, gap_opt,
                             G, B, D, K);
}

int pgmvae_vq_assign_ema(pgmvae_ctx* ctx, void* stream, const float* z, int64_t z_gs, int ldz, const float* e,
                         int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* counts, int64_t c_gs, float* dw,
                         int64_t dw_gs, int lddw, int G, int B, int D, int K) {
    PG_CHECK_ARG(ctx && z && e && idx && counts && dw);
    PG_CHECK_ARG(G >= 0 && B >= 0 && D > 0 && K > 0 && ldz >= D && lde >= D && lddw >= D);
    if (!pg_vq_assign_f16_supported(D, K)) {
        pgmvae_set_error("pgmvae_vq_assign_ema: D = %d exceeds the fused tensor-core kernel (D <= 126); "
                         "call pgmvae_vq_assign + pgmvae_ema_stats", D);
        return PGMVAE_EINVAL;
    }
    return pg_vq_assign_f16(ctx, pg_stream(ctx, stream), z, z_gs, ldz, e, e_gs, lde, idx, idx_gs, nullptr, nullptr,
                            counts, c_gs, dw, dw_gs, lddw, G, B, D, K);
}

/* rows of the last tensor-core vq_assign on this context that were re-scored in fp32 */
int pgmvae_vq_assign_rescored(pgmvae_ctx* ctx, int G, int K, int* out) {
    PG_CHECK_ARG(ctx && out);
    return pg_vq_assign_tc_last_flagged(ctx, G, K, out);
}

}  // extern "C"
