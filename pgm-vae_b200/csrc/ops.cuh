// Internal (non-exported) entry points shared between translation units.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

// exact-fp32 CUDA-core kernels (dense_simt.cu, vq.cu)
int pg_dense_fwd_fp32(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w,
                      int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, float* out, int64_t out_gs, int ldo,
                      int G, int B, int in, int out_dim, int act, const int* gidx = nullptr);
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
                      int B, int D, int K, const int* gidx = nullptr);

// tcgen05 tensor-core kernels (vq_tc.cu, dense_tc.cu)
bool pg_vq_assign_tc_supported(int prec, int D, int K, int ldz, int lde, const float* z, const float* e, int64_t z_gs,
                               int64_t e_gs);
int pg_vq_assign_tc(pgmvae_ctx* ctx, cudaStream_t st, int prec, const float* z, int64_t z_gs, int ldz, const float* e,
                    int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt, int G, int B,
                    int D, int K);
bool pg_vq_assign_f16_supported(int D, int K);
int pg_vq_assign_f16(pgmvae_ctx* ctx, cudaStream_t st, const float* z, int64_t z_gs, int ldz, const float* e,
                     int64_t e_gs, int lde, int32_t* idx, int64_t idx_gs, float* best_opt, float* gap_opt,
                     float* cnt_opt, int64_t c_gs, float* dw_opt, int64_t dw_gs, int lddw, int G, int B, int D, int K,
                     float* q_opt = nullptr, __nv_bfloat16* stb_opt = nullptr, int64_t q_gs = 0, int ldq = 0,
                     double* loss_opt = nullptr);
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

// every weight-gradient GEMM of a step in one launch (dense_tc.cu)
struct PgWgradProblem {
    const float* x; int64_t x_gs; int ldx;          // layer input  [G][B][in]  (x_gs = 0: shared by all variables)
    const float* dy; int64_t dy_gs; int lddy;       // d(loss)/d(pre-activation) [G][B][out]
    float* dw; int64_t dw_gs; int lddw;             // += x^T dy   [G][in][out]
    float* db; int64_t db_gs;                       // += column sums of dy [G][out]
    int G, B, in, out, zero_row_base;
};
bool pg_dense_wgrad_multi_supported(const PgWgradProblem* pr, int n);
int pg_dense_wgrad_multi_tc(pgmvae_ctx* ctx, cudaStream_t st, const PgWgradProblem* pr, int n);

// chain kernels (chain_tc.cu): a stack of dense layers per launch, activations resident in TMEM
#define PG_CHAIN_MAX_STAGES 20
enum { PG_CHAIN_FWD = 0, PG_CHAIN_ENCODE = 1, PG_CHAIN_BWD = 2, PG_CHAIN_TRAIN = 3 };   // TRAIN = FWD stages followed by BWD stages
enum { PG_CHAIN_EPI_SELU = 0, PG_CHAIN_EPI_SIGMOID_MSE = 1, PG_CHAIN_EPI_DGRAD = 2 };
struct PgChainStage {
    // filled by the caller
    int K;                 // padded contraction width = columns of the A operand in TMEM (multiple of 8)
    int pout;              // padded output width (multiple of 8)
    int k_valid, n_valid;  // logical extents of the weight matrix along K and N (TMA zero-fills the rest)
    int b_mn;              // 1: W is [K][N] with N contiguous (forward); 0: W is [N][K] with K contiguous (dgrad)
    int kind;              // PG_CHAIN_EPI_*
    int add_commit;        // dgrad at the VQ boundary: add cscale * (z - q) before act'
    const float* w; long long w_gs; int ldw;
    const float* bias; long long bias_gs;
    float* outp; long long out_gs; int ldo;              // row written to HBM (activation / gradient), may be null
    const float* aux; long long aux_gs; int ldaux;       // dgrad: activation of the layer below (for act')
    // filled by pg_chain_launch
    int N, ksteps, kblocks, a_col, d_col, bias_off;
    unsigned kb_bytes;
};
struct PgChainArgs {
    int mode, nst;
    PgChainStage st[PG_CHAIN_MAX_STAGES];
    int G, g0, B, V, Vp, D, Dp, K, vq_stage;            // vq_stage < 0: no VQ in this chain
    const float* a0; long long a0_gs; int lda0, a0_cols; // first operand: rows of width a0_cols (multiple of 8)
    const float* yf; int ldyf;                           // targets of the MSE stage
    const uint8_t* y8; int ldy8;                         // encode: raw data for the PLL histogram
    const float* E; long long e_gs;                      // codebook [G][K][Dp]
    float* q; float* stq; long long zq_gs; int ldzq;
    int32_t* idx; long long idx_gs;
    float* stat_c; float* stat_w;                        // fused EMA statistics [G][K], [G][K][Dp] (nullable)
    double* acc;                                         // [0] sum sq err, [1] sum abs err, [2] sum (q - z)^2
    float gscale, cscale;
    unsigned long long* n1; unsigned long long* n0;      // encode: [G][K] histograms (nullable)
    const float* z; const float* qv;                     // backward: latent and quantised latent
};
int pg_chain_launch(pgmvae_ctx* ctx, cudaStream_t st, const PgChainArgs& a);
bool pg_chain_supported(const int* pin, const int* pout, int nlayers, int Vp, int Dp, int K, size_t smem_optin);

// bf16 tensor-core grouped GEMMs (dense_bf16.cu): persistent tcgen05 kernels on bf16 copies of the operands.
//   x / dy / h: row-major bf16 activations [G][B][ld];  wt: transposed weight shadow [G][out][ld(in)], or with
//   w_mn = 1 the shadow AS STORED [G][in][ld(out)] read as an MN-major operand (no transposed copy needed);
//   w: weight shadow as stored [G][in][ld(out)];  outputs in bf16 and / or fp32 (nullable)
int pg_bf16_fwd(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* x, int64_t x_gs, int ldx, const __nv_bfloat16* wt,
                int64_t wt_gs, int ldwt, const float* bias, int64_t bias_gs, __nv_bfloat16* outb, int64_t outb_gs, int ldob,
                float* outf, int64_t outf_gs, int ldof, int G, int B, int in, int out_dim, int act, int w_mn = 0);
int pg_bf16_fwd_sigmoid_mse(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* x, int64_t x_gs, int ldx,
                            const __nv_bfloat16* wt, int64_t wt_gs, int ldwt, const float* bias, int64_t bias_gs,
                            const uint32_t* ybits, int ldbits, __nv_bfloat16* dpre, int64_t dpre_gs, int ldd, float* out_opt,
                            int64_t out_gs, int ldo, double* acc2, int G, int g0, int B, int in, int V, float grad_scale, int w_mn = 0);
int pg_bf16_dgrad(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* dy, int64_t dy_gs, int lddy, const __nv_bfloat16* w,
                  int64_t w_gs, int ldw, const __nv_bfloat16* hb, int64_t hb_gs, int ldhb, const float* hf, int64_t hf_gs,
                  int ldhf, const float* z, const float* q, int64_t zq_gs, int ldzq, float cscale, __nv_bfloat16* dxb,
                  int64_t dxb_gs, int lddxb, float* dxf, int64_t dxf_gs, int lddxf, int G, int B, int in, int out_dim,
                  int act_below);
int pg_bf16_wgrad(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* x, int64_t x_gs, int ldx, const __nv_bfloat16* dy,
                  int64_t dy_gs, int lddy, float* dw, int64_t dw_gs, int lddw, float* db, int64_t db_gs, int G, int B, int in,
                  int out_dim, int zero_row_base, int accumulate);
int pg_bf16_colsum(pgmvae_ctx* ctx, cudaStream_t st, const __nv_bfloat16* dy, int64_t dy_gs, int lddy, float* db, int64_t db_gs,
                   int G, int B, int N, int accumulate);
// operator-level wrappers on fp32 tensors (dense_bf16_ops.cu): same signatures as the tf32 / fp32 flavours
int pg_dense_fwd_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w, int64_t w_gs,
                      int ldw, const float* bias, int64_t bias_gs, float* out, int64_t out_gs, int ldo, int G, int B, int in,
                      int out_dim, int act);
int pg_dense_fwd_sigmoid_mse_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* w,
                                  int64_t w_gs, int ldw, const float* bias, int64_t bias_gs, const float* y, int ldy,
                                  float* dpre, int64_t dpre_gs, int ldd, float* out_opt, double* acc2, int G, int g0, int B,
                                  int in, int V, float grad_scale);
int pg_dense_dgrad_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* dy, int64_t dy_gs, int lddy, const float* w, int64_t w_gs,
                        int ldw, const float* h_in, int64_t h_gs, int ldh, const float* z, const float* q, int64_t zq_gs,
                        int ldzq, float cscale, float* dx, int64_t dx_gs, int lddx, int G, int B, int in, int out_dim,
                        int act_below);
int pg_dense_wgrad_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* x, int64_t x_gs, int ldx, const float* dy, int64_t dy_gs,
                        int lddy, float* dw, int64_t dw_gs, int lddw, float* db, int64_t db_gs, int G, int B, int in,
                        int out_dim, int zero_row_base);
int pg_f32_to_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* src, int64_t s_gs, int lds, __nv_bfloat16* dst, int64_t d_gs,
                   int ldd, int G, int rows, int cols);
int pg_bf16_shadow(pgmvae_ctx* ctx, cudaStream_t st, const float* w, int64_t w_gs, int lds, int rows, int cols,
                   __nv_bfloat16* wt, int64_t t_gs, int ldt, __nv_bfloat16* wc, int64_t c_gs, int ldc, int G);
int pg_flat_to_bf16(pgmvae_ctx* ctx, cudaStream_t st, const float* src, __nv_bfloat16* dst, int64_t n);
int pg_adam_step_shadow(pgmvae_ctx* ctx, cudaStream_t stream, float* p, const float* g, float* m, float* v, int64_t n,
                        float alpha, double b1, double b2, double eps, __nv_bfloat16* wb, int64_t nwb);
int pg_y_to_bits(pgmvae_ctx* ctx, cudaStream_t st, const uint8_t* y, int ldy, uint32_t* bits, int ldbits, int B, int V);
int pg_f32_to_bits(pgmvae_ctx* ctx, cudaStream_t st, const float* y, int ldy, uint32_t* bits, int ldbits, int B, int V);
int pg_y_to_bf16(pgmvae_ctx* ctx, cudaStream_t st, const uint8_t* y, int ldy, __nv_bfloat16* out, int ld, int B, int V);
