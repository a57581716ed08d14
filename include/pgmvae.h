/*
 * pgmvae.h  --  C-ABI of libpgmvae.so, the B200 (sm_100a) implementation of pgm-vae's
 * two-stage hot path (packed per-variable auto-encoders + VQ codebook training,
 * pseudo-log-likelihood evaluation).
 *
 * The reference (motionlife/pgm-vae) has no FFI: its boundary is the Python surface
 * core/dense.py, core/quantizer.py, core/model.py and run.py, all running on TensorFlow ops.
 * Each entry point below replaces the TensorFlow op group of one reference call site
 * (cited as file:line, relative to the reference root).  The Python host mirror in
 * pgm-vae_b200/core/ binds these symbols with ctypes (pgm-vae_b200/pgmvae/_ffi.py).
 *
 * Conventions
 *   - every function returns 0 on success, a PGMVAE_E* code otherwise;
 *     pgmvae_last_error() gives a thread-local message.  No exceptions cross the ABI.
 *   - the caller owns every buffer it passes; the library never frees caller memory and
 *     keeps pointers past a call only inside an explicit handle (ctx / model / comm).
 *   - operator entry points take DEVICE pointers and a `stream` (cudaStream_t as void*,
 *     NULL = the context's own stream) and are asynchronous.
 *   - tensors are fp32, row-major; "gs" is the element stride between two variables
 *     (groups), "ld" the element stride between two rows.  gs == 0 shares one matrix
 *     between all groups (the raw data matrix y of layer 0).
 *   - one handle is used by one host thread at a time; one process per GPU under DP.
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef PGMVAE_H
#define PGMVAE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGMVAE_VERSION 100 /* 0.1.0 */

enum {
    PGMVAE_OK = 0,
    PGMVAE_EINVAL = 1,   /* bad argument */
    PGMVAE_ECUDA = 2,    /* CUDA runtime / driver error */
    PGMVAE_ENOMEM = 3,
    PGMVAE_ENODEV = 4,   /* no CUDA device (there is no CPU fallback) */
    PGMVAE_ENCCL = 5,
    PGMVAE_ESTATE = 6
};

/* activation ids (core/model.py:19,36: 'selu' for fd0..fd8, 'sigmoid' for fd9) */
enum { PGMVAE_ACT_NONE = 0, PGMVAE_ACT_SELU = 1, PGMVAE_ACT_SIGMOID = 2 };

/* arithmetic of the GEMM-shaped kernels */
enum {
    PGMVAE_PREC_FP32 = 0, /* CUDA-core fp32 FMA: closest to the reference's fp32 maths */
    PGMVAE_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 accumulate in TMEM */
    PGMVAE_PREC_BF16 = 2  /* tcgen05 kind::f16 on 16-bit copies, fp32 accumulate in TMEM (the VQ
                             assignment uses fp16: same rate as bf16, 8x tighter error bound) */
};

typedef struct pgmvae_ctx pgmvae_ctx;
typedef struct pgmvae_model pgmvae_model;
typedef struct pgmvae_comm pgmvae_comm;

/* ------------------------------------------------------------------ context */
int pgmvae_version(void);
const char* pgmvae_last_error(void);
int pgmvae_device_count(int* n);
int pgmvae_ctx_create(int device, pgmvae_ctx** out);   /* run.py:27-31 device selection */
int pgmvae_ctx_destroy(pgmvae_ctx* ctx);
int pgmvae_ctx_sync(pgmvae_ctx* ctx);
void* pgmvae_ctx_stream(pgmvae_ctx* ctx);
int pgmvae_ctx_set_precision(pgmvae_ctx* ctx, int prec);
int pgmvae_ctx_get_precision(pgmvae_ctx* ctx);
/* Leave n SMs to other work (the NCCL kernels of the data-parallel exchange): the persistent kernels of the
 * library size their grids to the remaining SMs.  They assign tiles statically, one CTA per SM, so a CTA that
 * cannot be placed because a communication kernel holds its SM would start only when that kernel ends --
 * the overlap of exchange and compute needs the room to exist.  n = 0 restores the whole device. */
int pgmvae_ctx_reserve_sms(pgmvae_ctx* ctx, int n);
/* number of library kernels launched through this context since creation */
int64_t pgmvae_ctx_launch_count(pgmvae_ctx* ctx);

/* Optional per-kernel profiler: between begin and end every library launch on this context is
 * bracketed by CUDA events on its stream.  end() synchronises and writes a JSON array
 * [{"name","launches","ms","bytes","flops"}] (bytes/flops = ALGORITHMIC work of the launches). */
int pgmvae_ctx_profile_begin(pgmvae_ctx* ctx);
int pgmvae_ctx_profile_end(pgmvae_ctx* ctx, char* json_out, size_t cap);

/* device memory owned by the caller through the library allocator */
int pgmvae_malloc(pgmvae_ctx* ctx, size_t bytes, void** dptr);
int pgmvae_free(pgmvae_ctx* ctx, void* dptr);
int pgmvae_malloc_host(pgmvae_ctx* ctx, size_t bytes, void** hptr); /* pinned */
int pgmvae_free_host(pgmvae_ctx* ctx, void* hptr);
int pgmvae_memcpy_h2d(pgmvae_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int pgmvae_memcpy_d2h(pgmvae_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int pgmvae_memcpy_d2d(pgmvae_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int pgmvae_memset(pgmvae_ctx* ctx, void* dst, int byte, size_t bytes, void* stream);
/* device timing on the context stream (CUDA events) */
int pgmvae_timer_start(pgmvae_ctx* ctx);
int pgmvae_timer_stop_ms(pgmvae_ctx* ctx, float* ms);

/* ------------------------------------------------------ kernel (a): packed dense
 * Replaces tf.matmul(inputs, kernel) + bias -> activation of FatDense.call
 * (core/dense.py:99-111) for G independent networks:
 *     out[g] = act(x[g] [B,in] @ w[g] [in,out] + bias[g] [out])                        */
int pgmvae_dense_fwd(pgmvae_ctx* ctx, void* stream,
                     const float* x, int64_t x_gs, int ldx,
                     const float* w, int64_t w_gs, int ldw,
                     const float* bias, int64_t bias_gs,
                     float* out, int64_t out_gs, int ldo,
                     int G, int B, int in, int out_dim, int act);

/* Last layer fused with the Keras 'mse' loss / 'mae' metric and their gradient
 * (core/model.py:53 fd9 + run.py:61).  Variable g0+g reconstructs every column of y
 * except its own (leave-one-out, run.py:46-50), so column n == g0+g is masked:
 *     o      = sigmoid(x[g] @ w[g] + bias[g])                    [B, V]
 *     acc[0]+= sum_{n != g0+g} (o - y)^2 ;  acc[1] += sum |o - y|
 *     dpre   = grad_scale * (o - y) * o * (1 - o)   (0 in the masked column)
 * out_opt (may be NULL) receives o.                                                    */
int pgmvae_dense_fwd_sigmoid_mse(pgmvae_ctx* ctx, void* stream,
                                 const float* x, int64_t x_gs, int ldx,
                                 const float* w, int64_t w_gs, int ldw,
                                 const float* bias, int64_t bias_gs,
                                 const float* y, int ldy,
                                 float* dpre, int64_t dpre_gs, int ldd,
                                 float* out_opt,
                                 double* acc2,
                                 int G, int g0, int B, int in, int V, float grad_scale);

/* Input gradient of a dense layer, fused with the activation derivative of the layer
 * below (autodiff of core/dense.py:106-110, implicit in model.fit, run.py:62):
 *     dx[g] = (dy[g] [B,out] @ w[g]^T  +  cscale * (z - q)) * act'(h_in[g])
 * h_in is the OUTPUT of the activation below (TF SeluGrad uses outputs); NULL = identity.
 * z,q (may be NULL) add the commitment-loss gradient at the VQ boundary
 * (core/quantizer.py:50-53 / :142,153,156: straight-through + beta * d e_latent / dz).  */
int pgmvae_dense_dgrad(pgmvae_ctx* ctx, void* stream,
                       const float* dy, int64_t dy_gs, int lddy,
                       const float* w, int64_t w_gs, int ldw,
                       const float* h_in, int64_t h_gs, int ldh,
                       const float* z, const float* q, int64_t zq_gs, int ldzq, float cscale,
                       float* dx, int64_t dx_gs, int lddx,
                       int G, int B, int in, int out_dim, int act_below);

/* Weight / bias gradient:  dw[g] += x[g]^T [in,B] @ dy[g] [B,out] ;  db[g] += sum_b dy[g].
 * ACCUMULATES (zero dw/db first).  zero_row_base >= 0 keeps row (zero_row_base+g) of dw[g]
 * at zero: the leave-one-out mask of layer 0 when x is the shared raw matrix y.         */
int pgmvae_dense_wgrad(pgmvae_ctx* ctx, void* stream,
                       const float* x, int64_t x_gs, int ldx,
                       const float* dy, int64_t dy_gs, int lddy,
                       float* dw, int64_t dw_gs, int lddw,
                       float* db, int64_t db_gs,
                       int G, int B, int in, int out_dim, int zero_row_base);

/* ------------------------------------------------------ kernel (b): VQ assignment
 * Replaces distances + argmin (core/quantizer.py:44-47, :135-138):
 *     idx[g,b] = argmin_k ( (|z|^2 - 2 z.e_k) + |e_k|^2 ),  lowest index on ties.
 * The codebook is passed CODE-MAJOR: e[g] is [K, D] (the reference variable is [D,K]).
 * best_opt/gap_opt (may be NULL): smallest distance and distance gap to the runner-up.  */
int pgmvae_vq_assign(pgmvae_ctx* ctx, void* stream,
                     const float* z, int64_t z_gs, int ldz,
                     const float* e, int64_t e_gs, int lde,
                     int32_t* idx, int64_t idx_gs,
                     float* best_opt, float* gap_opt,
                     int G, int B, int D, int K);

/* Rows of the last tensor-core pgmvae_vq_assign on this context (precision TF32/BF16) whose
 * top-2 gap fell inside the low-precision error bound and were re-scored in exact fp32.
 * Synchronises the context stream.                                                       */
int pgmvae_vq_assign_rescored(pgmvae_ctx* ctx, int G, int K, int* out);

/* gather + losses + straight-through (core/quantizer.py:49-53, :141-142,156):
 *     q = e[idx];  st = z + (q - z);  loss_acc[0] += sum (q - z)^2                       */
int pgmvae_vq_quantize(pgmvae_ctx* ctx, void* stream,
                       const float* z, int64_t z_gs, int ldz,
                       const float* e, int64_t e_gs, int lde,
                       const int32_t* idx, int64_t idx_gs,
                       float* q, float* st, int64_t q_gs, int ldq,
                       double* loss_acc,
                       int G, int B, int D, int K);

/* codebook gradient of the non-EMA layer (core/quantizer.py:51: q_latent_loss):
 *     de[g,k,:] += scale * sum_{b: idx=k} (q - z)                                       */
int pgmvae_vq_codebook_grad(pgmvae_ctx* ctx, void* stream,
                            const float* z, const float* q, int64_t zq_gs, int ldzq,
                            const int32_t* idx, int64_t idx_gs,
                            float* de, int64_t de_gs, int ldde, float scale,
                            int G, int B, int D, int K);

/* ------------------------------------------------------ kernel (c): EMA statistics
 * Replaces reduce_sum(one_hot) and matmul(inputs^T, one_hot) (core/quantizer.py:144-146)
 * by a segmented scatter-add:  counts[g,k] += 1 ; dw[g,k,:] += z[g,b,:]  for k = idx[g,b].
 * ACCUMULATES (zero first); under data parallelism all-reduce counts/dw before ema_apply. */
int pgmvae_ema_stats(pgmvae_ctx* ctx, void* stream,
                     const float* z, int64_t z_gs, int ldz,
                     const int32_t* idx, int64_t idx_gs,
                     float* counts, int64_t c_gs,
                     float* dw, int64_t dw_gs, int lddw,
                     int G, int B, int D, int K);

/* Kernels (b)+(c) fused (BASELINE.json configs[3]: "fused distance-GEMM + argmin + EMA scatter"):
 * assignment as pgmvae_vq_assign, and in the same kernel counts[g,k] += 1 ; dw[g,k,:] += z[g,b,:]
 * for the chosen k, while the row is still hot (core/quantizer.py:135-138 + :144-146).
 * Tensor-core fp16 path only (needs D <= 126); ACCUMULATES into counts/dw.              */
int pgmvae_vq_assign_ema(pgmvae_ctx* ctx, void* stream,
                         const float* z, int64_t z_gs, int ldz,
                         const float* e, int64_t e_gs, int lde,
                         int32_t* idx, int64_t idx_gs,
                         float* counts, int64_t c_gs,
                         float* dw, int64_t dw_gs, int lddw,
                         int G, int B, int D, int K);

/* TF assign_moving_average(zero_debias=True) x2 + Laplace smoothing + normalise + write-back
 * (core/quantizer.py:144-152).  step = value of the hidden local_step AFTER this update.  */
int pgmvae_ema_apply(pgmvae_ctx* ctx, void* stream,
                     const float* counts, const float* dw,
                     float* biased_c, float* biased_w,
                     float* ema_c, float* ema_w,
                     float* e,
                     int G, int K, int D, int ld,
                     double decay, double epsilon, int step, int zero_debias);

/* ------------------------------------------------------ kernel (d): PLL
 * Replaces VqVAE.count's one-hot/boolean_mask/reduce_sum loops (core/model.py:58-82) by a
 * histogram over (g, idx[g,b], y[b,g0+g]):  n1 += [y != 0], n0 += [y == 0].  ACCUMULATES.  */
int pgmvae_pll_count(pgmvae_ctx* ctx, void* stream,
                     const int32_t* idx, int64_t idx_gs,
                     const uint8_t* y, int ldy, int g0,
                     unsigned long long* n1, unsigned long long* n0,
                     int G, int B, int K);
/* cpt (core/model.py:88): dist = (n1 + 0.8) / (n1 + n0 + 1.6), float64 */
int pgmvae_cpt(pgmvae_ctx* ctx, void* stream, const unsigned long long* n1,
               const unsigned long long* n0, double* dist, int64_t count);
/* core/model.py:93-96: *out_sum = sum n1*log(dist+1e-5) + n0*log(1-dist+1e-5)  (float64;
 * the caller divides by N).  out_sum is a device pointer.                               */
int pgmvae_pll_reduce(pgmvae_ctx* ctx, void* stream, const unsigned long long* n1,
                      const unsigned long long* n0, const double* dist, int64_t count,
                      double* out_sum);

/* ------------------------------------------------------ optimiser / data
 * Keras Adam, fused ResourceApplyAdam form (run.py:60):
 *   m += (g-m)(1-b1); v += (g*g-v)(1-b2); p -= alpha*m/(sqrt(v)+eps), alpha computed by caller */
int pgmvae_adam_step(pgmvae_ctx* ctx, void* stream, float* p, const float* g, float* m,
                     float* v, int64_t n, float alpha, double b1, double b2, double eps);
/* y [B,V] uint8 -> fp32 [B,ld] (pad columns zeroed): the only input the path reads;
 * the reference's materialised [N,V,V-1] tensor (run.py:48-50) is never built.          */
int pgmvae_y_to_f32(pgmvae_ctx* ctx, void* stream, const uint8_t* y, int ldy, float* out,
                    int ld, int B, int V);

/* ------------------------------------------------------ model handle
 * Device-resident state of one VqVAE (core/model.py:17-37) and the fused loops that
 * drive the kernels: one Keras fit step (run.py:62) and VqVAE.count (core/model.py:58-82). */
int pgmvae_model_create(pgmvae_ctx* ctx, const int* units4, int nvar, int dim, int k,
                        double cost, double decay, double epsilon, int ema, int max_batch,
                        pgmvae_model** out);
/* device-side initialisation with the reference's initialiser distributions (Keras fans
 * on the rank-3 shapes: he_uniform fd0..fd8, glorot_uniform fd9, VarianceScaling uniform
 * codebook, zero biases; core/model.py:19-36, core/quantizer.py:36,112-117) from a
 * counter-based generator keyed by seed.  TF's RNG stream itself is not reproducible.   */
int pgmvae_model_init(pgmvae_model* m, uint64_t seed);
int pgmvae_model_destroy(pgmvae_model* m);
/* tensors by reference name and in the REFERENCE layout (host fp32):
 *   "fd<i>.kernel" [V,in,out]  "fd<i>.bias" [V,1,out]  (core/dense.py:78-95)
 *   "vq.embeddings" [V,D,K]  "vq.ema_w" [V,D,K]  "vq.ema_cluster_size" [V,K] (core/quantizer.py:111-117)
 *   "vq.biased_w" [V,D,K]  "vq.biased_c" [V,K]  (TF's hidden zero-debias accumulators)
 *   "grad.fd<i>.kernel" / "grad.fd<i>.bias" / "grad.vq.embeddings" (read-only, last step)   */
int pgmvae_model_tensor_size(pgmvae_model* m, const char* name, int64_t* count);
int pgmvae_model_set_tensor(pgmvae_model* m, const char* name, const float* host, int64_t count);
int pgmvae_model_get_tensor(pgmvae_model* m, const char* name, float* host, int64_t count);
int pgmvae_model_set_ema_steps(pgmvae_model* m, int step_c, int step_w);
int pgmvae_model_set_adam_step(pgmvae_model* m, int64_t t);

/* Peer-to-peer gradient exchange (single node, data parallel; new work like pgmvae_comm_*; the reference is
 * single-device, run.py:27-31).  Every rank maps the gradient / parameter / bf16-mirror / Adam-moment buffers and the
 * flag block of every other rank (CUDA IPC).  Two kernels use the mapping instead of NCCL all-reduces + Adam:
 *   narrow models (one variable group, chain kernels): ONE kernel per rank reads the gradient buffers of all ranks
 *     over NVLink, sums them in rank order and applies the Adam update (two ranks by default; PGMVAE_P2P=1 forces it);
 *   wide models (per-group path): per variable group, reduce-scatter + Adam + all-gather as ONE kernel -- a rank sums
 *     ITS shard of the group's gradients from all peers, updates it and writes what the other ranks compute with
 *     (the bf16 mirror of a kernel; fp32 biases / fp32-mode parameters) into every rank's buffers; the CTAs are small enough to run next to the GEMMs of the following group
 *     (any rank count up to 8; PGMVAE_P2P_SHARD=0 keeps NCCL).  The replicas stay bit-identical either way.
 * export: writes 6 x 64 bytes (CUDA IPC handles: gradients, flag block, parameters, bf16 mirror, Adam m, Adam v);
 * the host exchanges them (any side channel) and passes all of them, in rank order, to import.  Optional: without
 * it pgmvae_model_train_step reduces the gradients through the communicator.  Every rank creates its model with the
 * same shapes, max_batch and group-size settings (the ownership of the shards follows the variable groups).    */
int pgmvae_model_p2p_export(pgmvae_model* m, void* handles_out384);
int pgmvae_model_p2p_import(pgmvae_model* m, int rank, int nranks, const void* all_handles);
/* After steps of the sharded exchange the Adam moments -- and, in bf16 mode, the fp32 master copy of the dense
 * kernels (every rank computes with their bf16 mirror, which IS exchanged) -- are complete only on the rank that owns
 * a shard (state_sharded() == 1); sync_state() -- COLLECTIVE over the ranks of the mapping -- completes them
 * everywhere (before get_tensor / a checkpoint / the fp32 readers such as the Gibbs sampler; fit() ends with it). */
int pgmvae_model_p2p_state_sharded(pgmvae_model* m);
/* which variables [*lo, *hi) of the variable group [g0, g0 + Gn) rank `rank` of `nranks` owns in the sharded exchange */
int pgmvae_p2p_shard_bounds(int g0, int Gn, int rank, int nranks, int* lo, int* hi);
int pgmvae_model_p2p_sync_state(pgmvae_model* m);
/* turn the peer-to-peer exchange off again (a rank failed to map its peers): NCCL is used instead */
int pgmvae_model_p2p_disable(pgmvae_model* m);
/* cudaDeviceCanAccessPeer(device, peer): whether the CUDA-IPC mapping behind pgmvae_model_p2p_import can work */
int pgmvae_device_can_access_peer(int device, int peer, int* out);

/* One training step on a batch y [B,V] uint8 (host or device pointer).
 * global_B is the batch size over all data-parallel ranks (== B without DP); comm may be NULL.
 * flags: bit0 = skip the optimiser/EMA update (gradients only).
 * metrics (host, may be NULL -> no sync): {loss, mse, mae, vq_loss} of this step.       */
int pgmvae_model_train_step(pgmvae_model* m, const uint8_t* y, int y_on_device, int B,
                            int global_B, float lr, pgmvae_comm* comm, int flags,
                            double* metrics4);
/* VqVAE.call (core/model.py:39-55) without the backward pass.  out_dev (device, may be NULL)
 * receives the reconstruction EXPANDED over all data columns: [V][max_batch][P(V)] with
 * P(V) = V rounded up to 8; column v of net v is the masked leave-one-out slot.
 * training != 0 also performs the EMA codebook update, as the reference call does.      */
int pgmvae_model_forward(pgmvae_model* m, const uint8_t* y, int y_on_device, int B, int training,
                         float* out_dev, double* metrics4);
/* codes of a batch: idx [V,B] int32 on the device (core/model.py:48 code_only path) */
int pgmvae_model_encode(pgmvae_model* m, const uint8_t* y, int y_on_device, int B,
                        int32_t* idx_dev);
/* VqVAE.count over N samples (core/model.py:58-82): n1,n0 [V,K] uint64 on the HOST. */
int pgmvae_model_count(pgmvae_model* m, const uint8_t* y, int y_on_device, int64_t N,
                       unsigned long long* n1_host, unsigned long long* n0_host);
/* The same over the variables [v0, v1) only (rows v0..v1-1 of n1 / n0 are written, the others left untouched):
 * stage 2 sharded over VARIABLE groups -- every rank evaluates its own nets on all samples and only the float64
 * PLL scalar crosses ranks (north_star; core/model.py:91-96 is a sum over variables). */
int pgmvae_model_count_vars(pgmvae_model* m, const uint8_t* y, int y_on_device, int64_t N, int v0, int v1,
                            unsigned long long* n1_host, unsigned long long* n0_host);
/* The same as a stream of chunks, for data sets larger than host or device memory (run.py:53: the author's TODO):
 * _begin zeroes the device counters, every _add enqueues the copy (host pointers: cudaMemcpyAsync, truly asynchronous
 * from pinned memory) and the kernels of one chunk and returns without waiting, _end downloads n1 / n0.  The caller
 * keeps a chunk's host buffer alive until pgmvae_ctx_sync or _end. */
int pgmvae_model_count_begin(pgmvae_model* m);
int pgmvae_model_count_add(pgmvae_model* m, const uint8_t* y, int y_on_device, int64_t N, int v0, int v1);
int pgmvae_model_count_end(pgmvae_model* m, int v0, int v1, unsigned long long* n1_host, unsigned long long* n0_host);
/* Sub-net path (reference `fts` branches: core/dense.py:104-105, core/quantizer.py:134, core/model.py:41-48): codes of the
 * networks fts[f] on inputs x_exp [F][B][V] float32 (HOST; expanded over all V data columns, column fts[f] of block f is
 * ignored), idx_host [F][B] int32.  The weights are read in place on the device through a per-group network index. */
int pgmvae_model_fts_encode(pgmvae_model* m, const float* x_exp_host, const int32_t* fts_host, int F, int B,
                            int32_t* idx_host);
/* VqVAE.conditional_marginal_log_likelihood (core/model.py:110-148; run.py:74) with the block-Gibbs sampler resident on
 * the device: x [B][V] uint8 and dist [V][K] float64 on the HOST; uniform_host (nullable) injects the U[0,1) draws
 * [num_smp * p1][blocks][B] float32, otherwise a counter-based generator seeded with `seed` draws them on the device. */
int pgmvae_model_gibbs_cmll(pgmvae_model* m, const uint8_t* x_host, int B, int p1, int num_smp, int burn_in,
                            const double* dist_host, uint64_t seed, const float* uniform_host, double* cmll_out);
/* arithmetic of the model's GEMM-shaped kernels, fixed at creation: 0 = fp32 CUDA cores, 1 = tf32 tcgen05
 * (per-layer or chain kernels), 2 = bf16 tcgen05 (wide networks under PGMVAE_PREC_BF16) */
int pgmvae_model_arithmetic(pgmvae_model* m);
/* bytes of device memory held by the model (weights, optimiser, workspace) */
int64_t pgmvae_model_device_bytes(pgmvae_model* m);
/* variables processed per group inside a step (workspace is sized for one group) */
int pgmvae_model_group_size(pgmvae_model* m);

/* ------------------------------------------------------ data-parallel communicator
 * New work (the reference is single-device): NCCL over NVLink, one rank per process.     */
int pgmvae_comm_unique_id(void* out128);                 /* ncclUniqueId, 128 bytes */
int pgmvae_comm_create(pgmvae_ctx* ctx, int rank, int nranks, const void* id128,
                       pgmvae_comm** out);
int pgmvae_comm_destroy(pgmvae_comm* c);
int pgmvae_comm_allreduce_f32(pgmvae_comm* c, float* buf, int64_t n, void* stream);
int pgmvae_comm_allreduce_f64(pgmvae_comm* c, double* buf, int64_t n, void* stream);
int pgmvae_comm_allreduce_u64(pgmvae_comm* c, unsigned long long* buf, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PGMVAE_H */
