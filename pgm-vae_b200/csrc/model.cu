// Model handle: device-resident state of one VqVAE (reference core/model.py:17-37) and the
// fused loops that drive the kernels:
//   train_step  = one Keras fit step (run.py:62): forward (core/model.py:39-55), MSE + VQ loss,
//                 backward, Adam, EMA codebook update (core/quantizer.py:144-152)
//   count       = VqVAE.count (core/model.py:58-82): encoder + assignment + histogram
//
// HBM layout (P(n) = n rounded up to 8 floats):
//   layer l weights   W_l [V][P(in_l)][P(out_l)], bias b_l [V][P(out_l)]; zero padding.
//     Layer 0 and 9 are stored EXPANDED over all V data columns: net v reads the raw data
//     matrix y[B,V] with weight row v fixed at zero (fd0) / reconstructs all V columns with
//     column v masked out of the loss (fd9).  This is the reference's leave-one-out input
//     (run.py:46-50) without ever materialising [N,V,V-1].
//   codebook          E [V][K][P(D)] code-major (reference variable is [V,D,K]).
//   activations       H_l [Vg][B][P(out_l)] for one group of Vg variables at a time; networks
//     of different variables are independent, so a step walks the variables group by group
//     and the workspace is sized for one group.
//
// Data parallelism (new work; the reference is single-device, run.py:27-31): NCCL all-reduces
// of gradients / EMA statistics / loss sums (comm.cu), or -- ranks of one node that can map
// each other's buffers -- the peer-to-peer kernels below: a whole-buffer sum + Adam for narrow
// models, and for wide models a per-variable-group reduce-scatter + Adam + all-gather in one
// kernel that runs next to the GEMMs of the following group.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "ops.cuh"

struct pgmvae_comm;
int pg_comm_allreduce(pgmvae_comm* c, void* buf, int64_t n, int dtype, cudaStream_t st);  // comm.cu
int pg_comm_group_begin(pgmvae_comm* c);
int pg_comm_group_end(pgmvae_comm* c);

namespace {

struct Layer {
    int in, out;        // logical (expanded for layer 0 / 9)
    int pin, pout;      // padded
    int ref_in, ref_out;  // reference shapes (V-1 for the leave-one-out dims)
    size_t w_off, b_off;  // float offsets into the parameter buffer
    int act;
};

__global__ void init_uniform_kernel(float* __restrict__ dst, long long n, int rows, int cols, int prow, int pcol,
                                    int zero_row_is_v, int zero_col_is_v, float limit, unsigned long long seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % pcol);
    const long long t = i / pcol;
    const int r = (int)(t % prow);
    const int v = (int)(t / prow);
    float val = 0.f;
    if (r < rows && c < cols && !(zero_row_is_v && r == v) && !(zero_col_is_v && c == v)) {
        // splitmix64 on (seed, element index) -> uniform in [-limit, limit)
        unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        x = x ^ (x >> 31);
        const float u = (float)(x >> 40) * (1.0f / 16777216.0f);   // [0,1)
        val = (2.0f * u - 1.0f) * limit;
    }
    dst[i] = val;
}

// ---- peer-to-peer sum + Adam ------------------------------------------------------------------
struct P2pArgs {
    float* const* peer_grads; int* const* peer_flags; int* my_flags;
    int rank, R, step;
    long long n;
    float *p, *m, *v;
    float alpha, omb1, omb2, eps;
    unsigned int* counter; int* err;
};

__device__ __forceinline__ int ld_volatile_i32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// One launch per rank and step: (1) tell every peer that this rank's gradients are complete, (2) wait for theirs,
// (3) g = sum over ranks in rank order (identical on every rank, so the replicas stay bit-identical), read straight
// from the peers' HBM over NVLink, and the Keras-form Adam update on it, (4) tell every peer that their buffers have
// been read.  Waits are bounded: a dead peer raises *err instead of hanging the GPU.
__global__ void __launch_bounds__(256) p2p_sum_adam_kernel(const P2pArgs a) {
    __shared__ int ok;
    if (blockIdx.x == 0 && threadIdx.x < a.R) {
        __threadfence_system();
        *reinterpret_cast<volatile int*>(a.peer_flags[threadIdx.x] + a.rank) = a.step;          // ready[rank] at peer
    }
    if (threadIdx.x == 0) {
        int good = 1;
        for (int q = 0; q < a.R && good; ++q) {
            long long spins = 0;
            while (ld_volatile_i32(a.my_flags + q) < a.step) {
                if (++spins > 100000000ll) { good = 0; break; }
                __nanosleep(100);
            }
        }
        __threadfence_system();
        ok = good;
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(a.err) = 1;
        return;
    }
    const long long n4 = a.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 gg = reinterpret_cast<const float4*>(a.peer_grads[0])[i];
        for (int q = 1; q < a.R; ++q) {
            const float4 t = reinterpret_cast<const float4*>(a.peer_grads[q])[i];
            gg.x += t.x; gg.y += t.y; gg.z += t.z; gg.w += t.w;
        }
        float4 pp = reinterpret_cast<float4*>(a.p)[i];
        float4 mm = reinterpret_cast<float4*>(a.m)[i];
        float4 vv = reinterpret_cast<float4*>(a.v)[i];
#define PG_ADAM1(c)                                          \
        mm.c += (gg.c - mm.c) * a.omb1;                      \
        vv.c += (gg.c * gg.c - vv.c) * a.omb2;               \
        pp.c -= (mm.c * a.alpha) / (sqrtf(vv.c) + a.eps);
        PG_ADAM1(x) PG_ADAM1(y) PG_ADAM1(z) PG_ADAM1(w)
#undef PG_ADAM1
        reinterpret_cast<float4*>(a.p)[i] = pp;
        reinterpret_cast<float4*>(a.m)[i] = mm;
        reinterpret_cast<float4*>(a.v)[i] = vv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(a.counter, 1u) == gridDim.x - 1) {           // last block: every read of the peers is done
            *a.counter = 0u;
            for (int q = 0; q < a.R; ++q)
                *reinterpret_cast<volatile int*>(a.peer_flags[q] + a.R + a.rank) = a.step;      // done[rank] at peer
        }
    }
}

// before this rank overwrites its gradient buffer: every peer has finished reading the previous step's gradients
__global__ void p2p_wait_done_kernel(const int* my_flags, int R, int step, int* err) {
    if ((int)threadIdx.x < R) {
        long long spins = 0;
        while (ld_volatile_i32(my_flags + R + threadIdx.x) < step) {
            if (++spins > 100000000ll) { *reinterpret_cast<volatile int*>(err) = 1; break; }
            __nanosleep(100);
        }
    }
}

// ---- sharded peer-to-peer exchange fused with Adam (wide models: one launch per variable group) ----------------------
// The reduce-scatter + optimiser + all-gather of a data-parallel step as ONE kernel over peer memory (NVLink /
// NVSwitch): rank r OWNS the variables [v_lo, v_hi) of the group.  For its shard of every gradient slice (20 dense
// slices per group, + the codebook gradient without EMA) it reads the R partial gradients straight from the peers'
// HBM, sums them in rank order, applies the Keras-form Adam update to ITS fp32 master weights and moments, and writes
// what the other ranks compute with into the buffers of ALL ranks: the bf16 mirror of a kernel slice (bf16 mode; the
// fp32 master of a kernel stays with its owner until pgmvae_model_p2p_sync_state), the fp32 values of everything else
// (biases, fp32 / tf32 models, the codebook).  Per kernel parameter a rank moves 4 (R-1)/R bytes in and 2 (R-1)/R
// bytes out over NVLink (an all-reduce moves 8 (R-1)/R each way) and runs 1/R of the optimiser's HBM traffic.
// The CTAs are small (128 threads, 80 registers, no dynamic shared memory) so that one of them fits on every SM NEXT
// TO the resident CTA of the persistent GEMM / VQ kernels of the following group: the exchange takes issue slots and
// memory bandwidth, not SMs, from the compute it overlaps.
// Flag words (ints of the 256-byte block): [0,16) the whole-buffer exchange above; [16,24) ready2[q]: rank q's
// gradients of exchange `seq` are complete; [24,32) done2[q]: rank q has read this rank's gradients and written its
// shard of the parameters for exchange `seq`; [32,40) the barrier of the on-demand gather of the sharded state.
enum { SHARD_F32_ALL = 0, SHARD_BF16_ALL = 1, SHARD_BOTH_ALL = 2 };
struct ShardSlice { long long off, per_var; int mode; };
struct P2pShardArgs {
    float* const* peer_grads; float* const* peer_params; __nv_bfloat16* const* peer_wb; int* const* peer_flags; int* my_flags;
    float *m, *v;
    int rank, R, seq, nslices;
    long long v_lo, v_hi;
    ShardSlice s[24];
    float alpha, omb1, omb2, eps;
    unsigned int* counter; int* err;
};
constexpr int P2P_SHARD_THREADS = 128;
enum { P2P_READY2 = 16, P2P_DONE2 = 24, P2P_READY3 = 32 };

// The exchange streams GBs per group through an L2 that the GEMMs it runs next to rely on for their operand re-reads:
// every access carries an evict-first policy (and stays out of L1).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream_u2(uint2* p, const uint2& v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" :: "l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}

// RMAX >= R ranks, U elements of 16 bytes per thread and source in flight ((RMAX + 3) * U loads of 16 bytes: gradients of
// every rank, parameter, two moments; two ranks take the deeper unroll)
template <int RMAX, int U>
__global__ void __maxnreg__(80) p2p_shard_adam_kernel(const __grid_constant__ P2pShardArgs a) {
    __shared__ float* sg[8];
    __shared__ float* sp[8];
    __shared__ __nv_bfloat16* sw[8];
    __shared__ int ok;
    const int tid = threadIdx.x;
    if (tid < a.R) {
        sg[tid] = a.peer_grads[tid];
        sp[tid] = a.peer_params[tid];
        sw[tid] = a.peer_wb ? a.peer_wb[tid] : nullptr;
    }
    if (blockIdx.x == 0 && tid < a.R) {
        __threadfence_system();
        *reinterpret_cast<volatile int*>(a.peer_flags[tid] + P2P_READY2 + a.rank) = a.seq;
    }
    if (tid == 0) {
        int good = 1;
        for (int q = 0; q < a.R && good; ++q) {
            long long spins = 0;
            while (ld_volatile_i32(a.my_flags + P2P_READY2 + q) < a.seq) {
                if (++spins > 100000000ll) { good = 0; break; }
                __nanosleep(200);
            }
        }
        __threadfence_system();
        ok = good;
    }
    __syncthreads();
    if (!ok) {
        if (tid == 0) *reinterpret_cast<volatile int*>(a.err) = 1;
        return;
    }
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long gt = (long long)blockIdx.x * blockDim.x + tid;
    float* const pl = sp[a.rank];
    const uint64_t pol = l2_evict_first_policy();
    for (int si = 0; si < a.nslices; ++si) {
        const long long base4 = (a.s[si].off + a.v_lo * a.s[si].per_var) >> 2;
        const long long cnt4 = ((a.v_hi - a.v_lo) * a.s[si].per_var) >> 2;
        const bool f32_all = a.s[si].mode != SHARD_BF16_ALL, bf16_all = a.s[si].mode != SHARD_F32_ALL;
        for (long long i0 = gt; i0 < cnt4; i0 += nthreads * U) {
            // every load of the U elements goes out before the first use: one round trip per iteration, not one per source
            float4 t[U][RMAX], pp[U], mm[U], vv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long e = base4 + i0 + u * nthreads;
                if (i0 + u * nthreads < cnt4) {
#pragma unroll
                    for (int q = 0; q < RMAX; ++q)
                        if (q < a.R) t[u][q] = ld_stream_f4(reinterpret_cast<const float4*>(sg[q]) + e, pol);
                    pp[u] = ld_stream_f4(reinterpret_cast<const float4*>(pl) + e, pol);
                    mm[u] = ld_stream_f4(reinterpret_cast<const float4*>(a.m) + e, pol);
                    vv[u] = ld_stream_f4(reinterpret_cast<const float4*>(a.v) + e, pol);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long e = base4 + i0 + u * nthreads;
                if (i0 + u * nthreads >= cnt4) break;
                float4 gg = t[u][0];
#pragma unroll
                for (int q = 1; q < RMAX; ++q)
                    if (q < a.R) { gg.x += t[u][q].x; gg.y += t[u][q].y; gg.z += t[u][q].z; gg.w += t[u][q].w; }
#define PG_ADAM1(c)                                                \
                mm[u].c += (gg.c - mm[u].c) * a.omb1;              \
                vv[u].c += (gg.c * gg.c - vv[u].c) * a.omb2;       \
                pp[u].c -= (mm[u].c * a.alpha) / (sqrtf(vv[u].c) + a.eps);
                PG_ADAM1(x) PG_ADAM1(y) PG_ADAM1(z) PG_ADAM1(w)
#undef PG_ADAM1
                st_stream_f4(reinterpret_cast<float4*>(a.m) + e, mm[u], pol);
                st_stream_f4(reinterpret_cast<float4*>(a.v) + e, vv[u], pol);
                const __nv_bfloat162 lo = __floats2bfloat162_rn(pp[u].x, pp[u].y), hi = __floats2bfloat162_rn(pp[u].z, pp[u].w);
                const uint2 pk = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                if (!f32_all) st_stream_f4(reinterpret_cast<float4*>(pl) + e, pp[u], pol);      // the master stays with its owner
#pragma unroll
                for (int q = 0; q < RMAX; ++q)
                    if (q < a.R) {
                        if (f32_all) st_stream_f4(reinterpret_cast<float4*>(sp[q]) + e, pp[u], pol);
                        if (bf16_all) st_stream_u2(reinterpret_cast<uint2*>(sw[q]) + e, pk, pol);
                    }
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        if (atomicAdd(a.counter, 1u) == gridDim.x - 1) {           // last block: every read and write of the peers is done
            *a.counter = 0u;
            __threadfence_system();
            for (int q = 0; q < a.R; ++q)
                *reinterpret_cast<volatile int*>(a.peer_flags[q] + P2P_DONE2 + a.rank) = a.seq;
        }
    }
}

// wait until every rank has posted `seq` in the flag words flags[0..R)
__global__ void p2p_wait_flags_kernel(const int* flags, int R, int seq, int* err) {
    if ((int)threadIdx.x < R) {
        long long spins = 0;
        while (ld_volatile_i32(flags + threadIdx.x) < seq) {
            if (++spins > 100000000ll) { *reinterpret_cast<volatile int*>(err) = 1; break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
}

// The moments (and, in bf16 mode, the fp32 master of the kernels) live with the owner of a shard; checkpoints and
// the fp32 readers want them everywhere: every rank writes its shards into the buffers of all peers (on demand:
// pgmvae_model_p2p_sync_state)
struct P2pStateArgs {
    float* const* peer_p; float* const* peer_m; float* const* peer_v;
    int rank, R, nslices;
    long long v_lo, v_hi;
    ShardSlice s[24];
};
__global__ void __launch_bounds__(256) p2p_push_state_kernel(const __grid_constant__ P2pStateArgs a) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float* pl = a.peer_p[a.rank];
    const float* ml = a.peer_m[a.rank];
    const float* vl = a.peer_v[a.rank];
    for (int si = 0; si < a.nslices; ++si) {
        const long long base4 = (a.s[si].off + a.v_lo * a.s[si].per_var) >> 2;
        const long long cnt4 = ((a.v_hi - a.v_lo) * a.s[si].per_var) >> 2;
        const bool push_p = a.s[si].mode == SHARD_BF16_ALL;
        for (long long i = gt; i < cnt4; i += nthreads) {
            const float4 mm = reinterpret_cast<const float4*>(ml)[base4 + i];
            const float4 vv = reinterpret_cast<const float4*>(vl)[base4 + i];
            float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (push_p) pp = reinterpret_cast<const float4*>(pl)[base4 + i];
            for (int q = 0; q < a.R; ++q) {
                if (q == a.rank) continue;
                reinterpret_cast<float4*>(a.peer_m[q])[base4 + i] = mm;
                reinterpret_cast<float4*>(a.peer_v[q])[base4 + i] = vv;
                if (push_p) reinterpret_cast<float4*>(a.peer_p[q])[base4 + i] = pp;
            }
        }
    }
}
// every earlier write of this stream to the peers is complete (kernel boundary): tell them, and wait for theirs
__global__ void p2p_barrier_kernel(int* const* peer_flags, const int* my_flags, int word, int rank, int R, int seq, int* err) {
    if ((int)threadIdx.x < R) {
        __threadfence_system();
        *reinterpret_cast<volatile int*>(peer_flags[threadIdx.x] + word + rank) = seq;
        long long spins = 0;
        while (ld_volatile_i32(my_flags + word + threadIdx.x) < seq) {
            if (++spins > 100000000ll) { *reinterpret_cast<volatile int*>(err) = 1; break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
}

}  // namespace

struct pgmvae_model {
    pgmvae_ctx* ctx = nullptr;
    int V = 0, D = 0, K = 0, Vp = 0, Dp = 0;
    int units[4] = {0, 0, 0, 0};
    double cost = 0.25, decay = 0.99, epsilon = 1e-5;
    int ema = 1, max_batch = 0, Vg = 0;
    bool chain_ok = false;   // narrow enough for the TMEM-resident chain kernels (chain_tc.cu)
    Layer L[10];
    size_t n_dense = 0, e_off = 0, n_params = 0;   // floats
    float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
    float *ema_w = nullptr, *biased_w = nullptr, *stat_w = nullptr;    // [V][K][Dp]
    float *ema_c = nullptr, *biased_c = nullptr, *stat_c = nullptr;    // [V][K]
    int step_c = 0, step_w = 0;
    int64_t adam_t = 0;
    // workspace
    uint8_t* y_u8 = nullptr;
    float* yf = nullptr;
    float* H[10] = {};
    float* Gd[10] = {};
    float *q = nullptr, *st = nullptr;
    int32_t* idx = nullptr;
    double* acc = nullptr;          // device [4]
    double* acc_host = nullptr;     // pinned [4]
    unsigned long long *n1 = nullptr, *n0 = nullptr;
    int64_t device_bytes = 0;
    std::vector<void*> allocs;
    // data-parallel overlap: all-reduces run on their own stream behind events of the compute stream
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_compute = nullptr, ev_comm = nullptr, ev_stats = nullptr;
    // the ten wgrad launches of a step are independent of each other: they are spread over three streams so that
    // the tail of one overlaps the head of the next
    // peer-to-peer gradient exchange fused with Adam (single node, NVLink): every rank reads the gradient buffers
    // of all ranks directly and applies the identical update -- no NCCL launch on the critical path
    bool p2p = false;
    int p2p_rank = 0, p2p_n = 1, p2p_step = 0;
    float** peer_grads = nullptr;      // device array [R] of gradient buffers (own + IPC-mapped peers)
    int** peer_flags = nullptr;        // device array [R] of flag blocks: ready[R] | done[R]
    int* p2p_flags = nullptr;          // this rank's flag block (peers write into it)
    unsigned int* p2p_counter = nullptr;
    int* p2p_err = nullptr;            // pinned host word: a peer barrier timed out
    std::vector<void*> ipc_opened;
    // sharded exchange (wide models): the peers' parameter / mirror / moment buffers as well
    bool p2p_chain = false, p2p_shard = false, state_sharded = false;
    int p2p_seq = 0, p2p_seq3 = 0;
    float **peer_params = nullptr, **peer_m = nullptr, **peer_v = nullptr;
    __nv_bfloat16** peer_wb = nullptr;
    // stage 2 (count) walks the data in slabs larger than the training batch: nothing but codes and counts
    // leaves the SM, so the slab only needs its own copy of the data (uint8 + fp32)
    uint8_t* cnt_y8 = nullptr; float* cnt_yf = nullptr; int cnt_rows = 0;
    cudaStream_t aux_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    // the codebook update only needs the statistics of the forward pass: it runs on a side stream under the
    // weight-gradient kernel
    cudaStream_t ema_stream = nullptr;
    cudaEvent_t ev_ema_fork = nullptr, ev_ema = nullptr;
    // bf16 tensor-core mode (csrc/dense_bf16.cu): networks too wide for the chain kernels (cfg3).  Operands live in
    // bf16: the data matrix, every activation / pre-activation gradient of the current variable group, and a bf16
    // mirror of the weights in the stored [V][in][out] orientation (K-major operand of the dgrad GEMMs, MN-major
    // operand of the forward GEMMs), written by the Adam kernel together with the fp32 master weights.
    bool bf16 = false, shadow_dirty = true;
    __nv_bfloat16* yb = nullptr;            // [max_batch][Vp]
    uint32_t* ybits = nullptr; int ldbits = 0;   // the same data bit-packed (targets of the fused MSE stage), [max_batch][ldbits]
    __nv_bfloat16* Hb[10] = {};             // activations (layer 4 = the latent stays fp32 in H[4])
    __nv_bfloat16* Gb[10] = {};             // d(loss)/d(pre-activation)
    __nv_bfloat16* stb = nullptr;           // straight-through output of the VQ layer
    __nv_bfloat16* wb = nullptr;            // bf16 mirror of the dense parameters, same offsets as `params`
    __nv_bfloat16* cnt_yb = nullptr;

    float* E() const { return params + e_off; }
    float* dE() const { return grads + e_off; }
};

namespace {

int dev_alloc(pgmvae_model* m, void** p, size_t bytes) {
    PG_TRY(pgmvae_malloc(m->ctx, bytes, p));
    m->allocs.push_back(*p);
    m->device_bytes += (int64_t)bytes;
    cudaError_t e = cudaMemsetAsync(*p, 0, bytes, m->ctx->stream);
    if (e != cudaSuccess) {
        pgmvae_set_error("memset: %s", cudaGetErrorString(e));
        return PGMVAE_ECUDA;
    }
    return PGMVAE_OK;
}

// ---- reference-layout <-> internal-layout conversion (host side) -------------------
struct TensorRef {
    enum Kind { KERNEL, BIAS, CODEBOOK, CODESIZE } kind;
    int layer = -1;
    float* base = nullptr;   // device pointer of the internal tensor
    int64_t ref_count = 0, int_count = 0;
};

int resolve(pgmvae_model* m, const char* name, TensorRef* t) {
    std::string s(name);
    float* pbase = m->params;
    if (s.rfind("grad.", 0) == 0) { pbase = m->grads; s = s.substr(5); }
    else if (s.rfind("adam_m.", 0) == 0) { pbase = m->adam_m; s = s.substr(7); }
    else if (s.rfind("adam_v.", 0) == 0) { pbase = m->adam_v; s = s.substr(7); }
    const int64_t V = m->V;
    if (s.size() >= 5 && s[0] == 'f' && s[1] == 'd' && s[2] >= '0' && s[2] <= '9' && s[3] == '.') {
        const int l = s[2] - '0';
        const Layer& L = m->L[l];
        const std::string f = s.substr(4);
        t->layer = l;
        if (f == "kernel") {
            t->kind = TensorRef::KERNEL;
            t->base = pbase + L.w_off;
            t->ref_count = V * L.ref_in * L.ref_out;
            t->int_count = V * (int64_t)L.pin * L.pout;
            return PGMVAE_OK;
        }
        if (f == "bias") {
            t->kind = TensorRef::BIAS;
            t->base = pbase + L.b_off;
            t->ref_count = V * L.ref_out;
            t->int_count = V * (int64_t)L.pout;
            return PGMVAE_OK;
        }
    }
    const int64_t cb = V * (int64_t)m->K * m->Dp;
    auto codebook = [&](float* base) {
        t->kind = TensorRef::CODEBOOK; t->base = base;
        t->ref_count = V * (int64_t)m->D * m->K; t->int_count = cb;
        return PGMVAE_OK;
    };
    auto codesize = [&](float* base) {
        t->kind = TensorRef::CODESIZE; t->base = base;
        t->ref_count = V * (int64_t)m->K; t->int_count = t->ref_count;
        return PGMVAE_OK;
    };
    if (s == "vq.embeddings") {
        if (pbase != m->params && m->ema) {
            pgmvae_set_error("tensor '%s': the EMA codebook has no gradient / Adam slots", name);
            return PGMVAE_EINVAL;
        }
        return codebook(pbase + m->e_off);
    }
    if (m->ema && pbase == m->params) {
        if (s == "vq.ema_w") return codebook(m->ema_w);
        if (s == "vq.biased_w") return codebook(m->biased_w);
        if (s == "vq.stat_w") return codebook(m->stat_w);
        if (s == "vq.ema_cluster_size") return codesize(m->ema_c);
        if (s == "vq.biased_c") return codesize(m->biased_c);
        if (s == "vq.stat_c") return codesize(m->stat_c);
    }
    pgmvae_set_error("unknown tensor name '%s'", name);
    return PGMVAE_EINVAL;
}

// maps a reference index j in [0, V-1) of net v to its expanded position in [0, V)
inline int expand_idx(int j, int v) { return j + (j >= v ? 1 : 0); }

void to_internal(const pgmvae_model* m, const TensorRef& t, const float* ref, float* in) {
    const int V = m->V;
    memset(in, 0, sizeof(float) * (size_t)t.int_count);
    if (t.kind == TensorRef::KERNEL) {
        const Layer& L = m->L[t.layer];
        for (int v = 0; v < V; ++v)
            for (int r = 0; r < L.ref_in; ++r) {
                const int ir = t.layer == 0 ? expand_idx(r, v) : r;
                const float* src = ref + ((size_t)v * L.ref_in + r) * L.ref_out;
                float* dst = in + ((size_t)v * L.pin + ir) * L.pout;
                if (t.layer == 9) for (int c = 0; c < L.ref_out; ++c) dst[expand_idx(c, v)] = src[c];
                else memcpy(dst, src, sizeof(float) * L.ref_out);
            }
    } else if (t.kind == TensorRef::BIAS) {
        const Layer& L = m->L[t.layer];
        for (int v = 0; v < V; ++v)
            for (int c = 0; c < L.ref_out; ++c)
                in[(size_t)v * L.pout + (t.layer == 9 ? expand_idx(c, v) : c)] = ref[(size_t)v * L.ref_out + c];
    } else if (t.kind == TensorRef::CODEBOOK) {
        for (int v = 0; v < V; ++v)
            for (int d = 0; d < m->D; ++d)
                for (int k = 0; k < m->K; ++k)
                    in[((size_t)v * m->K + k) * m->Dp + d] = ref[((size_t)v * m->D + d) * m->K + k];
    } else {
        memcpy(in, ref, sizeof(float) * (size_t)t.ref_count);
    }
}

void to_reference(const pgmvae_model* m, const TensorRef& t, const float* in, float* ref) {
    const int V = m->V;
    if (t.kind == TensorRef::KERNEL) {
        const Layer& L = m->L[t.layer];
        for (int v = 0; v < V; ++v)
            for (int r = 0; r < L.ref_in; ++r) {
                const int ir = t.layer == 0 ? expand_idx(r, v) : r;
                float* dst = ref + ((size_t)v * L.ref_in + r) * L.ref_out;
                const float* src = in + ((size_t)v * L.pin + ir) * L.pout;
                if (t.layer == 9) for (int c = 0; c < L.ref_out; ++c) dst[c] = src[expand_idx(c, v)];
                else memcpy(dst, src, sizeof(float) * L.ref_out);
            }
    } else if (t.kind == TensorRef::BIAS) {
        const Layer& L = m->L[t.layer];
        for (int v = 0; v < V; ++v)
            for (int c = 0; c < L.ref_out; ++c)
                ref[(size_t)v * L.ref_out + c] = in[(size_t)v * L.pout + (t.layer == 9 ? expand_idx(c, v) : c)];
    } else if (t.kind == TensorRef::CODEBOOK) {
        for (int v = 0; v < V; ++v)
            for (int d = 0; d < m->D; ++d)
                for (int k = 0; k < m->K; ++k)
                    ref[((size_t)v * m->D + d) * m->K + k] = in[((size_t)v * m->K + k) * m->Dp + d];
    } else {
        memcpy(ref, in, sizeof(float) * (size_t)t.ref_count);
    }
}

int upload_batch(pgmvae_model* m, const uint8_t* y, int on_device, int B, const uint8_t** y_dev) {
    cudaStream_t st = m->ctx->stream;
    if (on_device) {
        *y_dev = y;
    } else {
        PG_CUDA(cudaMemcpyAsync(m->y_u8, y, (size_t)B * m->V, cudaMemcpyHostToDevice, st));
        *y_dev = m->y_u8;
    }
    if (m->bf16) {
        PG_TRY(pg_y_to_bits(m->ctx, st, *y_dev, m->V, m->ybits, m->ldbits, B, m->V));
        return pg_y_to_bf16(m->ctx, st, *y_dev, m->V, m->yb, m->Vp, B, m->V);
    }
    return pgmvae_y_to_f32(m->ctx, st, *y_dev, m->V, m->yf, m->Vp, B, m->V);
}

// bf16 mirror of the weights, rebuilt from the fp32 master copy when that changed outside the optimiser (init,
// set_tensor); the Adam kernel keeps it current during training
int refresh_shadows(pgmvae_model* m) {
    if (!m->bf16 || !m->shadow_dirty) return PGMVAE_OK;
    PG_TRY(pg_flat_to_bf16(m->ctx, m->ctx->stream, m->params, m->wb, (int64_t)m->n_dense));
    m->shadow_dirty = false;
    return PGMVAE_OK;
}

// all-reduce `buf` on the communication stream once everything issued so far on the compute stream is done
int overlapped_allreduce(pgmvae_model* m, pgmvae_comm* comm, void* buf, int64_t n, int dtype) {
    cudaStream_t st = m->ctx->stream;
    if (!m->comm_stream) {
        // highest priority: the exchange kernels take the first SM slots that free up instead of queueing behind
        // the thousands of CTAs of the weight-gradient launch that was issued before them
        int prio_lo = 0, prio_hi = 0;
        PG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        PG_CUDA(cudaStreamCreateWithPriority(&m->comm_stream, cudaStreamNonBlocking, prio_hi));
        PG_CUDA(cudaEventCreateWithFlags(&m->ev_compute, cudaEventDisableTiming));
        PG_CUDA(cudaEventCreateWithFlags(&m->ev_comm, cudaEventDisableTiming));
        PG_CUDA(cudaEventCreateWithFlags(&m->ev_stats, cudaEventDisableTiming));
    }
    PG_CUDA(cudaEventRecord(m->ev_compute, st));
    PG_CUDA(cudaStreamWaitEvent(m->comm_stream, m->ev_compute, 0));
    return pg_comm_allreduce(comm, buf, n, dtype, m->comm_stream);
}

// a peer-to-peer exchange that timed out is fatal for this model: the replicas have diverged (some blocks skipped
// the update) and the flag protocol is out of step.  The error word lives in pinned host memory, so every entry point
// can look at it without synchronising.
int p2p_check(const pgmvae_model* m) {
    if (m->p2p_err && *reinterpret_cast<volatile int*>(m->p2p_err)) {
        pgmvae_set_error("peer-to-peer gradient exchange: a rank did not reach the barrier (timed out); the model's replicas "
                         "are no longer consistent -- re-create the model (PGMVAE_P2P=0 selects the NCCL exchange)");
        return PGMVAE_ENCCL;
    }
    return PGMVAE_OK;
}

// the parameter slices a variable group exchanges: per layer the kernel [V][pin][pout] and the bias [V][pout] (bf16
// mirror at the same offsets in bf16 mode), and the codebook gradient when the codebook is trained by Adam.  Every
// rank computes with the bf16 mirror of a kernel but with the fp32 values of a bias (and of everything in fp32 / tf32
// mode): that decides what the owner of a shard hands to the other ranks.
int shard_slices(const pgmvae_model* m, ShardSlice* s) {
    int n = 0;
    for (int l = 0; l < 10; ++l) {
        const Layer& L = m->L[l];
        s[n++] = ShardSlice{(long long)L.w_off, (long long)L.pin * L.pout, m->bf16 ? SHARD_BF16_ALL : SHARD_F32_ALL};
        s[n++] = ShardSlice{(long long)L.b_off, (long long)L.pout, m->bf16 ? SHARD_BOTH_ALL : SHARD_F32_ALL};
    }
    if (!m->ema) s[n++] = ShardSlice{(long long)m->e_off, (long long)m->K * m->Dp, SHARD_F32_ALL};
    return n;
}

bool use_chain(const pgmvae_model* m) {
    return !m->bf16 && m->chain_ok && getenv("PGMVAE_NO_CHAIN") == nullptr && m->ctx->precision != PGMVAE_PREC_FP32;
}

// forward stages fd[l0, l1) of variables [g0, g0 + Gn) as chain stages
void chain_fwd_stages(const pgmvae_model* m, int g0, int l0, int l1, PgChainArgs& a) {
    const int64_t MB = m->max_batch;
    for (int l = l0; l < l1; ++l) {
        const Layer& L = m->L[l];
        PgChainStage& S = a.st[l - l0];
        S = PgChainStage{};
        S.K = L.pin; S.pout = L.pout; S.k_valid = L.in; S.n_valid = L.out; S.b_mn = 1;
        S.kind = l == 9 ? PG_CHAIN_EPI_SIGMOID_MSE : PG_CHAIN_EPI_SELU;
        S.w = m->params + L.w_off + (size_t)g0 * L.pin * L.pout; S.w_gs = (int64_t)L.pin * L.pout; S.ldw = L.pout;
        S.bias = m->params + L.b_off + (size_t)g0 * L.pout; S.bias_gs = L.pout;
        S.outp = l == 9 ? m->Gd[9] : m->H[l]; S.out_gs = MB * L.pout; S.ldo = L.pout;
    }
    a.nst = l1 - l0;
}

void chain_common(const pgmvae_model* m, int g0, int Gn, int B, PgChainArgs& a) {
    a.G = Gn; a.g0 = g0; a.B = B; a.V = m->V; a.Vp = m->Vp; a.D = m->D; a.Dp = m->Dp; a.K = m->K;
    a.yf = m->yf; a.ldyf = m->Vp;
    a.E = m->E() + (size_t)g0 * m->K * m->Dp; a.e_gs = (int64_t)m->K * m->Dp;
    a.q = m->q; a.stq = m->st; a.zq_gs = (int64_t)m->max_batch * m->Dp; a.ldzq = m->Dp;
    a.idx = m->idx; a.idx_gs = B;
}

// the dgrad stages fd9 .. fd1 appended at a.st[j0 ...]
void chain_bwd_stages(const pgmvae_model* m, int g0, int j0, PgChainArgs& a) {
    const int64_t MB = m->max_batch;
    for (int j = 0; j < 9; ++j) {
        const int l = 9 - j;
        const Layer& L = m->L[l];
        const Layer& P = m->L[l - 1];
        PgChainStage& S = a.st[j0 + j];
        S = PgChainStage{};
        S.K = L.pout; S.pout = L.pin; S.k_valid = L.out; S.n_valid = L.in; S.b_mn = 0;
        S.kind = PG_CHAIN_EPI_DGRAD; S.add_commit = l == 5;
        S.w = m->params + L.w_off + (size_t)g0 * L.pin * L.pout; S.w_gs = (int64_t)L.pin * L.pout; S.ldw = L.pout;
        S.outp = m->Gd[l - 1]; S.out_gs = MB * P.pout; S.ldo = P.pout;
        S.aux = m->H[l - 1]; S.aux_gs = MB * P.pout; S.ldaux = P.pout;
    }
    a.nst = j0 + 9;
}

// encoder + assignment (+ PLL histogram when n1/n0 are given) in one launch
int chain_encode(pgmvae_model* m, int g0, int Gn, int B, const uint8_t* y_dev, unsigned long long* n1,
                 unsigned long long* n0, const float* yf = nullptr) {
    PgChainArgs a{};
    a.mode = PG_CHAIN_ENCODE;
    chain_fwd_stages(m, g0, 0, 5, a);
    for (int j = 0; j < 5; ++j) a.st[j].outp = nullptr;        // nothing but the codes leaves the SM
    chain_common(m, g0, Gn, B, a);
    a.vq_stage = 4;
    a.a0 = yf ? yf : m->yf; a.a0_gs = 0; a.lda0 = m->Vp; a.a0_cols = m->Vp;
    a.y8 = y_dev; a.ldy8 = m->V; a.n1 = n1; a.n0 = n0;
    if (n1) a.idx = nullptr;                                   // counting: the codes never leave the SM
    return pg_chain_launch(m->ctx, m->ctx->stream, a);
}

// bf16 mode: forward layers fd[l0, l1) of variables [g0, g0 + Gn) on the tensor cores (csrc/dense_bf16.cu); the
// latent (layer 4) is written in fp32 for the VQ step, everything else as the bf16 operand of the next GEMM
int bf16_forward_layers(pgmvae_model* m, int g0, int Gn, int B, int l0, int l1, const __nv_bfloat16* yb) {
    pgmvae_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t MB = m->max_batch;
    for (int l = l0; l < l1; ++l) {
        const Layer& L = m->L[l];
        const __nv_bfloat16* x = l == 0 ? yb : (l == 5 ? m->stb : m->Hb[l - 1]);
        const int ldx = l == 0 ? m->Vp : (l == 5 ? m->Dp : m->L[l - 1].pout);
        const int64_t x_gs = l == 0 ? 0 : MB * ldx;
        PG_TRY(pg_bf16_fwd(ctx, st, x, x_gs, ldx, m->wb + L.w_off + (size_t)g0 * L.pin * L.pout, (int64_t)L.pin * L.pout, L.pout,
                           m->params + L.b_off + (size_t)g0 * L.pout, L.pout, l == 4 ? nullptr : m->Hb[l], MB * L.pout, L.pout,
                           l == 4 ? m->H[4] : nullptr, MB * L.pout, L.pout, Gn, B, L.in, L.out, L.act, 1));
    }
    return PGMVAE_OK;
}

// encoder fd0..fd4 + assignment for variables [g0, g0+Gn) of the current batch (in m->yf / m->yb)
int encode_group(pgmvae_model* m, int g0, int Gn, int B, const __nv_bfloat16* yb = nullptr) {
    pgmvae_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    if (m->bf16) {
        PG_TRY(bf16_forward_layers(m, g0, Gn, B, 0, 5, yb ? yb : m->yb));
        return pg_vq_assign_f16(ctx, st, m->H[4], (int64_t)m->max_batch * m->Dp, m->Dp, m->E() + (size_t)g0 * m->K * m->Dp,
                                (int64_t)m->K * m->Dp, m->Dp, m->idx, B, nullptr, nullptr, nullptr, 0, nullptr, 0, 0, Gn, B,
                                m->D, m->K);
    }
    for (int l = 0; l < 5; ++l) {
        const Layer& L = m->L[l];
        const float* x = l == 0 ? m->yf : m->H[l - 1];
        const int64_t x_gs = l == 0 ? 0 : (int64_t)m->max_batch * m->L[l - 1].pout;
        const int ldx = l == 0 ? m->Vp : m->L[l - 1].pout;
        PG_TRY(pgmvae_dense_fwd(ctx, st, x, x_gs, ldx, m->params + L.w_off + (size_t)g0 * L.pin * L.pout,
                                (int64_t)L.pin * L.pout, L.pout, m->params + L.b_off + (size_t)g0 * L.pout, L.pout,
                                m->H[l], (int64_t)m->max_batch * L.pout, L.pout, Gn, B, L.in, L.out, L.act));
    }
    PG_TRY(pgmvae_vq_assign(ctx, st, m->H[4], (int64_t)m->max_batch * m->Dp, m->Dp, m->E() + (size_t)g0 * m->K * m->Dp,
                            (int64_t)m->K * m->Dp, m->Dp, m->idx, B, nullptr, nullptr, Gn, B, m->D, m->K));
    return PGMVAE_OK;
}

}  // namespace

// read-only view for gibbs.cu (the sub-net path reads the weights in place)
struct PgModelView {
    pgmvae_ctx* ctx;
    int V, Vp, D, Dp, K;
    const float* params;
    const float* E;
    struct { int in, out, pin, pout; size_t w_off, b_off; } L[5];
};
int pg_model_view(pgmvae_model* m, PgModelView* v) {
    PG_CHECK_ARG(m && v);
    PG_TRY(p2p_check(m));
    if (m->p2p && m->state_sharded && m->bf16) {
        pgmvae_set_error("the sub-net path reads the fp32 master weights, which are sharded over the data-parallel ranks; call "
                         "pgmvae_model_p2p_sync_state (VqVAE.sync_state) on EVERY rank first");
        return PGMVAE_EINVAL;
    }
    v->ctx = m->ctx; v->V = m->V; v->Vp = m->Vp; v->D = m->D; v->Dp = m->Dp; v->K = m->K;
    v->params = m->params; v->E = m->E();
    for (int l = 0; l < 5; ++l) {
        v->L[l].in = m->L[l].in; v->L[l].out = m->L[l].out; v->L[l].pin = m->L[l].pin; v->L[l].pout = m->L[l].pout;
        v->L[l].w_off = m->L[l].w_off; v->L[l].b_off = m->L[l].b_off;
    }
    return PGMVAE_OK;
}

extern "C" {

int pgmvae_model_create(pgmvae_ctx* ctx, const int* units4, int nvar, int dim, int k, double cost, double decay,
                        double epsilon, int ema, int max_batch, pgmvae_model** out) {
    PG_CHECK_ARG(ctx && units4 && out);
    PG_CHECK_ARG(nvar >= 2 && dim >= 1 && k >= 1 && max_batch >= 1);
    for (int i = 0; i < 4; ++i) PG_CHECK_ARG(units4[i] >= 1);
    PG_CUDA(cudaSetDevice(ctx->device));
    pgmvae_model* m = new pgmvae_model();
    m->ctx = ctx; m->V = nvar; m->D = dim; m->K = k;
    m->Vp = pg_round_up(nvar, 8); m->Dp = pg_round_up(dim, 8);
    memcpy(m->units, units4, sizeof(int) * 4);
    m->cost = cost; m->decay = decay; m->epsilon = epsilon; m->ema = ema ? 1 : 0; m->max_batch = max_batch;
    // core/model.py:21-36: V-1 -> u0 -> u1 -> u2 -> u3 -> D | D -> u3 -> u2 -> u1 -> u0 -> V-1
    const int chain[11] = {nvar - 1, units4[0], units4[1], units4[2], units4[3], dim,
                           units4[3], units4[2], units4[1], units4[0], nvar - 1};
    size_t off = 0;
    for (int l = 0; l < 10; ++l) {
        Layer& L = m->L[l];
        L.ref_in = chain[l]; L.ref_out = chain[l + 1];
        L.in = l == 0 ? nvar : chain[l];
        L.out = l == 9 ? nvar : chain[l + 1];
        L.pin = pg_round_up(L.in, 8); L.pout = pg_round_up(L.out, 8);
        L.act = l == 9 ? PGMVAE_ACT_SIGMOID : PGMVAE_ACT_SELU;
        L.w_off = off; off += (size_t)nvar * L.pin * L.pout;
        L.b_off = off; off += (size_t)nvar * L.pout;
    }
    m->n_dense = off;
    m->e_off = off;
    off += (size_t)nvar * k * m->Dp;
    m->n_params = off;
    {
        int pin[10], pout[10];
        for (int l = 0; l < 10; ++l) { pin[l] = m->L[l].pin; pout[l] = m->L[l].pout; }
        m->chain_ok = pg_chain_supported(pin, pout, 10, m->Vp, m->Dp, k, ctx->smem_optin);
    }

    int rc = PGMVAE_OK;
    auto A = [&](void** p, size_t bytes) { if (rc == PGMVAE_OK) rc = dev_alloc(m, p, bytes); };
    const size_t trainable = m->ema ? m->n_dense : m->n_params;
    A((void**)&m->params, m->n_params * 4);
    A((void**)&m->grads, trainable * 4);
    A((void**)&m->adam_m, trainable * 4);
    A((void**)&m->adam_v, trainable * 4);
    const size_t cb = (size_t)nvar * k * m->Dp, cs = (size_t)nvar * k;
    if (m->ema) {
        A((void**)&m->ema_w, cb * 4); A((void**)&m->biased_w, cb * 4); A((void**)&m->stat_w, cb * 4);
        A((void**)&m->ema_c, cs * 4); A((void**)&m->biased_c, cs * 4); A((void**)&m->stat_c, cs * 4);
    }
    // bf16 mode: decided once, at creation (the operand buffers differ)
    m->bf16 = ctx->precision == PGMVAE_PREC_BF16 && (!m->chain_ok || getenv("PGMVAE_NO_CHAIN") != nullptr);
    // group size: keep the per-group activation workspace under ~12 GB (PGMVAE_WS_GB), or PGMVAE_GROUP_VARS variables
    size_t per_vb = 0;   // bytes per (variable, sample)
    if (m->bf16) {
        for (int l = 0; l < 9; ++l) per_vb += (size_t)m->L[l].pout * (l == 4 ? 4 : 2);     // Hb_0..Hb_8 (latent fp32)
        for (int l = 0; l < 10; ++l) per_vb += (size_t)m->L[l].pout * 2;                    // Gb_0..Gb_9
        per_vb += (size_t)m->Dp * 6 + 4;                                                     // q (fp32), st (bf16), idx
    } else {
        for (int l = 0; l < 9; ++l) per_vb += (size_t)m->L[l].pout * 4;       // H_0..H_8
        for (int l = 0; l < 10; ++l) per_vb += (size_t)m->L[l].pout * 4;      // Gd_0..Gd_9
        per_vb += (2 * (size_t)m->Dp + 1) * 4;                                // q, st, idx
    }
    size_t budget = (size_t)12 << 30;
    if (const char* ev = getenv("PGMVAE_WS_GB")) budget = atof(ev) > 0.0 ? (size_t)(atof(ev) * (double)(1ull << 30)) : budget;
    size_t vg = budget / (per_vb * (size_t)max_batch);
    // one group = one SM count of variables when the budget allows more: every layer's tile count is then a whole
    // number of waves of the persistent GEMMs, and under data parallelism more (smaller) groups leave less of the
    // gradient exchange exposed behind the last one
    if (vg > (size_t)ctx->sm_count) vg = ctx->sm_count;
    if (const char* ev = getenv("PGMVAE_GROUP_VARS")) vg = atoi(ev) > 0 ? (size_t)atoi(ev) : vg;
    if (vg < 1) vg = 1;
    if (vg > (size_t)nvar) vg = nvar;
    m->Vg = (int)vg;
    A((void**)&m->y_u8, (size_t)max_batch * nvar);
    if (m->bf16) {
        A((void**)&m->yb, (size_t)max_batch * m->Vp * 2);
        m->ldbits = m->Vp / 32 + 2;
        A((void**)&m->ybits, (size_t)max_batch * m->ldbits * 4);
        for (int l = 0; l < 10; ++l) {
            if (l < 9 && l != 4) A((void**)&m->Hb[l], vg * max_batch * m->L[l].pout * 2);
            A((void**)&m->Gb[l], vg * max_batch * m->L[l].pout * 2);
        }
        A((void**)&m->wb, m->n_dense * 2);
        A((void**)&m->H[4], vg * max_batch * m->Dp * 4);
        A((void**)&m->stb, vg * max_batch * m->Dp * 2);
    } else {
        A((void**)&m->yf, (size_t)max_batch * m->Vp * 4);
        for (int l = 0; l < 10; ++l) {
            if (l < 9) A((void**)&m->H[l], vg * max_batch * m->L[l].pout * 4);
            A((void**)&m->Gd[l], vg * max_batch * m->L[l].pout * 4);
        }
        A((void**)&m->st, vg * max_batch * m->Dp * 4);
    }
    A((void**)&m->q, vg * max_batch * m->Dp * 4);
    // idx holds all V variables of a batch (encode / count expose [V,B])
    A((void**)&m->idx, (size_t)nvar * max_batch * 4);
    A((void**)&m->acc, 4 * sizeof(double));
    A((void**)&m->n1, cs * 8);
    A((void**)&m->n0, cs * 8);
    if (rc == PGMVAE_OK && cudaMallocHost((void**)&m->acc_host, 4 * sizeof(double)) != cudaSuccess) {
        pgmvae_set_error("cudaMallocHost failed");
        rc = PGMVAE_ECUDA;
    }
    if (rc != PGMVAE_OK) {
        pgmvae_model_destroy(m);
        return rc;
    }
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = m;
    return PGMVAE_OK;
}

int pgmvae_model_destroy(pgmvae_model* m) {
    if (!m) return PGMVAE_OK;
    cudaStreamSynchronize(m->ctx->stream);
    for (void* p : m->allocs) cudaFree(p);
    if (m->acc_host) cudaFreeHost(m->acc_host);
    for (void* q : m->ipc_opened) cudaIpcCloseMemHandle(q);
    if (m->p2p_err) cudaFreeHost(m->p2p_err);
    if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
    for (int i = 0; i < 2; ++i) {
        if (m->aux_stream[i]) cudaStreamDestroy(m->aux_stream[i]);
        if (m->ev_join[i]) cudaEventDestroy(m->ev_join[i]);
    }
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ema_stream) cudaStreamDestroy(m->ema_stream);
    if (m->ev_ema_fork) cudaEventDestroy(m->ev_ema_fork);
    if (m->ev_ema) cudaEventDestroy(m->ev_ema);
    if (m->ev_compute) cudaEventDestroy(m->ev_compute);
    if (m->ev_comm) cudaEventDestroy(m->ev_comm);
    if (m->ev_stats) cudaEventDestroy(m->ev_stats);
    delete m;
    return PGMVAE_OK;
}

int64_t pgmvae_model_device_bytes(pgmvae_model* m) { return m ? m->device_bytes : 0; }
int pgmvae_model_group_size(pgmvae_model* m) { return m ? m->Vg : 0; }

int pgmvae_model_init(pgmvae_model* m, uint64_t seed) {
    PG_CHECK_ARG(m != nullptr);
    pgmvae_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const double V = m->V;
    for (int l = 0; l < 10; ++l) {
        const Layer& L = m->L[l];
        // Keras fans on the reference shape [V, in, out]: fan_in = V*in, fan_out = V*out
        const double fan_in = V * L.ref_in, fan_out = V * L.ref_out;
        const float limit = (float)(l < 9 ? sqrt(6.0 / fan_in) : sqrt(6.0 / (fan_in + fan_out)));
        const long long n = (long long)m->V * L.pin * L.pout;
        init_uniform_kernel<<<(unsigned)pg_cdiv(n, 256), 256, 0, st>>>(m->params + L.w_off, n, L.in, L.out, L.pin,
                                                                      L.pout, l == 0, l == 9, limit,
                                                                      seed * 1000003ull + (unsigned long long)l);
        PG_LAUNCHED(ctx);
        PG_CUDA(cudaMemsetAsync(m->params + L.b_off, 0, (size_t)m->V * L.pout * 4, st));
    }
    {
        const float limit = (float)sqrt(3.0 / (V * m->D));
        const long long n = (long long)m->V * m->K * m->Dp;
        init_uniform_kernel<<<(unsigned)pg_cdiv(n, 256), 256, 0, st>>>(m->E(), n, m->K, m->D, m->K, m->Dp, 0, 0, limit,
                                                                      seed * 1000003ull + 100ull);
        PG_LAUNCHED(ctx);
    }
    const size_t trainable = m->ema ? m->n_dense : m->n_params;
    PG_CUDA(cudaMemsetAsync(m->adam_m, 0, trainable * 4, st));
    PG_CUDA(cudaMemsetAsync(m->adam_v, 0, trainable * 4, st));
    m->adam_t = 0;
    if (m->ema) {
        const size_t cb = (size_t)m->V * m->K * m->Dp, cs = (size_t)m->V * m->K;
        PG_CUDA(cudaMemcpyAsync(m->ema_w, m->E(), cb * 4, cudaMemcpyDeviceToDevice, st));   // core/quantizer.py:117
        PG_CUDA(cudaMemsetAsync(m->biased_w, 0, cb * 4, st));
        PG_CUDA(cudaMemsetAsync(m->ema_c, 0, cs * 4, st));
        PG_CUDA(cudaMemsetAsync(m->biased_c, 0, cs * 4, st));
        m->step_c = m->step_w = 0;
    }
    PG_CUDA(cudaStreamSynchronize(st));
    m->shadow_dirty = true;
    return PGMVAE_OK;
}

int pgmvae_model_tensor_size(pgmvae_model* m, const char* name, int64_t* count) {
    PG_CHECK_ARG(m && name && count);
    TensorRef t;
    PG_TRY(resolve(m, name, &t));
    *count = t.ref_count;
    return PGMVAE_OK;
}

int pgmvae_model_set_tensor(pgmvae_model* m, const char* name, const float* host, int64_t count) {
    PG_CHECK_ARG(m && name && host);
    if (m->p2p && m->state_sharded) {
        // (the bf16 mirror would be refreshed from fp32 master weights that are stale on the ranks that do not own them)
        pgmvae_set_error("set_tensor '%s': the model's fp32 master weights are sharded over the data-parallel ranks; call "
                         "pgmvae_model_p2p_sync_state (VqVAE.sync_state) on EVERY rank first", name);
        return PGMVAE_EINVAL;
    }
    TensorRef t;
    PG_TRY(resolve(m, name, &t));
    if (count != t.ref_count) {
        pgmvae_set_error("set_tensor '%s': expected %lld elements, got %lld", name, (long long)t.ref_count,
                         (long long)count);
        return PGMVAE_EINVAL;
    }
    std::vector<float> in((size_t)t.int_count);
    to_internal(m, t, host, in.data());
    PG_CUDA(cudaStreamSynchronize(m->ctx->stream));
    PG_CUDA(cudaMemcpy(t.base, in.data(), sizeof(float) * (size_t)t.int_count, cudaMemcpyHostToDevice));
    m->shadow_dirty = true;
    return PGMVAE_OK;
}

int pgmvae_model_get_tensor(pgmvae_model* m, const char* name, float* host, int64_t count) {
    PG_CHECK_ARG(m && name && host);
    PG_TRY(p2p_check(m));
    if (m->p2p && m->state_sharded) {
        // after steps of the sharded exchange these tensors are complete only on the ranks that own their shards
        const std::string s(name);
        const bool moment = s.rfind("adam_m.", 0) == 0 || s.rfind("adam_v.", 0) == 0;
        const bool master = m->bf16 && s.rfind("fd", 0) == 0 && s.size() > 7 && s.compare(s.size() - 7, 7, ".kernel") == 0;
        if (moment || master) {
            pgmvae_set_error("get_tensor '%s': after data-parallel steps with the sharded peer-to-peer exchange this tensor is "
                             "complete only on the ranks that own its shards; call pgmvae_model_p2p_sync_state "
                             "(VqVAE.sync_state) on EVERY rank first", name);
            return PGMVAE_EINVAL;
        }
    }
    TensorRef t;
    PG_TRY(resolve(m, name, &t));
    if (count != t.ref_count) {
        pgmvae_set_error("get_tensor '%s': expected %lld elements, got %lld", name, (long long)t.ref_count,
                         (long long)count);
        return PGMVAE_EINVAL;
    }
    std::vector<float> in((size_t)t.int_count);
    PG_CUDA(cudaStreamSynchronize(m->ctx->stream));
    PG_CUDA(cudaMemcpy(in.data(), t.base, sizeof(float) * (size_t)t.int_count, cudaMemcpyDeviceToHost));
    to_reference(m, t, in.data(), host);
    return PGMVAE_OK;
}

/* handles_out: 6 x 64 bytes (cudaIpcMemHandle_t of the gradient buffer, this rank's flag block, the parameters, their
 * bf16 mirror (zeros without one), and the two Adam moment buffers) */
int pgmvae_model_p2p_export(pgmvae_model* m, void* handles_out) {
    PG_CHECK_ARG(m && handles_out);
    PG_CUDA(cudaSetDevice(m->ctx->device));
    if (!m->p2p_flags) {
        PG_TRY(dev_alloc(m, (void**)&m->p2p_flags, 256));
        PG_TRY(dev_alloc(m, (void**)&m->p2p_counter, 256));
        PG_CUDA(cudaMallocHost((void**)&m->p2p_err, sizeof(int)));
        *m->p2p_err = 0;
        PG_CUDA(cudaStreamSynchronize(m->ctx->stream));
    }
    cudaIpcMemHandle_t h[6];
    memset(h, 0, sizeof(h));
    PG_CUDA(cudaIpcGetMemHandle(&h[0], m->grads));
    PG_CUDA(cudaIpcGetMemHandle(&h[1], m->p2p_flags));
    PG_CUDA(cudaIpcGetMemHandle(&h[2], m->params));
    if (m->wb) PG_CUDA(cudaIpcGetMemHandle(&h[3], m->wb));
    PG_CUDA(cudaIpcGetMemHandle(&h[4], m->adam_m));
    PG_CUDA(cudaIpcGetMemHandle(&h[5], m->adam_v));
    memcpy(handles_out, h, sizeof(h));
    return PGMVAE_OK;
}

/* all_handles: nranks x 384 bytes in rank order (what every rank exported); at most 8 ranks on one node */
int pgmvae_model_p2p_import(pgmvae_model* m, int rank, int nranks, const void* all_handles) {
    PG_CHECK_ARG(m && all_handles && nranks >= 2 && nranks <= 8 && rank >= 0 && rank < nranks && m->p2p_flags);
    PG_CUDA(cudaSetDevice(m->ctx->device));
    void* ptr[6][8];
    void* own[6] = {m->grads, m->p2p_flags, m->params, m->wb, m->adam_m, m->adam_v};
    for (int q = 0; q < nranks; ++q) {
        if (q == rank) {
            for (int k = 0; k < 6; ++k) ptr[k][q] = own[k];
            continue;
        }
        cudaIpcMemHandle_t h[6];
        memcpy(h, (const char*)all_handles + (size_t)q * sizeof(h), sizeof(h));
        for (int k = 0; k < 6; ++k) {
            ptr[k][q] = nullptr;
            if (k == 3 && !m->wb) continue;
            PG_CUDA(cudaIpcOpenMemHandle(&ptr[k][q], h[k], cudaIpcMemLazyEnablePeerAccess));
            m->ipc_opened.push_back(ptr[k][q]);
        }
    }
    void** dst[6] = {(void**)&m->peer_grads, (void**)&m->peer_flags, (void**)&m->peer_params, (void**)&m->peer_wb,
                     (void**)&m->peer_m, (void**)&m->peer_v};
    for (int k = 0; k < 6; ++k) {
        PG_TRY(dev_alloc(m, dst[k], 8 * sizeof(void*)));
        PG_CUDA(cudaMemcpyAsync(*dst[k], ptr[k], nranks * sizeof(void*), cudaMemcpyHostToDevice, m->ctx->stream));
    }
    PG_CUDA(cudaStreamSynchronize(m->ctx->stream));
    m->p2p_rank = rank; m->p2p_n = nranks; m->p2p_step = 0; m->p2p_seq = 0; m->p2p_seq3 = 0; m->p2p = true;
    // which exchanges use the mapping.  Narrow models (chain kernels, a few MB of gradients): the whole-buffer sum +
    // Adam kernel, measured faster than NCCL at two ranks only (every rank reads every buffer); PGMVAE_P2P=1 forces it.
    // Wide models (per-group path): the sharded exchange at any rank count; PGMVAE_P2P_SHARD=0 keeps NCCL there.
    const char* want = getenv("PGMVAE_P2P");
    m->p2p_chain = want ? atoi(want) == 1 : nranks == 2;
    const char* ws = getenv("PGMVAE_P2P_SHARD");
    m->p2p_shard = !(ws && atoi(ws) == 0);
    return PGMVAE_OK;
}

/* give up the peer-to-peer exchange (a rank could not map its peers): the NCCL path is used instead */
int pgmvae_model_p2p_disable(pgmvae_model* m) {
    PG_CHECK_ARG(m != nullptr);
    m->p2p = false; m->p2p_chain = false; m->p2p_shard = false;
    return PGMVAE_OK;
}

/* the variables [*lo, *hi) of the group [g0, g0 + Gn) that rank `rank` of `nranks` owns in the sharded exchange: the shards
 * of a group are contiguous, disjoint, cover it, and differ by at most one variable (no CUDA call: testable on any host) */
int pgmvae_p2p_shard_bounds(int g0, int Gn, int rank, int nranks, int* lo, int* hi) {
    PG_CHECK_ARG(lo && hi && g0 >= 0 && Gn >= 0 && nranks >= 1 && rank >= 0 && rank < nranks);
    *lo = g0 + (int)((long long)rank * Gn / nranks);
    *hi = g0 + (int)((long long)(rank + 1) * Gn / nranks);
    return PGMVAE_OK;
}

/* 1 when steps of the sharded exchange have left the Adam moments (and, in bf16 mode, the fp32 master of the dense
 * kernels) complete only on the rank that owns a shard: sync_state() below completes them everywhere */
int pgmvae_model_p2p_state_sharded(pgmvae_model* m) { return m && m->p2p && m->state_sharded ? 1 : 0; }

/* COLLECTIVE over the ranks of the mapping: every rank writes its shards of the Adam moments (and fp32 master
 * kernels) into the buffers of all peers; afterwards every tensor reads (get_tensor) complete on every rank */
int pgmvae_model_p2p_sync_state(pgmvae_model* m) {
    PG_CHECK_ARG(m != nullptr);
    if (!m->p2p || !m->state_sharded) return PGMVAE_OK;
    PG_TRY(p2p_check(m));
    pgmvae_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    PG_CUDA(cudaSetDevice(ctx->device));
    for (int g0 = 0; g0 < m->V; g0 += m->Vg) {
        const int Gn = std::min(m->Vg, m->V - g0);
        P2pStateArgs a{};
        a.peer_p = m->peer_params; a.peer_m = m->peer_m; a.peer_v = m->peer_v; a.rank = m->p2p_rank; a.R = m->p2p_n;
        {
            int lo = 0, hi = 0;
            PG_TRY(pgmvae_p2p_shard_bounds(g0, Gn, m->p2p_rank, m->p2p_n, &lo, &hi));
            a.v_lo = lo; a.v_hi = hi;
        }
        a.nslices = shard_slices(m, a.s);
        if (a.v_hi <= a.v_lo) continue;
        PG_KERNEL(ctx, st, "p2p_push_state", 0.0, 0.0);
        p2p_push_state_kernel<<<ctx->sm_total * 2, 256, 0, st>>>(a);
        PG_LAUNCHED(ctx);
    }
    p2p_barrier_kernel<<<1, 32, 0, st>>>(m->peer_flags, m->p2p_flags, P2P_READY3, m->p2p_rank, m->p2p_n, ++m->p2p_seq3, m->p2p_err);
    PG_LAUNCHED(ctx);
    PG_CUDA(cudaStreamSynchronize(st));
    PG_TRY(p2p_check(m));
    m->state_sharded = false;
    return PGMVAE_OK;
}

int pgmvae_model_set_ema_steps(pgmvae_model* m, int step_c, int step_w) {
    PG_CHECK_ARG(m && step_c >= 0 && step_w >= 0);
    m->step_c = step_c; m->step_w = step_w;
    return PGMVAE_OK;
}
int pgmvae_model_set_adam_step(pgmvae_model* m, int64_t t) {
    PG_CHECK_ARG(m && t >= 0);
    m->adam_t = t;
    return PGMVAE_OK;
}

}  // extern "C"

namespace {
enum { STEP_NO_UPDATE = 1, STEP_FWD_ONLY = 2, STEP_NO_EMA = 4 };

int run_step(pgmvae_model* m, const uint8_t* y, int y_on_device, int B, int global_B, float lr, pgmvae_comm* comm,
             int flags, float* out_dev, double* metrics4) {
    PG_CHECK_ARG(m && y);
    PG_CHECK_ARG(B >= 1 && B <= m->max_batch);
    PG_TRY(p2p_check(m));
    if (global_B <= 0) global_B = B;
    PG_CHECK_ARG(global_B >= B);
    pgmvae_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    PG_CUDA(cudaSetDevice(ctx->device));
    const int V = m->V, D = m->D, K = m->K, Dp = m->Dp;
    const uint8_t* y_dev = nullptr;
    PG_TRY(upload_batch(m, y, y_on_device, B, &y_dev));

    const size_t trainable = m->ema ? m->n_dense : m->n_params;
    if (m->p2p && m->p2p_step > 0) {
        p2p_wait_done_kernel<<<1, 32, 0, st>>>(m->p2p_flags, m->p2p_n, m->p2p_step, m->p2p_err);
        PG_LAUNCHED(ctx);
    }
    PG_TRY(refresh_shadows(m));
    // (bf16 mode: every gradient element is written exactly once per step, nothing accumulates into the buffer)
    if (!m->bf16) PG_CUDA(cudaMemsetAsync(m->grads, 0, trainable * 4, st));
    else if (!m->ema) PG_CUDA(cudaMemsetAsync(m->dE(), 0, (trainable - m->n_dense) * 4, st));   // the scatter accumulates
    PG_CUDA(cudaMemsetAsync(m->acc, 0, 4 * sizeof(double), st));
    if (m->ema) {
        PG_CUDA(cudaMemsetAsync(m->stat_w, 0, (size_t)V * K * Dp * 4, st));
        PG_CUDA(cudaMemsetAsync(m->stat_c, 0, (size_t)V * K * 4, st));
    }
    // global means: Keras mse over B*V*(V-1) outputs (run.py:61); VQ means over V*B*D
    const double n_out = (double)global_B * V * (V - 1);
    const double n_lat = (double)global_B * V * D;
    const float gscale = (float)(2.0 / n_out);
    const float cscale = (float)(m->cost * 2.0 / n_lat);

    const bool chain = use_chain(m) && out_dev == nullptr;
    bool overlapped = false, use_p2p = false, ema_done = false, ema_side = false;
    const bool do_update = !(flags & (STEP_NO_UPDATE | STEP_FWD_ONLY));
    // wide models under data parallelism on one node: per variable group, reduce-scatter + Adam + all-gather as one
    // kernel over peer memory (p2p_shard_adam_kernel) instead of NCCL all-reduces followed by a replicated Adam
    const bool use_shard = comm != nullptr && !chain && m->p2p && m->p2p_shard && do_update;
    if (do_update && !use_shard && m->p2p && m->state_sharded) {
        // a replicated optimiser step would update fp32 master weights that are stale on the ranks that do not own them
        pgmvae_set_error("train_step: the optimiser state of this model is sharded over the data-parallel ranks (sharded "
                         "peer-to-peer exchange); call pgmvae_model_p2p_sync_state (VqVAE.sync_state) on EVERY rank before "
                         "training it through another path");
        return PGMVAE_EINVAL;
    }
    const double b1 = 0.9, b2 = 0.999;
    float alpha = 0.f;
    if (do_update) {
        m->adam_t += 1;
        alpha = (float)((double)lr * sqrt(1.0 - pow(b2, (double)m->adam_t)) / (1.0 - pow(b1, (double)m->adam_t)));
    }
    ShardSlice shard_s[24];
    const int shard_n = use_shard ? shard_slices(m, shard_s) : 0;
    ctx->coresident = use_shard;
    auto ema_update = [&](cudaStream_t es) -> int {
        if (ema_done || (flags & (STEP_NO_UPDATE | STEP_FWD_ONLY)) || !m->ema || (flags & STEP_NO_EMA)) return PGMVAE_OK;
        ema_done = true;
        m->step_c += 1; m->step_w += 1;
        return pgmvae_ema_apply(ctx, es, m->stat_c, m->stat_w, m->biased_c, m->biased_w, m->ema_c, m->ema_w, m->E(), V, K,
                                Dp, Dp, m->decay, m->epsilon, m->step_c, 1);
    };
    for (int g0 = 0; g0 < V && chain; g0 += m->Vg) {
        const int Gn = std::min(m->Vg, V - g0);
        const int64_t MB = m->max_batch;
        const bool stats = m->ema && !(flags & STEP_NO_EMA);
        const bool fused = !(flags & STEP_FWD_ONLY) && getenv("PGMVAE_CHAIN_SPLIT") == nullptr;
        {   // forward chain: fd0..fd4, VQ (+ EMA statistics), fd5..fd9, loss and d(loss)/d(pre-activation)
            PgChainArgs a{};
            a.mode = PG_CHAIN_FWD;
            chain_fwd_stages(m, g0, 0, 10, a);
            if (fused) {        // forward and dgrad of the same rows in one pass (nothing global lies between them:
                                // the loss scale is a constant, the EMA statistics only feed the codebook update)
                a.mode = PG_CHAIN_TRAIN;
                chain_bwd_stages(m, g0, 10, a);
                a.z = m->H[4]; a.qv = m->q; a.cscale = cscale;
            }
            chain_common(m, g0, Gn, B, a);
            a.vq_stage = 4;
            a.a0 = m->yf; a.a0_gs = 0; a.lda0 = m->Vp; a.a0_cols = m->Vp;
            a.stat_c = stats ? m->stat_c + (size_t)g0 * K : nullptr;
            a.stat_w = stats ? m->stat_w + (size_t)g0 * K * Dp : nullptr;
            a.acc = m->acc; a.gscale = gscale;
            PG_TRY(pg_chain_launch(ctx, st, a));
        }
        // single variable group under data parallelism: every exchange is issued as soon as its operand is
        // final and runs on the communication stream under the kernels that follow
        const bool overlap = comm != nullptr && Gn == V;
        use_p2p = overlap && m->p2p && m->p2p_chain && !(flags & (STEP_NO_UPDATE | STEP_FWD_ONLY));
        if (overlap) {      // EMA statistics + loss accumulators: one fused NCCL launch
            PG_TRY(pg_comm_group_begin(comm));
            PG_TRY(overlapped_allreduce(m, comm, m->acc, 4, 1));
            if (stats) {
                PG_TRY(pg_comm_allreduce(comm, m->stat_w, (int64_t)V * K * Dp, 0, m->comm_stream));
                PG_TRY(pg_comm_allreduce(comm, m->stat_c, (int64_t)V * K, 0, m->comm_stream));
            }
            PG_TRY(pg_comm_group_end(comm));
            PG_CUDA(cudaEventRecord(m->ev_stats, m->comm_stream));
            overlapped = true;
            if (!ctx->profiling) PG_TRY(ema_update(m->comm_stream));      // right behind the exchanged statistics
        } else if (Gn == V && comm == nullptr && !ctx->profiling && m->ema && !(flags & (STEP_NO_UPDATE | STEP_FWD_ONLY | STEP_NO_EMA))) {
            // single GPU: the codebook update runs on a side stream under the kernels that follow
            if (!m->ema_stream) {
                int prio_lo = 0, prio_hi = 0;
                PG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
                PG_CUDA(cudaStreamCreateWithPriority(&m->ema_stream, cudaStreamNonBlocking, prio_hi));
                PG_CUDA(cudaEventCreateWithFlags(&m->ev_ema_fork, cudaEventDisableTiming));
                PG_CUDA(cudaEventCreateWithFlags(&m->ev_ema, cudaEventDisableTiming));
            }
            PG_CUDA(cudaEventRecord(m->ev_ema_fork, st));
            PG_CUDA(cudaStreamWaitEvent(m->ema_stream, m->ev_ema_fork, 0));
            PG_TRY(ema_update(m->ema_stream));
            PG_CUDA(cudaEventRecord(m->ev_ema, m->ema_stream));
            ema_side = true;
        }
        if (flags & STEP_FWD_ONLY) continue;
        if (!m->ema)    // q_latent_loss gradient wrt the codebook: 2 (q - z) / (V B D)   (core/quantizer.py:51)
            PG_TRY(pgmvae_vq_codebook_grad(ctx, st, m->H[4], m->q, MB * Dp, Dp, m->idx, B, m->dE() + (size_t)g0 * K * Dp,
                                           (int64_t)K * Dp, Dp, (float)(2.0 / n_lat), Gn, B, D, K));
        if (!fused) {   // backward chain: dgrad of fd9 .. fd1
            PgChainArgs a{};
            a.mode = PG_CHAIN_BWD;
            chain_bwd_stages(m, g0, 0, a);
            chain_common(m, g0, Gn, B, a);
            a.vq_stage = -1;
            a.a0 = m->Gd[9]; a.a0_gs = MB * m->L[9].pout; a.lda0 = m->L[9].pout; a.a0_cols = m->L[9].pout;
            a.z = m->H[4]; a.qv = m->q; a.cscale = cscale;
            PG_TRY(pg_chain_launch(ctx, st, a));
        }
        // weight gradients.  Without NCCL buckets to feed (single GPU, or the peer-to-peer exchange that reads the
        // whole gradient buffer at once) all ten GEMMs go out as ONE launch.
        {
            PgWgradProblem pr[10];
            for (int l = 9; l >= 0; --l) {
                const Layer& L = m->L[l];
                PgWgradProblem& q = pr[9 - l];
                q.x = l == 0 ? m->yf : (l == 5 ? m->st : m->H[l - 1]);
                q.ldx = l == 0 ? m->Vp : (l == 5 ? Dp : m->L[l - 1].pout);
                q.x_gs = l == 0 ? 0 : MB * q.ldx;
                q.dy = m->Gd[l]; q.dy_gs = MB * L.pout; q.lddy = L.pout;
                q.dw = m->grads + L.w_off + (size_t)g0 * L.pin * L.pout; q.dw_gs = (int64_t)L.pin * L.pout; q.lddw = L.pout;
                q.db = m->grads + L.b_off + (size_t)g0 * L.pout; q.db_gs = L.pout;
                q.G = Gn; q.B = B; q.in = L.in; q.out = L.out; q.zero_row_base = l == 0 ? g0 : -1;
            }
            // Data parallel through NCCL: one launch + ONE all-reduce of the whole gradient buffer behind it (a few MB:
            // latency-bound, ~the cost of the exposed last bucket) beats ten launches feeding three overlapped
            // buckets; PGMVAE_DP_BUCKETS=1 keeps the bucketed schedule (large models: bandwidth-bound exchange).
            const bool per_layer = getenv("PGMVAE_WGRAD_PER_LAYER") != nullptr;
            const bool buckets = overlap && !use_p2p && getenv("PGMVAE_DP_BUCKETS") != nullptr;
            if (!per_layer && !buckets && pg_dense_wgrad_multi_supported(pr, 10)) {
                PG_TRY(pg_dense_wgrad_multi_tc(ctx, st, pr, 10));
                if (overlap && !use_p2p) PG_TRY(overlapped_allreduce(m, comm, m->grads, (int64_t)trainable, 0));
                continue;
            }
        }
        if (!m->aux_stream[0]) {
            for (int i = 0; i < 2; ++i) {
                PG_CUDA(cudaStreamCreateWithFlags(&m->aux_stream[i], cudaStreamNonBlocking));
                PG_CUDA(cudaEventCreateWithFlags(&m->ev_join[i], cudaEventDisableTiming));
            }
            PG_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
        }
        PG_CUDA(cudaEventRecord(m->ev_fork, st));                 // the backward chain has been issued
        for (int i = 0; i < 2; ++i) PG_CUDA(cudaStreamWaitEvent(m->aux_stream[i], m->ev_fork, 0));
        size_t bucket_end = m->n_dense;
        for (int l = 9; l >= 0; --l) {
            const Layer& L = m->L[l];
            const float* x = l == 0 ? m->yf : (l == 5 ? m->st : m->H[l - 1]);
            const int ldx = l == 0 ? m->Vp : (l == 5 ? Dp : m->L[l - 1].pout);
            const int64_t x_gs = l == 0 ? 0 : MB * ldx;
            // (the per-kernel profiler wants every launch alone on the compute stream)
            cudaStream_t ws = (ctx->profiling || (9 - l) % 3 == 0) ? st : m->aux_stream[(9 - l) % 3 - 1];
            PG_TRY(pgmvae_dense_wgrad(ctx, ws, x, x_gs, ldx, m->Gd[l], MB * L.pout, L.pout,
                                      m->grads + L.w_off + (size_t)g0 * L.pin * L.pout, (int64_t)L.pin * L.pout, L.pout,
                                      m->grads + L.b_off + (size_t)g0 * L.pout, L.pout, Gn, B, L.in, L.out,
                                      l == 0 ? g0 : -1));
            // gradients leave in three buckets (layers 9-7, 6-4, 3-0; parameters are laid out in layer order),
            // each as soon as its last wgrad is issued: few launches, and only the last bucket is exposed
            if (l == 0 || (overlap && (l == 7 || l == 4))) {
                for (int i = 0; i < 2; ++i) {                     // join the side streams into the compute stream
                    PG_CUDA(cudaEventRecord(m->ev_join[i], m->aux_stream[i]));
                    PG_CUDA(cudaStreamWaitEvent(st, m->ev_join[i], 0));
                }
                if (overlap && !use_p2p) {
                    PG_TRY(overlapped_allreduce(m, comm, m->grads + L.w_off, (int64_t)(bucket_end - L.w_off), 0));
                    bucket_end = L.w_off;
                }
            }
        }
        if (overlap && !m->ema && !use_p2p)
            PG_TRY(overlapped_allreduce(m, comm, m->dE(), (int64_t)V * K * Dp, 0));
    }
    // ---- layer-by-layer path (networks too wide for the chain kernels, or exact fp32): one variable group at a time.
    // Under data parallelism group g's gradient slices (and EMA statistics) are all-reduced on the communication
    // stream while group g+1 computes; only the last group's exchange is exposed.
    bool group_overlap = false;
    auto exchange_group = [&](int g0, int Gn) -> int {
        if (!comm) return PGMVAE_OK;
        if (!m->comm_stream) {
            int prio_lo = 0, prio_hi = 0;
            PG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            PG_CUDA(cudaStreamCreateWithPriority(&m->comm_stream, cudaStreamNonBlocking, prio_hi));
            PG_CUDA(cudaEventCreateWithFlags(&m->ev_compute, cudaEventDisableTiming));
            PG_CUDA(cudaEventCreateWithFlags(&m->ev_comm, cudaEventDisableTiming));
            PG_CUDA(cudaEventCreateWithFlags(&m->ev_stats, cudaEventDisableTiming));
        }
        PG_CUDA(cudaEventRecord(m->ev_compute, st));
        PG_CUDA(cudaStreamWaitEvent(m->comm_stream, m->ev_compute, 0));
        PG_TRY(pg_comm_group_begin(comm));
        if (!(flags & STEP_FWD_ONLY) && !use_shard) {
            for (int l = 0; l < 10; ++l) {
                const Layer& L = m->L[l];
                PG_TRY(pg_comm_allreduce(comm, m->grads + L.w_off + (size_t)g0 * L.pin * L.pout, (int64_t)Gn * L.pin * L.pout, 0,
                                         m->comm_stream));
                PG_TRY(pg_comm_allreduce(comm, m->grads + L.b_off + (size_t)g0 * L.pout, (int64_t)Gn * L.pout, 0, m->comm_stream));
            }
            if (!m->ema)
                PG_TRY(pg_comm_allreduce(comm, m->dE() + (size_t)g0 * K * Dp, (int64_t)Gn * K * Dp, 0, m->comm_stream));
        }
        if (m->ema && !(flags & STEP_NO_EMA)) {
            PG_TRY(pg_comm_allreduce(comm, m->stat_w + (size_t)g0 * K * Dp, (int64_t)Gn * K * Dp, 0, m->comm_stream));
            PG_TRY(pg_comm_allreduce(comm, m->stat_c + (size_t)g0 * K, (int64_t)Gn * K, 0, m->comm_stream));
        }
        if (g0 + Gn >= V) PG_TRY(pg_comm_allreduce(comm, m->acc, 4, 1, m->comm_stream));      // loss accumulators: after the last group
        PG_TRY(pg_comm_group_end(comm));
        // (measured at two GPUs: issuing the exchange BEFORE the NCCL launch -- so that it starts right at the group
        // boundary -- is slower, 113 vs 109 ms per cfg3 step)
        if (use_shard) {
            P2pShardArgs a{};
            a.peer_grads = m->peer_grads; a.peer_params = m->peer_params; a.peer_wb = m->bf16 ? m->peer_wb : nullptr;
            a.peer_flags = m->peer_flags; a.my_flags = m->p2p_flags;
            a.m = m->adam_m; a.v = m->adam_v;
            a.rank = m->p2p_rank; a.R = m->p2p_n; a.seq = ++m->p2p_seq; a.nslices = shard_n;
            {
                int lo = 0, hi = 0;
                PG_TRY(pgmvae_p2p_shard_bounds(g0, Gn, m->p2p_rank, m->p2p_n, &lo, &hi));
                a.v_lo = lo; a.v_hi = hi;
            }
            for (int i = 0; i < shard_n; ++i) a.s[i] = shard_s[i];
            a.alpha = alpha; a.omb1 = (float)(1.0 - b1); a.omb2 = (float)(1.0 - b2); a.eps = 1e-7f;
            a.counter = m->p2p_counter; a.err = m->p2p_err;
            // algorithmic bytes per owned parameter: R gradient reads, p / m / v read and written locally, then 4 bytes
            // (fp32) or 2 bytes (bf16 mirror of a kernel) to each of the other ranks
            double bytes = 0.0, own = 0.0;
            for (int i = 0; i < shard_n; ++i) {
                const double n = (double)(a.v_hi - a.v_lo) * (double)shard_s[i].per_var;
                const double out = shard_s[i].mode == SHARD_BF16_ALL ? 2.0 : (shard_s[i].mode == SHARD_BOTH_ALL ? 6.0 : 4.0);
                bytes += n * (4.0 * m->p2p_n + 24.0 + out * m->p2p_n);
                own += n;
            }
            PG_KERNEL(ctx, m->comm_stream, "p2p_shard_adam", bytes, (10.0 + m->p2p_n) * own);
            // An SM holds ONE shared-memory configuration at a time: without this hint an exchange CTA that reaches an idle SM
            // first leaves it at a small carve-out, and the GEMM CTA (224 KB) has to wait for it to finish.
            static bool carve[16] = {};
            if (!carve[ctx->device & 15]) {
                PG_CUDA(cudaFuncSetAttribute(p2p_shard_adam_kernel<2, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                PG_CUDA(cudaFuncSetAttribute(p2p_shard_adam_kernel<4, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                PG_CUDA(cudaFuncSetAttribute(p2p_shard_adam_kernel<8, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                carve[ctx->device & 15] = true;
            }
            int xg = ctx->sm_total;
            if (const char* ev = getenv("PGMVAE_P2P_CTAS")) xg = std::max(1, atoi(ev));        // (tuning: CTAs of the exchange kernel)
            if (m->p2p_n <= 2) p2p_shard_adam_kernel<2, 2><<<xg, P2P_SHARD_THREADS, 0, m->comm_stream>>>(a);
            else if (m->p2p_n <= 4) p2p_shard_adam_kernel<4, 1><<<xg, P2P_SHARD_THREADS, 0, m->comm_stream>>>(a);
            else p2p_shard_adam_kernel<8, 1><<<xg, P2P_SHARD_THREADS, 0, m->comm_stream>>>(a);
            PG_LAUNCHED(ctx);
            m->state_sharded = true;
        }
        group_overlap = true;
        return PGMVAE_OK;
    };
    for (int g0 = 0; g0 < V && !chain && m->bf16; g0 += m->Vg) {
        // bf16 tensor-core path (csrc/dense_bf16.cu, vq_tc16.cu)
        const int Gn = std::min(m->Vg, V - g0);
        const int64_t MB = m->max_batch;
        const int64_t zgs = MB * Dp;
        const bool stats = m->ema && !(flags & STEP_NO_EMA);
        PG_TRY(bf16_forward_layers(m, g0, Gn, B, 0, 5, m->yb));
        // VQ: assignment + EMA statistics + quantise + commitment loss in one kernel (core/quantizer.py:135-156)
        PG_TRY(pg_vq_assign_f16(ctx, st, m->H[4], zgs, Dp, m->E() + (size_t)g0 * K * Dp, (int64_t)K * Dp, Dp, m->idx, B, nullptr,
                                nullptr, stats ? m->stat_c + (size_t)g0 * K : nullptr, K,
                                stats ? m->stat_w + (size_t)g0 * K * Dp : nullptr, (int64_t)K * Dp, Dp, Gn, B, D, K, m->q, m->stb,
                                zgs, Dp, m->acc + 2));
        if (!m->ema && !(flags & STEP_FWD_ONLY))
            PG_TRY(pgmvae_vq_codebook_grad(ctx, st, m->H[4], m->q, zgs, Dp, m->idx, B, m->dE() + (size_t)g0 * K * Dp,
                                           (int64_t)K * Dp, Dp, (float)(2.0 / n_lat), Gn, B, D, K));
        PG_TRY(bf16_forward_layers(m, g0, Gn, B, 5, 9, m->yb));
        {   // fd9 + loss + d(loss)/d(pre-activation)
            const Layer& L = m->L[9];
            PG_TRY(pg_bf16_fwd_sigmoid_mse(ctx, st, m->Hb[8], MB * m->L[8].pout, m->L[8].pout,
                                           m->wb + L.w_off + (size_t)g0 * L.pin * L.pout, (int64_t)L.pin * L.pout, L.pout,
                                           m->params + L.b_off + (size_t)g0 * L.pout, L.pout, m->ybits, m->ldbits, m->Gb[9], MB * L.pout,
                                           L.pout, out_dev ? out_dev + (size_t)g0 * MB * L.pout : nullptr, MB * L.pout, L.pout,
                                           m->acc, Gn, g0, B, L.in, V, gscale, 1));
        }
        for (int l = (flags & STEP_FWD_ONLY) ? -1 : 9; l >= 0; --l) {
            const Layer& L = m->L[l];
            const __nv_bfloat16* x = l == 0 ? m->yb : (l == 5 ? m->stb : m->Hb[l - 1]);
            const int ldx = l == 0 ? m->Vp : (l == 5 ? Dp : m->L[l - 1].pout);
            const int64_t x_gs = l == 0 ? 0 : MB * ldx;
            PG_TRY(pg_bf16_wgrad(ctx, st, x, x_gs, ldx, m->Gb[l], MB * L.pout, L.pout,
                                 m->grads + L.w_off + (size_t)g0 * L.pin * L.pout, (int64_t)L.pin * L.pout, L.pout,
                                 m->grads + L.b_off + (size_t)g0 * L.pout, L.pout, Gn, B, L.in, L.out, l == 0 ? g0 : -1, 0));
            if (l > 0) {
                const Layer& P = m->L[l - 1];
                const bool vqb = (l == 5);
                PG_TRY(pg_bf16_dgrad(ctx, st, m->Gb[l], MB * L.pout, L.pout, m->wb + L.w_off + (size_t)g0 * L.pin * L.pout,
                                     (int64_t)L.pin * L.pout, L.pout, vqb ? nullptr : m->Hb[l - 1], MB * P.pout, P.pout,
                                     vqb ? m->H[4] : nullptr, MB * P.pout, P.pout, vqb ? m->H[4] : nullptr, vqb ? m->q : nullptr,
                                     zgs, Dp, cscale, m->Gb[l - 1], MB * P.pout, P.pout, nullptr, 0, 0, Gn, B, L.in, L.out,
                                     PGMVAE_ACT_SELU));
            }
        }
        PG_TRY(exchange_group(g0, Gn));
    }
    for (int g0 = 0; g0 < V && !chain && !m->bf16; g0 += m->Vg) {
        const int Gn = std::min(m->Vg, V - g0);
        PG_TRY(encode_group(m, g0, Gn, B));
        // m->idx rows [0, Gn) now hold this group's codes
        const float* Eg = m->E() + (size_t)g0 * K * Dp;
        const int64_t MB = m->max_batch;   // activation strides are fixed so that pad columns stay zero
        const int64_t zgs = MB * Dp;
        PG_TRY(pgmvae_vq_quantize(ctx, st, m->H[4], zgs, Dp, Eg, (int64_t)K * Dp, Dp, m->idx, B, m->q, m->st, zgs, Dp,
                                  m->acc + 2, Gn, B, D, K));
        if (m->ema) {
            if (!(flags & STEP_NO_EMA))
            PG_TRY(pgmvae_ema_stats(ctx, st, m->H[4], zgs, Dp, m->idx, B, m->stat_c + (size_t)g0 * K, K,
                                    m->stat_w + (size_t)g0 * K * Dp, (int64_t)K * Dp, Dp, Gn, B, D, K));
        } else if (!(flags & STEP_FWD_ONLY)) {
            // q_latent_loss gradient wrt the codebook: 2 (q - z) / (V B D)   (core/quantizer.py:51)
            PG_TRY(pgmvae_vq_codebook_grad(ctx, st, m->H[4], m->q, zgs, Dp, m->idx, B, m->dE() + (size_t)g0 * K * Dp,
                                           (int64_t)K * Dp, Dp, (float)(2.0 / n_lat), Gn, B, D, K));
        }
        // decoder fd5..fd8
        for (int l = 5; l < 9; ++l) {
            const Layer& L = m->L[l];
            const float* x = l == 5 ? m->st : m->H[l - 1];
            const int ldx = l == 5 ? Dp : m->L[l - 1].pout;
            PG_TRY(pgmvae_dense_fwd(ctx, st, x, MB * ldx, ldx, m->params + L.w_off + (size_t)g0 * L.pin * L.pout,
                                    (int64_t)L.pin * L.pout, L.pout, m->params + L.b_off + (size_t)g0 * L.pout, L.pout,
                                    m->H[l], MB * L.pout, L.pout, Gn, B, L.in, L.out, L.act));
        }
        {   // fd9 + loss + d(loss)/d(pre-activation)
            const Layer& L = m->L[9];
            PG_TRY(pgmvae_dense_fwd_sigmoid_mse(
                ctx, st, m->H[8], MB * m->L[8].pout, m->L[8].pout,
                m->params + L.w_off + (size_t)g0 * L.pin * L.pout, (int64_t)L.pin * L.pout, L.pout,
                m->params + L.b_off + (size_t)g0 * L.pout, L.pout, m->yf, m->Vp, m->Gd[9], MB * L.pout, L.pout,
                out_dev ? out_dev + (size_t)g0 * MB * L.pout : nullptr, m->acc, Gn, g0, B, L.in, V, gscale));
        }
        // backward
        for (int l = (flags & STEP_FWD_ONLY) ? -1 : 9; l >= 0; --l) {
            const Layer& L = m->L[l];
            const float* x = l == 0 ? m->yf : (l == 5 ? m->st : m->H[l - 1]);
            const int ldx = l == 0 ? m->Vp : (l == 5 ? Dp : m->L[l - 1].pout);
            const int64_t x_gs = l == 0 ? 0 : MB * ldx;
            PG_TRY(pgmvae_dense_wgrad(ctx, st, x, x_gs, ldx, m->Gd[l], MB * L.pout, L.pout,
                                      m->grads + L.w_off + (size_t)g0 * L.pin * L.pout, (int64_t)L.pin * L.pout, L.pout,
                                      m->grads + L.b_off + (size_t)g0 * L.pout, L.pout, Gn, B, L.in, L.out,
                                      l == 0 ? g0 : -1));
            if (l > 0) {
                const Layer& P = m->L[l - 1];
                const bool vqb = (l == 5);
                PG_TRY(pgmvae_dense_dgrad(ctx, st, m->Gd[l], MB * L.pout, L.pout,
                                          m->params + L.w_off + (size_t)g0 * L.pin * L.pout, (int64_t)L.pin * L.pout,
                                          L.pout, m->H[l - 1], MB * P.pout, P.pout, vqb ? m->H[4] : nullptr,
                                          vqb ? m->q : nullptr, zgs, Dp, cscale, m->Gd[l - 1], MB * P.pout,
                                          P.pout, Gn, B, L.in, L.out, PGMVAE_ACT_SELU));
            }
        }
        PG_TRY(exchange_group(g0, Gn));
    }

    if (comm && overlapped) {
        // join: the statistics were exchanged long ago, so the codebook update runs while the gradient exchange is
        // still in flight; the optimiser waits for that one
        PG_CUDA(cudaStreamWaitEvent(st, m->ev_stats, 0));
        PG_TRY(ema_update(st));
        PG_CUDA(cudaEventRecord(m->ev_comm, m->comm_stream));
        PG_CUDA(cudaStreamWaitEvent(st, m->ev_comm, 0));
    } else if (comm && group_overlap) {
        // every group's exchange has been issued; the optimiser and the codebook update wait for the last one
        PG_CUDA(cudaEventRecord(m->ev_comm, m->comm_stream));
        PG_CUDA(cudaStreamWaitEvent(st, m->ev_comm, 0));
        if (use_shard) {
            // ... and, sharded exchange: until every peer has written its shards of the new parameters into this rank's
            // buffers and has read this rank's gradients (they are overwritten by the next step)
            p2p_wait_flags_kernel<<<1, 32, 0, st>>>(m->p2p_flags + P2P_DONE2, m->p2p_n, m->p2p_seq, m->p2p_err);
            PG_LAUNCHED(ctx);
        }
    } else if (comm) {
        if (!(flags & STEP_FWD_ONLY)) PG_TRY(pg_comm_allreduce(comm, m->grads, (int64_t)trainable, 0, st));
        if (m->ema && !(flags & STEP_NO_EMA)) {
            PG_TRY(pg_comm_allreduce(comm, m->stat_w, (int64_t)V * K * Dp, 0, st));
            PG_TRY(pg_comm_allreduce(comm, m->stat_c, (int64_t)V * K, 0, st));
        }
        PG_TRY(pg_comm_allreduce(comm, m->acc, 4, 1, st));
    }

    if (do_update && !use_shard) {
        if (use_p2p) {
            P2pArgs a{};
            a.peer_grads = m->peer_grads; a.peer_flags = m->peer_flags; a.my_flags = m->p2p_flags;
            a.rank = m->p2p_rank; a.R = m->p2p_n; a.step = ++m->p2p_step;
            a.n = (long long)trainable; a.p = m->params; a.m = m->adam_m; a.v = m->adam_v;
            a.alpha = alpha; a.omb1 = (float)(1.0 - b1); a.omb2 = (float)(1.0 - b2); a.eps = 1e-7f;
            a.counter = m->p2p_counter; a.err = m->p2p_err;
            int blocks = (int)std::min<int64_t>(pg_cdiv((int64_t)trainable, 256 * 4 * 2), (int64_t)ctx->sm_count * 4);
            if (blocks < 1) blocks = 1;
            PG_KERNEL(ctx, st, "p2p_sum_adam", 4.0 * trainable * (m->p2p_n + 6.0), (10.0 + m->p2p_n) * trainable);
            p2p_sum_adam_kernel<<<blocks, 256, 0, st>>>(a);
            PG_LAUNCHED(ctx);
        } else {
            PG_TRY(pg_adam_step_shadow(ctx, st, m->params, m->grads, m->adam_m, m->adam_v, (int64_t)trainable, alpha, b1, b2,
                                       1e-7, m->bf16 ? m->wb : nullptr, (int64_t)m->n_dense));
        }
        if (use_p2p) m->shadow_dirty = true;          // (the peer-to-peer kernel does not write the bf16 mirror)
    }
    PG_TRY(ema_update(st));
    if (ema_side) PG_CUDA(cudaStreamWaitEvent(st, m->ev_ema, 0));       // later work sees the new codebook
    if (metrics4) {
        PG_CUDA(cudaMemcpyAsync(m->acc_host, m->acc, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaStreamSynchronize(st));
        PG_TRY(p2p_check(m));
        const double mse = m->acc_host[0] / n_out, mae = m->acc_host[1] / n_out;
        const double e_latent = m->acc_host[2] / n_lat;
        const double vq = m->ema ? m->cost * e_latent : (1.0 + m->cost) * e_latent;
        metrics4[0] = mse + vq; metrics4[1] = mse; metrics4[2] = mae; metrics4[3] = vq;
    }
    return PGMVAE_OK;
}
}  // namespace

extern "C" {

int pgmvae_model_train_step(pgmvae_model* m, const uint8_t* y, int y_on_device, int B, int global_B, float lr,
                            pgmvae_comm* comm, int flags, double* metrics4) {
    return run_step(m, y, y_on_device, B, global_B, lr, comm, (flags & 1) ? STEP_NO_UPDATE : 0, nullptr, metrics4);
}

int pgmvae_model_forward(pgmvae_model* m, const uint8_t* y, int y_on_device, int B, int training, float* out_dev,
                         double* metrics4) {
    return run_step(m, y, y_on_device, B, B, 0.f, nullptr, STEP_FWD_ONLY | (training ? 0 : STEP_NO_EMA), out_dev,
                    metrics4);
}

int pgmvae_model_encode(pgmvae_model* m, const uint8_t* y, int y_on_device, int B, int32_t* idx_dev) {
    PG_CHECK_ARG(m && y && idx_dev);
    PG_CHECK_ARG(B >= 1 && B <= m->max_batch);
    PG_CUDA(cudaSetDevice(m->ctx->device));
    const uint8_t* y_dev = nullptr;
    PG_TRY(refresh_shadows(m));
    PG_TRY(upload_batch(m, y, y_on_device, B, &y_dev));
    for (int g0 = 0; g0 < m->V; g0 += m->Vg) {
        const int Gn = std::min(m->Vg, m->V - g0);
        if (use_chain(m)) PG_TRY(chain_encode(m, g0, Gn, B, y_dev, nullptr, nullptr));
        else PG_TRY(encode_group(m, g0, Gn, B));
        PG_CUDA(cudaMemcpyAsync(idx_dev + (size_t)g0 * B, m->idx, (size_t)Gn * B * 4, cudaMemcpyDeviceToDevice,
                                m->ctx->stream));
    }
    return PGMVAE_OK;
}

int pgmvae_model_arithmetic(pgmvae_model* m) {
    if (!m) return -1;
    if (m->bf16) return 2;
    return m->ctx->precision == PGMVAE_PREC_FP32 ? 0 : 1;
}

int pgmvae_model_count(pgmvae_model* m, const uint8_t* y, int y_on_device, int64_t N, unsigned long long* n1_host,
                       unsigned long long* n0_host) {
    PG_CHECK_ARG(m != nullptr);
    return pgmvae_model_count_vars(m, y, y_on_device, N, 0, m->V, n1_host, n0_host);
}

int pgmvae_model_count_vars(pgmvae_model* m, const uint8_t* y, int y_on_device, int64_t N, int v0, int v1,
                            unsigned long long* n1_host, unsigned long long* n0_host) {
    PG_CHECK_ARG(m && y && n1_host && n0_host && N >= 0);
    PG_TRY(pgmvae_model_count_begin(m));
    PG_TRY(pgmvae_model_count_add(m, y, y_on_device, N, v0, v1));
    return pgmvae_model_count_end(m, v0, v1, n1_host, n0_host);
}

int pgmvae_model_count_begin(pgmvae_model* m) {
    PG_CHECK_ARG(m != nullptr);
    PG_TRY(p2p_check(m));
    cudaStream_t st = m->ctx->stream;
    PG_CUDA(cudaSetDevice(m->ctx->device));
    const size_t cs = (size_t)m->V * m->K;
    PG_TRY(refresh_shadows(m));
    PG_CUDA(cudaMemsetAsync(m->n1, 0, cs * 8, st));
    PG_CUDA(cudaMemsetAsync(m->n0, 0, cs * 8, st));
    return PGMVAE_OK;
}

int pgmvae_model_count_add(pgmvae_model* m, const uint8_t* y, int y_on_device, int64_t N, int v0, int v1) {
    PG_CHECK_ARG(m && y && N >= 0);
    PG_CHECK_ARG(v0 >= 0 && v0 <= v1 && v1 <= m->V);
    pgmvae_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    PG_CUDA(cudaSetDevice(ctx->device));
    if (use_chain(m) && m->Vg >= m->V && N > m->max_batch && v0 == 0 && v1 == m->V) {
        // slabs of up to 32768 samples per launch: 86 tile triples per variable instead of 11, so the items divide
        // evenly over the SMs, and an eighth of the launches
        const int CB = (int)std::min<int64_t>(N, 32768);
        if (m->cnt_rows < CB) {
            PG_TRY(dev_alloc(m, (void**)&m->cnt_y8, (size_t)CB * m->V));
            PG_TRY(dev_alloc(m, (void**)&m->cnt_yf, (size_t)CB * m->Vp * 4));
            m->cnt_rows = CB;
        }
        for (int64_t s = 0; s < N; s += CB) {
            const int B = (int)std::min<int64_t>(CB, N - s);
            const uint8_t* y_dev = y + (size_t)s * m->V;
            if (!y_on_device) {
                PG_CUDA(cudaMemcpyAsync(m->cnt_y8, y_dev, (size_t)B * m->V, cudaMemcpyHostToDevice, st));
                y_dev = m->cnt_y8;
            }
            PG_TRY(pgmvae_y_to_f32(ctx, st, y_dev, m->V, m->cnt_yf, m->Vp, B, m->V));
            PG_TRY(chain_encode(m, 0, m->V, B, y_dev, m->n1, m->n0, m->cnt_yf));
        }
        return PGMVAE_OK;
    }
    for (int64_t s = 0; s < N; s += m->max_batch) {
        const int B = (int)std::min<int64_t>(m->max_batch, N - s);
        const uint8_t* y_dev = nullptr;
        PG_TRY(upload_batch(m, y + (size_t)s * m->V, y_on_device, B, &y_dev));
        for (int g0 = v0; g0 < v1; g0 += m->Vg) {
            const int Gn = std::min(m->Vg, v1 - g0);
            if (use_chain(m)) {      // encoder + assignment + histogram in one launch
                PG_TRY(chain_encode(m, g0, Gn, B, y_dev, m->n1 + (size_t)g0 * m->K, m->n0 + (size_t)g0 * m->K));
                continue;
            }
            PG_TRY(encode_group(m, g0, Gn, B));
            PG_TRY(pgmvae_pll_count(ctx, st, m->idx, B, y_dev, m->V, g0, m->n1 + (size_t)g0 * m->K,
                                    m->n0 + (size_t)g0 * m->K, Gn, B, m->K));
        }
    }
    return PGMVAE_OK;
}

int pgmvae_model_count_end(pgmvae_model* m, int v0, int v1, unsigned long long* n1_host, unsigned long long* n0_host) {
    PG_CHECK_ARG(m && n1_host && n0_host && v0 >= 0 && v0 <= v1 && v1 <= m->V);
    cudaStream_t st = m->ctx->stream;
    const size_t lo = (size_t)v0 * m->K, cnt = (size_t)(v1 - v0) * m->K;
    if (cnt) {
        PG_CUDA(cudaMemcpyAsync(n1_host + lo, m->n1 + lo, cnt * 8, cudaMemcpyDeviceToHost, st));
        PG_CUDA(cudaMemcpyAsync(n0_host + lo, m->n0 + lo, cnt * 8, cudaMemcpyDeviceToHost, st));
    }
    PG_CUDA(cudaStreamSynchronize(st));
    return PGMVAE_OK;
}

}  // extern "C"
