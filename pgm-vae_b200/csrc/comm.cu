// Data-parallel communicator: NCCL over NVLink 5 / NVSwitch, one rank per process.
// New work -- the reference is single-device (run.py:27-31).  Exchanges per training step
// (SURVEY.md 8e): sum-allreduce of the dense gradients, of the EMA statistics
// counts [V,K] + dw [V,K,D] BEFORE the EMA update (so that every rank writes the same
// codebook), and of the loss accumulators; PLL counts are reduced once per evaluation.
//
// libnccl is resolved at run time (dlopen) so that libpgmvae.so loads on machines without
// NCCL; PGMVAE_NCCL_LIB overrides the library path (the Python host points it at the
// NCCL bundled with torch).
#include <dlfcn.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { NCCL_UINT64 = 5, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return PGMVAE_OK;
    const char* env = getenv("PGMVAE_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        pgmvae_set_error("cannot load libnccl (%s); set PGMVAE_NCCL_LIB", dlerror());
        return PGMVAE_ENCCL;
    }
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce =
        (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
        pgmvae_set_error("libnccl lacks a required symbol");
        dlclose(h);
        return PGMVAE_ENCCL;
    }
    g_nccl.handle = h;
    return PGMVAE_OK;
}

int nccl_fail(const char* what, int rc) {
    pgmvae_set_error("%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
    return PGMVAE_ENCCL;
}

}  // namespace

struct pgmvae_comm {
    pgmvae_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
};

int pg_comm_allreduce(pgmvae_comm* c, void* buf, int64_t n, int dtype, cudaStream_t st) {
    if (!c || c->nranks <= 1 || n <= 0) return PGMVAE_OK;
    const int dt = dtype == 0 ? NCCL_FLOAT32 : (dtype == 1 ? NCCL_FLOAT64 : NCCL_UINT64);
    const int rc = g_nccl.AllReduce(buf, buf, (size_t)n, dt, NCCL_SUM, c->comm, st);
    if (rc != ncclSuccess_) return nccl_fail("ncclAllReduce", rc);
    return PGMVAE_OK;
}

// several all-reduces fused into one NCCL launch (no-ops without a communicator / symbols)
int pg_comm_group_begin(pgmvae_comm* c) {
    if (!c || c->nranks <= 1 || !g_nccl.GroupStart) return PGMVAE_OK;
    const int rc = g_nccl.GroupStart();
    return rc == ncclSuccess_ ? PGMVAE_OK : nccl_fail("ncclGroupStart", rc);
}
int pg_comm_group_end(pgmvae_comm* c) {
    if (!c || c->nranks <= 1 || !g_nccl.GroupEnd) return PGMVAE_OK;
    const int rc = g_nccl.GroupEnd();
    return rc == ncclSuccess_ ? PGMVAE_OK : nccl_fail("ncclGroupEnd", rc);
}

extern "C" {

int pgmvae_comm_unique_id(void* out128) {
    PG_CHECK_ARG(out128 != nullptr);
    PG_TRY(load_nccl());
    ncclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc != ncclSuccess_) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(out128, &id, 128);
    return PGMVAE_OK;
}

int pgmvae_comm_create(pgmvae_ctx* ctx, int rank, int nranks, const void* id128, pgmvae_comm** out) {
    PG_CHECK_ARG(ctx && out && nranks >= 1 && rank >= 0 && rank < nranks);
    PG_CHECK_ARG(nranks == 1 || id128 != nullptr);
    if (nranks > 1) {
        PG_TRY(load_nccl());
        PG_CUDA(cudaSetDevice(ctx->device));
    }
    pgmvae_comm* c = new pgmvae_comm();        // (nothing below returns without deleting it)
    c->ctx = ctx; c->rank = rank; c->nranks = nranks;
    if (nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        const int rc = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
        if (rc != ncclSuccess_) {
            delete c;
            return nccl_fail("ncclCommInitRank", rc);
        }
    }
    *out = c;
    return PGMVAE_OK;
}

int pgmvae_comm_destroy(pgmvae_comm* c) {
    if (!c) return PGMVAE_OK;
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return PGMVAE_OK;
}

int pgmvae_comm_allreduce_f32(pgmvae_comm* c, float* buf, int64_t n, void* stream) {
    PG_CHECK_ARG(c && buf);
    return pg_comm_allreduce(c, buf, n, 0, pg_stream(c->ctx, stream));
}
int pgmvae_comm_allreduce_f64(pgmvae_comm* c, double* buf, int64_t n, void* stream) {
    PG_CHECK_ARG(c && buf);
    return pg_comm_allreduce(c, buf, n, 1, pg_stream(c->ctx, stream));
}
int pgmvae_comm_allreduce_u64(pgmvae_comm* c, unsigned long long* buf, int64_t n, void* stream) {
    PG_CHECK_ARG(c && buf);
    return pg_comm_allreduce(c, buf, n, 2, pg_stream(c->ctx, stream));
}

}  // extern "C"
