// Context, memory and error plumbing of the C-ABI (include/pgmvae.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void pgmvae_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void pg_prof_begin(pgmvae_ctx* ctx, cudaStream_t st, const char* name, double bytes, double flops) {
    pg_prof_rec r{name, nullptr, nullptr, st, bytes, flops};
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    ctx->prof.push_back(r);
    ctx->prof_open = true;
}
void pg_prof_end(pgmvae_ctx* ctx) {
    if (ctx->prof.empty() || !ctx->prof_open) return;
    ctx->prof_open = false;
    pg_prof_rec& r = ctx->prof.back();
    cudaEventRecord(r.e1, r.st);
}

extern "C" {

int pgmvae_ctx_profile_begin(pgmvae_ctx* ctx) {
    PG_CHECK_ARG(ctx != nullptr);
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->prof.clear();
    ctx->profiling = true;
    return PGMVAE_OK;
}

// JSON: [{"name":..,"launches":n,"ms":total,"bytes":total,"flops":total}, ...]
int pgmvae_ctx_profile_end(pgmvae_ctx* ctx, char* json_out, size_t cap) {
    PG_CHECK_ARG(ctx != nullptr && json_out != nullptr && cap > 2);
    ctx->profiling = false;
    PG_CUDA(cudaDeviceSynchronize());
    struct Agg { std::string name; long n; double ms, bytes, flops; };
    std::vector<Agg> agg;
    // PGMVAE_PROF_TIMELINE=<prefix> (file <prefix>.dev<device>.csv): every launch with its stream and its start / end relative to the first one (the
    // overlap of the communication stream with the compute stream is read from this)
    if (const char* path = getenv("PGMVAE_PROF_TIMELINE")) {
        const std::string file = std::string(path) + ".dev" + std::to_string(ctx->device) + ".csv";
        if (FILE* f = ctx->prof.empty() ? nullptr : fopen(file.c_str(), "a")) {
            fprintf(f, "name,stream,start_ms,end_ms\n");
            for (pg_prof_rec& r : ctx->prof) {
                float t0 = 0.f, t1 = 0.f;
                cudaEventElapsedTime(&t0, ctx->prof[0].e0, r.e0);
                cudaEventElapsedTime(&t1, ctx->prof[0].e0, r.e1);
                fprintf(f, "%s,%p,%.4f,%.4f\n", r.name, (void*)r.st, t0, t1);
            }
            fclose(f);
        }
    }
    for (pg_prof_rec& r : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        Agg* a = nullptr;
        for (Agg& x : agg) if (x.name == r.name) { a = &x; break; }
        if (!a) { agg.push_back({r.name, 0, 0, 0, 0}); a = &agg.back(); }
        a->n++; a->ms += ms; a->bytes += r.bytes; a->flops += r.flops;
    }
    ctx->prof.clear();
    std::string out = "[";
    char buf[512];
    for (size_t i = 0; i < agg.size(); ++i) {
        snprintf(buf, sizeof(buf), "%s{\"name\":\"%s\",\"launches\":%ld,\"ms\":%.6f,\"bytes\":%.0f,\"flops\":%.0f}",
                 i ? "," : "", agg[i].name.c_str(), agg[i].n, agg[i].ms, agg[i].bytes, agg[i].flops);
        out += buf;
    }
    out += "]";
    if (out.size() + 1 > cap) {
        pgmvae_set_error("profile_end: buffer too small (%zu needed)", out.size() + 1);
        return PGMVAE_EINVAL;
    }
    memcpy(json_out, out.c_str(), out.size() + 1);
    return PGMVAE_OK;
}

int pgmvae_version(void) { return PGMVAE_VERSION; }
const char* pgmvae_last_error(void) { return g_err; }

int pgmvae_device_count(int* n) {
    PG_CHECK_ARG(n != nullptr);
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        c = 0;
    }
    *n = c;
    return PGMVAE_OK;
}

int pgmvae_device_can_access_peer(int device, int peer, int* out) {
    PG_CHECK_ARG(out != nullptr);
    PG_CUDA(cudaDeviceCanAccessPeer(out, device, peer));
    return PGMVAE_OK;
}

int pgmvae_ctx_create(int device, pgmvae_ctx** out) {
    PG_CHECK_ARG(out != nullptr);
    int n = 0;
    pgmvae_device_count(&n);
    if (n <= 0) {
        pgmvae_set_error("pgmvae_ctx_create: no CUDA device visible; libpgmvae has no CPU fallback");
        return PGMVAE_ENODEV;
    }
    if (device < 0 || device >= n) {
        pgmvae_set_error("pgmvae_ctx_create: device %d out of range (%d visible); "
                         "device -1 (the reference's CPU path) is not supported", device, n);
        return PGMVAE_ENODEV;
    }
    PG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        pgmvae_set_error("pgmvae_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                         device, prop.major, prop.minor);
        return PGMVAE_ENODEV;
    }
    pgmvae_ctx* c = new pgmvae_ctx();
    c->device = device;
    c->sm_count = c->sm_total = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    PG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PG_CUDA(cudaEventCreate(&c->ev0));
    PG_CUDA(cudaEventCreate(&c->ev1));
    *out = c;
    return PGMVAE_OK;
}

int pgmvae_ctx_destroy(pgmvae_ctx* ctx) {
    if (!ctx) return PGMVAE_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scratch_b) cudaFree(ctx->scratch_b);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PGMVAE_OK;
}

int pgmvae_ctx_sync(pgmvae_ctx* ctx) {
    PG_CHECK_ARG(ctx != nullptr);
    PG_CUDA(cudaStreamSynchronize(ctx->stream));
    return PGMVAE_OK;
}

void* pgmvae_ctx_stream(pgmvae_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int pgmvae_ctx_set_precision(pgmvae_ctx* ctx, int prec) {
    PG_CHECK_ARG(ctx != nullptr);
    PG_CHECK_ARG(prec == PGMVAE_PREC_FP32 || prec == PGMVAE_PREC_TF32 || prec == PGMVAE_PREC_BF16);
    ctx->precision = prec;
    return PGMVAE_OK;
}
int pgmvae_ctx_get_precision(pgmvae_ctx* ctx) { return ctx ? ctx->precision : -1; }

int pgmvae_ctx_reserve_sms(pgmvae_ctx* ctx, int n) {
    PG_CHECK_ARG(ctx && n >= 0 && n < ctx->sm_total - 1);
    int left = ctx->sm_total - n;
    left -= left & 1;                        // whole 2-CTA clusters
    ctx->sm_count = left;
    return PGMVAE_OK;
}
int64_t pgmvae_ctx_launch_count(pgmvae_ctx* ctx) { return ctx ? ctx->launches : 0; }

int pgmvae_malloc(pgmvae_ctx* ctx, size_t bytes, void** dptr) {
    PG_CHECK_ARG(ctx != nullptr && dptr != nullptr);
    PG_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        pgmvae_set_error("pgmvae_malloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
        return PGMVAE_ENOMEM;
    }
    return PGMVAE_OK;
}
int pgmvae_free(pgmvae_ctx* ctx, void* dptr) {
    PG_CHECK_ARG(ctx != nullptr);
    if (dptr) PG_CUDA(cudaFree(dptr));
    return PGMVAE_OK;
}
int pgmvae_malloc_host(pgmvae_ctx* ctx, size_t bytes, void** hptr) {
    PG_CHECK_ARG(ctx != nullptr && hptr != nullptr);
    PG_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
    return PGMVAE_OK;
}
int pgmvae_free_host(pgmvae_ctx* ctx, void* hptr) {
    PG_CHECK_ARG(ctx != nullptr);
    if (hptr) PG_CUDA(cudaFreeHost(hptr));
    return PGMVAE_OK;
}
int pgmvae_memcpy_h2d(pgmvae_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream) {
    PG_CHECK_ARG(ctx != nullptr);
    PG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, pg_stream(ctx, stream)));
    return PGMVAE_OK;
}
int pgmvae_memcpy_d2h(pgmvae_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream) {
    PG_CHECK_ARG(ctx != nullptr);
    cudaStream_t s = pg_stream(ctx, stream);
    PG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    PG_CUDA(cudaStreamSynchronize(s));
    return PGMVAE_OK;
}
int pgmvae_memcpy_d2d(pgmvae_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream) {
    PG_CHECK_ARG(ctx != nullptr);
    PG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, pg_stream(ctx, stream)));
    return PGMVAE_OK;
}
int pgmvae_memset(pgmvae_ctx* ctx, void* dst, int byte, size_t bytes, void* stream) {
    PG_CHECK_ARG(ctx != nullptr);
    PG_CUDA(cudaMemsetAsync(dst, byte, bytes, pg_stream(ctx, stream)));
    return PGMVAE_OK;
}
int pgmvae_timer_start(pgmvae_ctx* ctx) {
    PG_CHECK_ARG(ctx != nullptr);
    PG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return PGMVAE_OK;
}
int pgmvae_timer_stop_ms(pgmvae_ctx* ctx, float* ms) {
    PG_CHECK_ARG(ctx != nullptr && ms != nullptr);
    PG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    PG_CUDA(cudaEventSynchronize(ctx->ev1));
    PG_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return PGMVAE_OK;
}

}  // extern "C"
