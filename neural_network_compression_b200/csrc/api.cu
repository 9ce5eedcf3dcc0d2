// api.cu -- the C ABI (include/nnc.h): context, workspace arena, host/device staging, and the entry points
// that string the kernels together.
//
// Boundary being replaced (reference, Python): neural_network_compression/common/utility.py
//   prune_weigth :134-163, get_weight_distribution :334-392, get_quantized_weight :172-240
// and the inline NumPy of common/trainer.py:55-60 (survivor selection) and :195-206 (mask re-apply).
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <exception>
#include <new>

#include "common.cuh"
#include "internal.h"
#include "peer.cuh"

namespace nnc {

static thread_local char g_err[512] = "";

bool debug_sync() {
    static const bool on = getenv("NNC_DEBUG_SYNC") != nullptr;
    return on;
}

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- arena -----------------------------------------------------------------------------------------
// A bump allocator over one device block.  When a call needs more than the block holds, an overflow block is
// taken with cudaMalloc and kept until the next reset, at which point the main block is regrown to the
// high-water mark so that steady-state calls never allocate.
static const size_t ARENA_ALIGN = 256;
static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void arena_reset(nnc_ctx *ctx) {
    if (!ctx->overflow.empty()) {
        NNC_CUDA(cudaStreamSynchronize(ctx->stream));
        for (void *p : ctx->overflow) cudaFree(p);
        ctx->overflow.clear();
        if (ctx->ws) cudaFree(ctx->ws);
        ctx->ws = nullptr;
        ctx->ws_bytes = 0;
    }
    if (ctx->high_water > ctx->ws_bytes) {
        if (ctx->ws) {
            NNC_CUDA(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->ws);
            ctx->ws = nullptr;
            ctx->ws_bytes = 0;
        }
        size_t want = round_up(ctx->high_water + ctx->high_water / 16, 1 << 20);
        void *p = nullptr;
        if (cudaMalloc(&p, want) == cudaSuccess) {
            ctx->ws = static_cast<char *>(p);
            ctx->ws_bytes = want;
        } else {
            cudaGetLastError();  // fall back to per-request blocks
        }
    }
    ctx->ws_off = 0;
    ctx->call_bytes = 0;
}

void arena_reserve(nnc_ctx *ctx, size_t bytes) {
    if (ctx->ws_off == 0 && ctx->overflow.empty() && bytes > ctx->ws_bytes) {
        ctx->high_water = std::max(ctx->high_water, bytes);
        arena_reset(ctx);
    }
}

void *arena_alloc(nnc_ctx *ctx, size_t bytes) {
    bytes = round_up(std::max<size_t>(bytes, 1), ARENA_ALIGN);
    ctx->call_bytes += bytes;
    ctx->high_water = std::max(ctx->high_water, ctx->call_bytes);
    if (ctx->ws && ctx->ws_off + bytes <= ctx->ws_bytes) {
        void *p = ctx->ws + ctx->ws_off;
        ctx->ws_off += bytes;
        return p;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        set_error("workspace: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        throw Error{NNC_ERR_CUDA};
    }
    ctx->overflow.push_back(p);
    return p;
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

void func_dyn_smem(nnc_ctx *ctx, const void *fn, size_t bytes) {
    // (no shortcut for small sizes: the 48 KB default limit covers static + dynamic shared memory together)
    for (auto &e : ctx->func_smem) {
        if (e.first == fn) {
            if (e.second >= bytes) return;
            NNC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            e.second = bytes;
            return;
        }
    }
    NNC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    ctx->func_smem.emplace_back(fn, bytes);
}

Staged stage_in(nnc_ctx *ctx, const void *p, size_t bytes) {
    Staged s;
    s.bytes = bytes;
    if (is_device_ptr(p)) {
        s.dev = const_cast<void *>(p);
        return s;
    }
    s.dev = arena_alloc(ctx, bytes);
    NNC_CUDA(cudaMemcpyAsync(s.dev, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return s;
}
Staged stage_out(nnc_ctx *ctx, void *p, size_t bytes) {
    Staged s;
    s.bytes = bytes;
    if (is_device_ptr(p)) {
        s.dev = p;
        return s;
    }
    s.dev = arena_alloc(ctx, bytes);
    s.host = p;
    return s;
}
Staged stage_inout(nnc_ctx *ctx, void *p, size_t bytes) {
    Staged s = stage_in(ctx, p, bytes);
    if (s.dev != p) s.host = p;
    return s;
}
void stage_finish(nnc_ctx *ctx, const Staged &s) {
    if (s.host) NNC_CUDA(cudaMemcpyAsync(s.host, s.dev, s.bytes, cudaMemcpyDeviceToHost, ctx->stream));
}

// ---- phases ------------------------------------------------------------------------------------------
static void kfold(nnc_ctx *ctx);
static cudaEvent_t prof_event(nnc_ctx *ctx, int i) {
    while ((int)ctx->prof.ev.size() <= i) {
        cudaEvent_t e;
        NNC_CUDA(cudaEventCreate(&e));
        ctx->prof.ev.push_back(e);
    }
    return ctx->prof.ev[i];
}
void prof_begin(nnc_ctx *ctx) {
    ctx->prof.used = 0;
    ctx->prof.names.clear();
    NNC_CUDA(cudaEventRecord(prof_event(ctx, 0), ctx->stream));
    ctx->prof.used = 1;
}
void prof_mark(nnc_ctx *ctx, const char *name) {
    if (ctx->prof.used == 0) return;
    NNC_CUDA(cudaEventRecord(prof_event(ctx, ctx->prof.used), ctx->stream));
    ctx->prof.used++;
    ctx->prof.names.push_back(name);
}
void prof_end(nnc_ctx *ctx) {
    if (ctx->prof.used == 0) return;
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->prof_ms.clear();
    ctx->prof_names.clear();
    for (int i = 1; i < ctx->prof.used; ++i) {
        float ms = 0.f;
        NNC_CUDA(cudaEventElapsedTime(&ms, ctx->prof.ev[i - 1], ctx->prof.ev[i]));
        ctx->prof_ms.push_back(ms);
        if (i > 1) ctx->prof_names += ";";
        ctx->prof_names += ctx->prof.names[i - 1];
    }
    ctx->prof.used = 0;
    ctx->last_launches = ctx->launches;
    if (ctx->ktime) kfold(ctx);
}

static cudaEvent_t kevent(nnc_ctx *ctx) {
    if (ctx->kused == ctx->kev.size()) {
        cudaEvent_t e;
        NNC_CUDA(cudaEventCreate(&e));
        ctx->kev.push_back(e);
    }
    return ctx->kev[ctx->kused++];
}
void klaunch_begin(nnc_ctx *ctx, const char *name) {
    NNC_CUDA(cudaEventRecord(kevent(ctx), ctx->stream));
    ctx->knames.push_back(name);
}
void klaunch_end(nnc_ctx *ctx) { NNC_CUDA(cudaEventRecord(kevent(ctx), ctx->stream)); }
static void kfold(nnc_ctx *ctx) {  // after a stream synchronize: fold this call's launches into the running totals
    for (size_t i = 0; i < ctx->knames.size(); ++i) {
        float t = 0.f;
        NNC_CUDA(cudaEventElapsedTime(&t, ctx->kev[2 * i], ctx->kev[2 * i + 1]));
        std::string nm = ctx->knames[i];
        size_t j = 0;
        for (; j < ctx->kacc_names.size(); ++j)
            if (ctx->kacc_names[j] == nm) break;
        if (j == ctx->kacc_names.size()) {
            ctx->kacc_names.push_back(nm);
            ctx->kacc_ms.push_back(0.0);
            ctx->kacc_cnt.push_back(0);
        }
        ctx->kacc_ms[j] += t;
        ctx->kacc_cnt[j] += 1;
    }
    ctx->knames.clear();
    ctx->kused = 0;
}

// ---- NCCL, bound at run time (dlopen: the library is the one PyTorch already loaded, or the system's) ------------
// Only what is needed: unique id, communicator, int64 all-reduce.  Constants as in nccl.h (2.x ABI).
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUniqueIdBytes, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return nullptr;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.lib, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
        api.lib = nullptr;
        return nullptr;
    }
    return &api;
}

void comm_allreduce(nnc_ctx *ctx, int64_t *d_buf, int count, int op) {
    if (ctx->world <= 1 || count <= 0) return;
    if (ctx->nccl_comm) {
        NcclApi *api = nccl_api();
        const int nccl_op = op == 0 ? 0 /* ncclSum */ : (op == 1 ? 3 /* ncclMin */ : 2 /* ncclMax */);
        const int rc = api->AllReduce(d_buf, d_buf, (size_t)count, 4 /* ncclInt64 */, nccl_op, ctx->nccl_comm, ctx->stream);
        if (rc != 0) NNC_FAIL(NNC_ERR_COMM, "ncclAllReduce: %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
        return;
    }
    if (!ctx->allreduce) NNC_FAIL(NNC_ERR_COMM, "multi-rank context without an all-reduce callback");
    const int rc = ctx->allreduce(ctx->allreduce_user, d_buf, count, op, ctx->stream);
    if (rc != 0) NNC_FAIL(NNC_ERR_COMM, "all-reduce callback failed (%d)", rc);
}

void read_scalars(nnc_ctx *ctx) {
    NNC_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, ctx->stream));
    int comm_error = 0;
    if (ctx->d_comm_error) NNC_CUDA(cudaMemcpyAsync(&comm_error, ctx->d_comm_error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (comm_error) {
        // a rank did not arrive at an in-kernel exchange: the ranks' sequence counters may differ from here on, so the
        // mailboxes are not used again by this context (the NCCL path takes over) until they are reconnected
        ctx->peer_enabled = false;
        cudaMemsetAsync(ctx->d_comm_error, 0, sizeof(int), ctx->stream);
        NNC_FAIL(NNC_ERR_COMM, "a rank did not arrive at an in-kernel peer exchange (time-out); peer exchange disabled on this context");
    }
}

// ---- host-side float32 helpers (every rounding explicit: host code is built with -ffp-contract=off) ----
// np.linspace(start, stop, num) for float32 endpoints (numpy/_core/function_base.py): float32 step,
// y = arange * step + start, last sample forced to `stop`.
static void linspace_f32(float start, float stop, int num, float *out) {
    if (num == 1) {
        out[0] = start;
        return;
    }
    volatile float delta = stop - start;
    const float div = (float)(num - 1);
    volatile float step = delta / div;
    for (int i = 0; i < num; ++i) {
        volatile float y = (float)i;
        if (step == 0.f) {
            y = y / div;
            y = y * delta;
        } else {
            y = y * step;
        }
        volatile float r = y + start;
        out[i] = r;
    }
    out[num - 1] = stop;
}

struct Call {  // RAII: device selection, arena reset, launch accounting, profile bracket
    nnc_ctx *ctx;
    explicit Call(nnc_ctx *c) : ctx(c) {
        if (!c) NNC_FAIL(NNC_ERR_BAD_ARG, "null context");
        NNC_CUDA(cudaSetDevice(c->device));
        arena_reset(c);
        c->launches = 0;
        c->knames.clear();
        c->kused = 0;
        prof_begin(c);
    }
    void finish() {
        prof_end(ctx);
        ctx->last_launches = ctx->launches;
        ctx->total_launches += ctx->launches;
    }
};

// Establishes the shard of this call: one rank owns everything; otherwise the ranks agree on the global element
// count (all-reduce) and every rank must hold exactly the slice nnc_shard_range assigns to it.
static void shard_setup(nnc_ctx *ctx, int64_t n_local) {
    if (ctx->world <= 1) {
        ctx->sh.n_global = n_local;
        ctx->sh.begin = 0;
        ctx->sh.t0 = 0;
        ctx->sh.t1 = n_local > 0 ? np_plan(n_local).num_tiles : 0;
        return;
    }
    int64_t n_global = 0;
    if (ctx->hint_n_global > 0) {
        // the caller told every rank the size of the whole tensor (nnc_ctx_hint_global_n): no all-reduce, no host round trip;
        // a rank whose slice does not fit the shard map of that size fails below like without the hint
        n_global = ctx->hint_n_global;
    } else {
        int64_t *d = arena_alloc_t<int64_t>(ctx, 1);
        NNC_CUDA(cudaMemcpyAsync(d, &n_local, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        comm_allreduce(ctx, d, 1, 0);
        NNC_CUDA(cudaMemcpyAsync(&n_global, d, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    int64_t b = 0, e = 0;
    uint32_t t0 = 0, t1 = 0;
    if (n_global <= 0) NNC_FAIL(NNC_ERR_BAD_ARG, "sharded call on an empty tensor");
    np_shard_range(n_global, ctx->rank, ctx->world, &b, &e, &t0, &t1);
    if (e - b != n_local)
        NNC_FAIL(NNC_ERR_BAD_ARG, "rank %d of %d holds %lld elements of a %lld-element tensor; nnc_shard_range assigns [%lld, %lld)",
                 ctx->rank, ctx->world, (long long)n_local, (long long)n_global, (long long)b, (long long)e);
    ctx->sh.n_global = n_global;
    ctx->sh.begin = b;
    ctx->sh.t0 = t0;
    ctx->sh.t1 = t1;
}

}  // namespace nnc

using namespace nnc;

#define NNC_TRY try {
#define NNC_CATCH                                                   \
    }                                                               \
    catch (const nnc::Error &e) {                                   \
        cudaGetLastError();                                         \
        return e.code;                                              \
    }                                                               \
    catch (const std::bad_alloc &) {                                \
        nnc::set_error("host allocation failed");                   \
        return NNC_ERR_INTERNAL;                                    \
    }                                                               \
    catch (const std::exception &e) {                               \
        nnc::set_error("unexpected exception: %s", e.what());       \
        return NNC_ERR_INTERNAL;                                    \
    }                                                               \
    return NNC_OK;

extern "C" {

int nnc_version(void) { return NNC_VERSION; }
const char *nnc_last_error(void) { return g_err; }

int nnc_ctx_create(int device, nnc_ctx **out) {
    NNC_TRY
    if (!out) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_ctx_create: null out");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        NNC_FAIL(NNC_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= count) NNC_FAIL(NNC_ERR_BAD_ARG, "device %d outside [0, %d)", device, count);
    NNC_CUDA(cudaSetDevice(device));
    nnc_ctx *c = new nnc_ctx();
    c->device = device;
    cudaDeviceProp prop;
    NNC_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    NNC_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    NNC_CUDA(cudaMalloc(&c->d_scal, sizeof(DevScalars)));
    NNC_CUDA(cudaMallocHost(&c->h_scal, sizeof(DevScalars)));
    NNC_CUDA(cudaEventCreate(&c->ev0));
    NNC_CUDA(cudaEventCreate(&c->ev1));
    *out = c;
    NNC_CATCH
}

void nnc_ctx_destroy(nnc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->nccl_comm) {
        NcclApi *api = nccl_api();
        if (api) api->CommDestroy(ctx->nccl_comm);
    }
    for (int r = 0; r < 16; ++r)
        if (ctx->peer_mail[r] && ctx->peer_mail[r] != ctx->peer_local) cudaIpcCloseMemHandle(ctx->peer_mail[r]);
    if (ctx->peer_local) cudaFree(ctx->peer_local);
    for (void *p : ctx->overflow) cudaFree(p);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->desc_ptr) cudaFree(ctx->desc_ptr);
    if (ctx->d_comm_error) cudaFree(ctx->d_comm_error);
    if (ctx->d_scal) cudaFree(ctx->d_scal);
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    for (cudaEvent_t e : ctx->prof.ev) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->kev) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int nnc_ctx_set_stream(nnc_ctx *ctx, void *cuda_stream) {
    NNC_TRY
    if (!ctx) NNC_FAIL(NNC_ERR_BAD_ARG, "null context");
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    ctx->user_stream = cuda_stream != nullptr;
    NNC_CATCH
}

int nnc_ctx_reserve(nnc_ctx *ctx, size_t bytes) {
    NNC_TRY
    if (!ctx) NNC_FAIL(NNC_ERR_BAD_ARG, "null context");
    NNC_CUDA(cudaSetDevice(ctx->device));
    arena_reset(ctx);
    arena_reserve(ctx, bytes);
    NNC_CATCH
}

int nnc_timer_start(nnc_ctx *ctx) {
    NNC_TRY
    if (!ctx) NNC_FAIL(NNC_ERR_BAD_ARG, "null context");
    NNC_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    NNC_CATCH
}
int nnc_timer_stop(nnc_ctx *ctx, float *ms_out) {
    NNC_TRY
    if (!ctx || !ms_out) NNC_FAIL(NNC_ERR_BAD_ARG, "null argument");
    NNC_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    NNC_CUDA(cudaEventSynchronize(ctx->ev1));
    NNC_CUDA(cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
    NNC_CATCH
}

int nnc_last_profile(nnc_ctx *ctx, float *ms_out, int cap, int *n_out, const char **names_out, int64_t *launches_out) {
    NNC_TRY
    if (!ctx) NNC_FAIL(NNC_ERR_BAD_ARG, "null context");
    int n = (int)ctx->prof_ms.size();
    if (ms_out)
        for (int i = 0; i < n && i < cap; ++i) ms_out[i] = ctx->prof_ms[i];
    if (n_out) *n_out = n;
    if (names_out) *names_out = ctx->prof_names.c_str();
    if (launches_out) *launches_out = ctx->last_launches;
    NNC_CATCH
}

int nnc_ctx_set_kernel_timing(nnc_ctx *ctx, int on, const char *name_filter) {
    NNC_TRY
    if (!ctx) NNC_FAIL(NNC_ERR_BAD_ARG, "null context");
    ctx->ktime = on != 0;
    ctx->kfilter = name_filter ? name_filter : "";
    ctx->kacc_names.clear();
    ctx->kacc_ms.clear();
    ctx->kacc_cnt.clear();
    NNC_CATCH
}
int nnc_ctx_total_launches(nnc_ctx *ctx, int64_t *out) {
    NNC_TRY
    if (!ctx || !out) NNC_FAIL(NNC_ERR_BAD_ARG, "null argument");
    *out = ctx->total_launches;
    NNC_CATCH
}
int nnc_last_kernel_times(nnc_ctx *ctx, const char **out) {
    NNC_TRY
    if (!ctx || !out) NNC_FAIL(NNC_ERR_BAD_ARG, "null argument");
    ctx->ktimes.clear();
    char line[256];
    for (size_t j = 0; j < ctx->kacc_names.size(); ++j) {
        snprintf(line, sizeof(line), "%s%s:%lld:%.6f", j ? ";" : "", ctx->kacc_names[j].c_str(), ctx->kacc_cnt[j], ctx->kacc_ms[j]);
        ctx->ktimes += line;
    }
    *out = ctx->ktimes.c_str();
    NNC_CATCH
}

int nnc_comm_unique_id(char *out128) {
    NNC_TRY
    if (!out128) NNC_FAIL(NNC_ERR_BAD_ARG, "null argument");
    NcclApi *api = nccl_api();
    if (!api) NNC_FAIL(NNC_ERR_COMM, "libnccl.so.2 could not be loaded");
    const int rc = api->GetUniqueId(out128);
    if (rc != 0) NNC_FAIL(NNC_ERR_COMM, "ncclGetUniqueId failed (%d)", rc);
    NNC_CATCH
}

int nnc_ctx_init_nccl(nnc_ctx *ctx, const char *id128, int rank, int world) {
    NNC_TRY
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_ctx_init_nccl: bad argument");
    NcclApi *api = nccl_api();
    if (!api) NNC_FAIL(NNC_ERR_COMM, "libnccl.so.2 could not be loaded");
    NNC_CUDA(cudaSetDevice(ctx->device));
    NcclUniqueIdBytes id;
    memcpy(id.internal, id128, sizeof(id.internal));
    void *comm = nullptr;
    const int rc = api->CommInitRank(&comm, world, id, rank);
    if (rc != 0) NNC_FAIL(NNC_ERR_COMM, "ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
    if (ctx->nccl_comm) api->CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    NNC_CATCH
}

int nnc_peer_mailbox_create(nnc_ctx *ctx, int world, char *handle_out64) {
    NNC_TRY
    if (!ctx || !handle_out64 || world < 2 || world > PEER_MAX_WORLD) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_peer_mailbox_create: world = %d", world);
    NNC_CUDA(cudaSetDevice(ctx->device));
    if (ctx->peer_local) NNC_FAIL(NNC_ERR_BAD_ARG, "the context already has a mailbox");
    const size_t bytes = peer_mailbox_bytes(world);
    NNC_CUDA(cudaMalloc(&ctx->peer_local, bytes));
    NNC_CUDA(cudaMemset(ctx->peer_local, 0, bytes));
    NNC_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t hdl;
    NNC_CUDA(cudaIpcGetMemHandle(&hdl, ctx->peer_local));
    static_assert(sizeof(hdl) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out64, &hdl, sizeof(hdl));
    NNC_CATCH
}

int nnc_peer_mailbox_connect(nnc_ctx *ctx, const char *handles, int rank, int world) {
    NNC_TRY
    if (ctx && !handles) {  // disconnect: back to NCCL exchanges
        ctx->peer_enabled = false;
        return NNC_OK;
    }
    if (!ctx || !ctx->peer_local || world < 2 || world > PEER_MAX_WORLD || rank < 0 || rank >= world)
        NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_peer_mailbox_connect: bad argument");
    NNC_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            ctx->peer_mail[r] = ctx->peer_local;
            continue;
        }
        cudaIpcMemHandle_t hdl;
        memcpy(&hdl, handles + (size_t)r * sizeof(hdl), sizeof(hdl));
        void *p = nullptr;
        NNC_CUDA(cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_mail[r] = p;
    }
    ctx->peer_enabled = true;
    ctx->peer_seq = 0;
    NNC_CATCH
}

int nnc_ctx_hint_global_n(nnc_ctx *ctx, int64_t n_global) {
    NNC_TRY
    if (!ctx || n_global < 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_ctx_hint_global_n: bad argument");
    ctx->hint_n_global = n_global;
    NNC_CATCH
}

int nnc_shard_range(int64_t n, int rank, int world, int64_t *begin, int64_t *end) {
    NNC_TRY
    if (n <= 0 || world < 1 || rank < 0 || rank >= world || !begin || !end)
        NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_shard_range: n = %lld, rank %d / world %d", (long long)n, rank, world);
    np_shard_range(n, rank, world, begin, end, nullptr, nullptr);
    NNC_CATCH
}

int nnc_ctx_set_comm(nnc_ctx *ctx, int rank, int world, nnc_allreduce_i64_fn fn, void *user) {
    NNC_TRY
    if (!ctx || world < 1 || rank < 0 || rank >= world || (world > 1 && !fn))
        NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_ctx_set_comm: rank %d / world %d", rank, world);
    ctx->rank = rank;
    ctx->world = world;
    ctx->allreduce = fn;
    ctx->allreduce_user = user;
    NNC_CATCH
}

// ---- pruning -------------------------------------------------------------------------------------------
int nnc_stats_f32(nnc_ctx *ctx, const float *w, int64_t n, float *mean_out, float *var_out, float *std_out) {
    NNC_TRY
    Call call(ctx);
    if (!w || n <= 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_stats_f32: w = %p, n = %lld", (const void *)w, (long long)n);
    shard_setup(ctx, n);
    Staged sw = stage_in(ctx, w, sizeof(float) * (size_t)n);
    np_stats(ctx, static_cast<const float *>(sw.dev), n);
    prof_mark(ctx, "stats");
    read_scalars(ctx);
    if (mean_out) *mean_out = ctx->h_scal->mean;
    if (var_out) *var_out = ctx->h_scal->var;
    if (std_out) *std_out = ctx->h_scal->std_;
    call.finish();
    NNC_CATCH
}

int nnc_prune_f32(nnc_ctx *ctx, float *w, int64_t n, double threshold, int std_smooth, int threshold_mode, uint8_t *mask,
                  double *thr_out, int64_t *n_pruned_out) {
    NNC_TRY
    Call call(ctx);
    if (!w || !mask || n < 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_prune_f32: null buffer or n < 0");
    if (n == 0) {  // np.std([]) is NaN (with a warning); nothing to mask
        if (thr_out) *thr_out = std_smooth ? NAN : threshold;
        if (n_pruned_out) *n_pruned_out = 0;
        call.finish();
        return NNC_OK;
    }
    shard_setup(ctx, n);
    Staged sw = stage_inout(ctx, w, sizeof(float) * (size_t)n);
    Staged sm = stage_out(ctx, mask, (size_t)n);
    prof_mark(ctx, "h2d");
    prune_device(ctx, static_cast<float *>(sw.dev), n, threshold, std_smooth, threshold_mode, static_cast<uint8_t *>(sm.dev));
    stage_finish(ctx, sw);
    stage_finish(ctx, sm);
    read_scalars(ctx);
    prof_mark(ctx, "d2h");
    if (thr_out) *thr_out = ctx->h_scal->thr;
    if (n_pruned_out) *n_pruned_out = (int64_t)ctx->h_scal->n_pruned;
    call.finish();
    NNC_CATCH
}

int nnc_mask_apply_f32(nnc_ctx *ctx, float *w, const uint8_t *mask, int64_t n) {
    NNC_TRY
    Call call(ctx);
    if (n == 0) {
        call.finish();
        return NNC_OK;
    }
    if (!w || !mask || n < 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_mask_apply_f32: null buffer or n < 0");
    Staged sw = stage_inout(ctx, w, sizeof(float) * (size_t)n);
    Staged sm = stage_in(ctx, mask, (size_t)n);
    mask_apply_device(ctx, static_cast<float *>(sw.dev), static_cast<const uint8_t *>(sm.dev), n);
    prof_mark(ctx, "mask_apply");
    stage_finish(ctx, sw);
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    call.finish();
    NNC_CATCH
}

// ---- weight distribution -----------------------------------------------------------------------------------
int nnc_compact_nonzero_f32(nnc_ctx *ctx, const float *w, int64_t n, float *out, int64_t *n_nz_out) {
    NNC_TRY
    Call call(ctx);
    if (n < 0 || !n_nz_out || (n > 0 && (!w || !out))) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_compact_nonzero_f32: bad argument");
    *n_nz_out = 0;
    if (n > 0) {
        Staged sw = stage_in(ctx, w, sizeof(float) * (size_t)n);
        Staged so = stage_out(ctx, out, sizeof(float) * (size_t)n);
        int64_t c = compact_ordered_device(ctx, static_cast<const float *>(sw.dev), n, static_cast<float *>(so.dev));
        prof_mark(ctx, "compact");
        so.bytes = sizeof(float) * (size_t)c;
        if (c > 0) stage_finish(ctx, so);
        NNC_CUDA(cudaStreamSynchronize(ctx->stream));
        *n_nz_out = c;
    }
    call.finish();
    NNC_CATCH
}

int nnc_minmax_f32(nnc_ctx *ctx, const float *w, int64_t n, int skip_zeros, float *min_out, float *max_out,
                   int64_t *count_out) {
    NNC_TRY
    Call call(ctx);
    if (!w || n <= 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_minmax_f32: empty input");
    Staged sw = stage_in(ctx, w, sizeof(float) * (size_t)n);
    float mn, mx;
    int64_t cnt;
    minmax_device(ctx, static_cast<const float *>(sw.dev), n, skip_zeros, &mn, &mx, &cnt);
    prof_mark(ctx, "minmax");
    if (min_out) *min_out = mn;
    if (max_out) *max_out = mx;
    if (count_out) *count_out = cnt;
    call.finish();
    NNC_CATCH
}

int nnc_hist_edges_f32(nnc_ctx *ctx, const float *w, int64_t n, const float *edges, int n_edges, int skip_zeros,
                       int64_t *counts_out) {
    NNC_TRY
    Call call(ctx);
    if (!w || n <= 0 || !edges || !counts_out || n_edges < 2 || n_edges > 1025)
        NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_hist_edges_f32: bad argument (n_edges = %d)", n_edges);
    Staged sw = stage_in(ctx, w, sizeof(float) * (size_t)n);
    hist_edges_device(ctx, static_cast<const float *>(sw.dev), n, edges, n_edges, skip_zeros, counts_out);
    prof_mark(ctx, "hist");
    call.finish();
    NNC_CATCH
}

// get_weight_distribution (utility.py:334-392): the two sweeps (min/max, 31-bin histogram) run on the device;
// the 31-value CDF and its 300-point linear interpolation (scipy interp1d, float32 abscissae, float64
// ordinates) are a few hundred scalar operations done here on the host in the reference's own precision.
int nnc_weight_cdf_f32(nnc_ctx *ctx, const float *w, int64_t n, int skip_zeros, float *xnew300, double *cdf300) {
    NNC_TRY
    Call call(ctx);
    if (!w || n <= 0 || !xnew300 || !cdf300) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_weight_cdf_f32: bad argument");
    Staged sw = stage_in(ctx, w, sizeof(float) * (size_t)n);
    const float *d_w = static_cast<const float *>(sw.dev);
    float mn, mx;
    int64_t cnt;
    minmax_device(ctx, d_w, n, skip_zeros, &mn, &mx, &cnt);
    prof_mark(ctx, "minmax");
    if (cnt == 0) NNC_FAIL(NNC_ERR_NOT_ENOUGH, "weight distribution of an empty selection");
    float steps[32];
    linspace_f32(mn, mx, 32, steps);  // utility.py:365
    int64_t counts[31];
    hist_edges_device(ctx, d_w, n, steps, 32, skip_zeros, counts);  // utility.py:366-372
    prof_mark(ctx, "hist");
    int64_t tot = 0;
    for (int b = 0; b < 31; ++b) tot += counts[b];
    double cdf[31];
    for (int b = 0; b < 31; ++b) {  // :375-385 (tot == 0 gives NaNs exactly like 0/0 in NumPy)
        volatile double p = (double)counts[b] / (double)tot;
        cdf[b] = b == 0 ? p : p + cdf[b - 1];
    }
    const double last = cdf[30];
    for (int b = 0; b < 31; ++b) cdf[b] = cdf[b] / last;
    linspace_f32(steps[0], steps[30], 300, xnew300);  // :387
    for (int t = 0; t < 300; ++t) {                   // :388-390 interp1d(x, cdf)(xnew), kind="linear"
        int idx = 0;
        while (idx < 31 && steps[idx] < xnew300[t]) idx++;  // searchsorted, side="left"
        idx = std::min(30, std::max(1, idx));
        const float x_lo = steps[idx - 1], x_hi = steps[idx];
        volatile float num_hi = xnew300[t] - x_lo, num_lo = x_hi - xnew300[t], den = x_hi - x_lo;
        volatile float r_hi = num_hi / den, r_lo = num_lo / den;
        volatile double a = (double)r_hi * cdf[idx], b = (double)r_lo * cdf[idx - 1];
        cdf300[t] = a + b;
    }
    call.finish();
    NNC_CATCH
}

int nnc_gather_f32(nnc_ctx *ctx, const float *w, int64_t n, const int64_t *idx, int m, float *out) {
    NNC_TRY
    Call call(ctx);
    if (!w || n <= 0 || !idx || !out || m < 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_gather_f32: bad argument");
    if (m > 0) {
        if (is_device_ptr(w)) {
            gather_device(ctx, w, n, idx, m, out);
        } else {  // host tensor: no reason to ship it to the device for m loads
            for (int i = 0; i < m; ++i) {
                if (idx[i] < 0 || idx[i] >= n) NNC_FAIL(NNC_ERR_BAD_ARG, "gather: index %lld out of range", (long long)idx[i]);
                out[i] = w[idx[i]];
            }
        }
    }
    call.finish();
    NNC_CATCH
}

// ---- k-means ---------------------------------------------------------------------------------------------
// The k-means pipeline on a device-resident tensor (prologue -> compaction -> sort -> Lloyd -> emission); shared by
// nnc_kmeans1d_f32 and nnc_compress_f32.  Output pointers are the caller's (host or device); w is a device pointer.
// nz_bound: an upper bound of the non-zero count of the shard when the caller knows one (-1: none), sizes the buffers.
// prefilled: the prologue already ran fused with the pruning pass (prune_device with a QuantFuse): the survivors are in
// prefilled (capacity n) and DevScalars holds the prologue's results.
static void kmeans_on_device(nnc_ctx *ctx, const float *d_w, int64_t n, int64_t nz_bound, const float *init, int k, int max_iter,
                             double tol, int flags, float *centers, float *centred, int32_t *labels, float *ris, uint8_t *packed, int bits,
                             int64_t *hist, nnc_kmeans_info *info, float *prefilled = nullptr) {
    // n: elements of this rank's shard; n_global: of the whole tensor (they coincide on one rank)
    const bool init_linear = (flags & NNC_KM_INIT_LINEAR) != 0;
    const int64_t n_global = ctx->sh.n_global;
    if ((!init && !init_linear) || n <= 0 || k <= 0 || max_iter < 1 || tol < 0)
        NNC_FAIL(NNC_ERR_BAD_ARG, "k-means: bad argument (n = %lld, k = %d, max_iter = %d)", (long long)n, k, max_iter);
    if (k > NNC_KMAX) NNC_FAIL(NNC_ERR_UNSUPPORTED, "k = %d exceeds NNC_KMAX = %d", k, NNC_KMAX);
    if ((int64_t)k > n_global) NNC_FAIL(NNC_ERR_NOT_ENOUGH, "n_samples=%lld should be >= n_clusters=%d.", (long long)n_global, k);
    if (!init_linear)
        for (int j = 0; j < k; ++j)
            if (!isfinite(init[j])) NNC_FAIL(NNC_ERR_NONFINITE, "initial centroid %d is not finite", j);
    // 1. mean (NumPy pairwise), min/max, key range, survivor count -- and the survivors, compacted in the same read
    const int64_t cap = std::max<int64_t>(1, std::min<int64_t>(n, nz_bound >= 0 ? nz_bound : n));
    float *buf_a = prefilled ? prefilled : arena_alloc_t<float>(ctx, (size_t)cap);
    float *buf_b = arena_alloc_t<float>(ctx, (size_t)cap);
    if (!prefilled) {
        quant_prologue(ctx, d_w, n, buf_a, cap);
        read_scalars(ctx);
        prof_mark(ctx, "prologue");
    }
    const DevScalars sc = *ctx->h_scal;
    if (sc.n_nonfinite) NNC_FAIL(NNC_ERR_NONFINITE, "Input X contains NaN or infinity.");
    const int64_t n_nz = (int64_t)(ctx->world > 1 ? sc.n_nz_local : sc.n_nz);  // of this shard
    if (n_nz > (prefilled ? n : cap))
        NNC_FAIL(NNC_ERR_INTERNAL, "k-means: %lld survivors exceed the reserved %lld", (long long)n_nz, (long long)cap);
    // 2. survivors -> sorted
    // (a large narrow-range layer: as (value, multiplicity) runs from a key histogram, khist.cu; otherwise radix sort)
    const float *d_sorted = buf_a;
    SortedRuns runs;
    if (n_nz > 0) {
        if (hist_sort_applicable(n_nz, sc.amin_nz_m1 + 1u, sc.amax_bits)) {
            runs = hist_sort_f32(ctx, buf_a, buf_b, n_nz, sc.amin_nz_m1 + 1u, sc.amax_bits);
            d_sorted = runs.val;
        } else {
            d_sorted = radix_sort_f32(ctx, buf_a, buf_b, n_nz, sc.amin_nz_m1 + 1u, sc.amax_bits);
        }
        prof_mark(ctx, "sort");
    }
    // 3. Lloyd iterations on the sorted survivors + the zero run
    LloydHandle h;
    h.d_sorted = d_sorted;
    h.d_cnt = runs.cnt;
    h.n_ent = runs.n_ent;
    h.n_nz = n_nz;
    h.n0 = n - n_nz;
    h.n = n_global;
    h.k = k;
    h.mean = sc.mean;
    std::vector<float> c_final(k), c_emit(k), lin;
    if (init_linear) {  // utility.py:206-209: np.linspace(min, max, num=k) on float32 scalars
        lin.resize(k);
        linspace_f32(ord2f(sc.min_ord), ord2f(sc.max_ord), k, lin.data());
        init = lin.data();
    }
    LloydResult lr = lloyd_run(ctx, h, init, max_iter, tol, c_final.data(), c_emit.data(), hist);
    // 4. final E-step in original order
    volatile float xlo = ord2f(sc.min_ord) - sc.mean, xhi = ord2f(sc.max_ord) - sc.mean;
    const float xabs = fmaxf(fabsf(xlo), fabsf(xhi));
    double inertia = NAN;
    const bool want_emit = labels || ris || packed || (info && (flags & NNC_KM_INERTIA));
    if (want_emit) {
        Staged sl, sr, sp;
        if (labels) sl = stage_out(ctx, labels, sizeof(int32_t) * (size_t)n);
        if (ris) sr = stage_out(ctx, ris, sizeof(float) * (size_t)n);
        if (packed) sp = stage_out(ctx, packed, (size_t)((n * (int64_t)bits + 7) / 8));
        emit_device(ctx, d_w, n, c_emit.data(), c_final.data(), k, sc.mean, xabs, xlo, xhi, nullptr, static_cast<int32_t *>(sl.dev),
                    static_cast<float *>(sr.dev), static_cast<uint8_t *>(sp.dev), bits, nullptr,
                    (info && (flags & NNC_KM_INERTIA)) ? &inertia : nullptr);
        prof_mark(ctx, "emit");
        stage_finish(ctx, sl);
        stage_finish(ctx, sr);
        stage_finish(ctx, sp);
        NNC_CUDA(cudaStreamSynchronize(ctx->stream));
        prof_mark(ctx, "d2h");
    }
    for (int j = 0; j < k; ++j) {
        if (centred) centred[j] = c_final[j];
        if (centers) {
            volatile float v = c_final[j] + sc.mean;  // best_centers += X_mean (_kmeans.py:1546), float32
            centers[j] = v;
        }
    }
    if (info) {
        info->n_iter = lr.n_iter;
        info->strict = lr.strict;
        info->n_relocations = lr.n_reloc;
        info->fixed_exp = lr.fixed_exp;
        info->mean = sc.mean;
        info->tol = lr.tol;
        info->inertia = inertia;
        info->n_nonzero = (int64_t)sc.n_nz;
    }
}

int nnc_kmeans1d_f32(nnc_ctx *ctx, const float *w, int64_t n, const float *init, int k, int max_iter, double tol,
                     int flags, float *centers, float *centred, int32_t *labels, float *ris, uint8_t *packed, int bits,
                     int64_t *hist, nnc_kmeans_info *info) {
    NNC_TRY
    Call call(ctx);
    if (!w || n <= 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_kmeans1d_f32: empty input");
    shard_setup(ctx, n);
    Staged sw = stage_in(ctx, w, sizeof(float) * (size_t)n);
    prof_mark(ctx, "h2d");
    kmeans_on_device(ctx, static_cast<const float *>(sw.dev), n, -1, init, k, max_iter, tol, flags, centers, centred, labels, ris,
                     packed, bits, hist, info);
    call.finish();
    NNC_CATCH
}

// Prune + quantize in one call (Trainer._prune_parameters then Trainer.quantize on one tensor, trainer.py:177-193,
// :42-72): a host tensor crosses the bus once.
int nnc_compress_f32(nnc_ctx *ctx, float *w, int64_t n, double threshold, int std_smooth, int threshold_mode, int write_back,
                     uint8_t *mask, double *thr_out, int64_t *n_pruned_out, const float *init, int k, int max_iter, double tol,
                     int flags, float *centers, float *centred, uint8_t *packed, int bits, int64_t *hist, nnc_kmeans_info *info) {
    NNC_TRY
    Call call(ctx);
    if (!w || !mask || n <= 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_compress_f32: null buffer or empty input");
    shard_setup(ctx, n);
    Staged sw = stage_inout(ctx, w, sizeof(float) * (size_t)n);
    // NNC_KM_MASK_BITS: `mask` receives ceil(n / 8) bytes (bit i of byte i / 8), the byte mask stays in the workspace
    const bool mask_bits = (flags & NNC_KM_MASK_BITS) != 0;
    Staged sm;
    uint8_t *d_mask_bytes = nullptr;
    if (mask_bits) {
        d_mask_bytes = arena_alloc_t<uint8_t>(ctx, (size_t)n);
        sm = stage_out(ctx, mask, (size_t)((n + 7) / 8));
    } else {
        sm = stage_out(ctx, mask, (size_t)n);
        d_mask_bytes = static_cast<uint8_t *>(sm.dev);
    }
    prof_mark(ctx, "h2d");
    float *d_w = static_cast<float *>(sw.dev);
    // with the std-scaled threshold the k-means prologue of the pruned tensor rides on the apply pass (one sweep less)
    QuantFuse fuse;
    if (std_smooth && !getenv("NNC_NO_FUSE")) {
        fuse.out = arena_alloc_t<float>(ctx, (size_t)n);
        fuse.capacity = n;
    }
    prune_device(ctx, d_w, n, threshold, std_smooth, threshold_mode, d_mask_bytes, fuse.out ? &fuse : nullptr);
    if (write_back) stage_finish(ctx, sw);  // a device-resident tensor was pruned in place already
    if (mask_bits) pack_bits_device(ctx, d_mask_bytes, n, static_cast<uint8_t *>(sm.dev));
    stage_finish(ctx, sm);
    if (sw.host || sm.host) NNC_CUDA(cudaStreamSynchronize(ctx->stream));  // host copies complete (the scalars were read by prune_device)
    prof_mark(ctx, "prune_d2h");
    if (thr_out) *thr_out = ctx->h_scal->thr;
    if (n_pruned_out) *n_pruned_out = (int64_t)ctx->h_scal->n_pruned;
    // every pruned element is a zero now: on one rank n - n_pruned bounds the survivors (on several, n_pruned is global)
    const int64_t nz_bound = ctx->world == 1 ? n - (int64_t)ctx->h_scal->n_pruned : -1;
    kmeans_on_device(ctx, d_w, n, nz_bound, init, k, max_iter, tol, flags, centers, centred, nullptr, nullptr, packed, bits, hist, info,
                     fuse.done ? fuse.out : nullptr);
    call.finish();
    NNC_CATCH
}

int nnc_assign_f32(nnc_ctx *ctx, const float *w, int64_t n, const float *centred, int k, float mean, const float *values,
                   int32_t *labels, float *ris, uint8_t *packed, int bits, int64_t *hist, double *inertia_out) {
    NNC_TRY
    Call call(ctx);
    if (!w || !centred || n <= 0 || k <= 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_assign_f32: bad argument");
    if (k > NNC_KMAX) NNC_FAIL(NNC_ERR_UNSUPPORTED, "k = %d exceeds NNC_KMAX = %d", k, NNC_KMAX);
    Staged sw = stage_in(ctx, w, sizeof(float) * (size_t)n);
    Staged sl, sr, sp;
    if (labels) sl = stage_out(ctx, labels, sizeof(int32_t) * (size_t)n);
    if (ris) sr = stage_out(ctx, ris, sizeof(float) * (size_t)n);
    if (packed) sp = stage_out(ctx, packed, (size_t)((n * (int64_t)bits + 7) / 8));
    emit_device(ctx, static_cast<const float *>(sw.dev), n, centred, centred, k, mean, -1.f, 0.f, 0.f, values,
                static_cast<int32_t *>(sl.dev), static_cast<float *>(sr.dev), static_cast<uint8_t *>(sp.dev), bits, hist,
                inertia_out);
    prof_mark(ctx, "emit");
    stage_finish(ctx, sl);
    stage_finish(ctx, sr);
    stage_finish(ctx, sp);
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    call.finish();
    NNC_CATCH
}

int nnc_unpack_gather_f32(nnc_ctx *ctx, const uint8_t *packed, int64_t n, int bits, const float *values, int k, float *out) {
    NNC_TRY
    Call call(ctx);
    if (!packed || !values || !out || n <= 0 || k <= 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_unpack_gather_f32: bad argument");
    if (bits < 1 || bits > 16) NNC_FAIL(NNC_ERR_BAD_ARG, "unpack: bits = %d outside [1, 16]", bits);
    Staged sp = stage_in(ctx, packed, (size_t)((n * (int64_t)bits + 7) / 8));
    Staged so = stage_out(ctx, out, sizeof(float) * (size_t)n);
    unpack_gather_device(ctx, static_cast<const uint8_t *>(sp.dev), n, bits, values, k, static_cast<float *>(so.dev));
    prof_mark(ctx, "unpack");
    stage_finish(ctx, so);
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    call.finish();
    NNC_CATCH
}

int nnc_pack_bits_u8(nnc_ctx *ctx, const uint8_t *src, int64_t n, uint8_t *dst_bits) {
    NNC_TRY
    Call call(ctx);
    if (!src || !dst_bits || n <= 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_pack_bits_u8: bad argument");
    Staged ss = stage_in(ctx, src, (size_t)n);
    Staged sd = stage_out(ctx, dst_bits, (size_t)((n + 7) / 8));
    pack_bits_device(ctx, static_cast<const uint8_t *>(ss.dev), n, static_cast<uint8_t *>(sd.dev));
    prof_mark(ctx, "pack_bits");
    stage_finish(ctx, sd);
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    call.finish();
    NNC_CATCH
}

int nnc_grad_segsum_f32(nnc_ctx *ctx, const float *grad, const void *codes, int64_t n, int bits, int k, double *out) {
    NNC_TRY
    Call call(ctx);
    if (!grad || !codes || !out || n <= 0 || k <= 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_grad_segsum_f32: bad argument");
    if (bits < 0 || bits > 16) NNC_FAIL(NNC_ERR_BAD_ARG, "segsum: bits = %d outside [0, 16]", bits);
    Staged sg = stage_in(ctx, grad, sizeof(float) * (size_t)n);
    const size_t code_bytes = bits == 0 ? sizeof(int32_t) * (size_t)n : (size_t)((n * (int64_t)bits + 7) / 8);
    Staged sc = stage_in(ctx, codes, code_bytes);
    grad_segsum_device(ctx, static_cast<const float *>(sg.dev), sc.dev, n, bits, k, out);
    prof_mark(ctx, "segsum");
    call.finish();
    NNC_CATCH
}

}  // extern "C"
