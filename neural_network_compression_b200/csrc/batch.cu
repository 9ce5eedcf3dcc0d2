// batch.cu -- nnc_compress_many_f32: all tensors of a model in one call, on a pool of native worker threads.
//
// Reference context: the trainer prunes and quantises a model layer by layer, tensor by tensor (common/trainer.py:50-70,
// 177-193; models/le_net_5.py:17-34).  Every LeNet tensor is small (10 .. 627 200 weights): its ~25 kernel launches and
// handful of host round trips, not its bytes, are its cost.  The pool overlaps them across tensors: one persistent host
// thread per worker, each with its own context (stream, workspace arena, scalar mirrors), jobs handed out largest first.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "internal.h"

namespace nnc {

namespace {

struct BatchParams {
    int std_smooth, threshold_mode, bits, mode;
};

// utility.py:210-223: for each of the k + 1 targets t of np.linspace(0, 1, k + 1) the first cdf value closest to t (Python
// min over abs), then the x of the first cdf entry equal to that value (np.argmax of the equality)
void density_init(int k_lin, const float *xnew, const double *cdf, float *space) {
    const int num = k_lin + 1, div = num - 1;
    const double step = 1.0 / (double)div;
    for (int i = 0; i < num; ++i) {
        double t = (double)i * step + 0.0;  // np.linspace: arange * step + start ...
        if (i == num - 1) t = 1.0;          // ... and the last point set to stop
        int j = 0;
        double best = fabs(cdf[0] - t);
        for (int q = 1; q < 300; ++q) {
            const double d = fabs(cdf[q] - t);
            if (d < best) {
                best = d;
                j = q;
            }
        }
        int idx = 0;
        while (cdf[idx] != cdf[j]) ++idx;
        space[i] = xnew[idx];
    }
}

int index_bits(int k) {
    int b = 0;
    while ((1 << b) < k) ++b;
    return b < 1 ? 1 : b;
}

void run_job(nnc_ctx *c, nnc_tensor_job &j, const BatchParams &p) {
    j.status = NNC_OK;
    j.k = 0;
    j.code_bits = 0;
    j.error[0] = 0;
    j.thr = 0.0;
    j.n_pruned = 0;
    memset(&j.info, 0, sizeof(j.info));
    const int k_lin = 1 << p.bits;
    int rc = NNC_OK;
    if (j.n < (int64_t)k_lin + 1) {  // "not enough bits" (utility.py:202-204): pruned, not quantised
        if (j.prune) rc = nnc_prune_f32(c, j.w, j.n, j.threshold, p.std_smooth, p.threshold_mode, j.mask, &j.thr, &j.n_pruned);
    } else if (p.mode == 0) {
        const int cb = index_bits(k_lin);
        if (j.prune)
            rc = nnc_compress_f32(c, j.w, j.n, j.threshold, p.std_smooth, p.threshold_mode, 1, j.mask, &j.thr, &j.n_pruned, nullptr, k_lin,
                                  300, 1e-4, NNC_KM_INIT_LINEAR, j.centers, j.centred, j.packed, cb, j.hist, &j.info);
        else
            rc = nnc_kmeans1d_f32(c, j.w, j.n, nullptr, k_lin, 300, 1e-4, NNC_KM_INIT_LINEAR, j.centers, j.centred, nullptr, nullptr,
                                  j.packed, cb, j.hist, &j.info);
        if (rc == NNC_OK) {
            j.k = k_lin;
            j.code_bits = cb;
        }
    } else {
        const int k = k_lin + 1, cb = index_bits(k);
        if (j.prune) rc = nnc_prune_f32(c, j.w, j.n, j.threshold, p.std_smooth, p.threshold_mode, j.mask, &j.thr, &j.n_pruned);
        float xnew[300];
        double cdf[300];
        if (rc == NNC_OK) rc = nnc_weight_cdf_f32(c, j.w, j.n, 1, xnew, cdf);
        if (rc == NNC_OK) {
            std::vector<float> space((size_t)k);
            density_init(k_lin, xnew, cdf, space.data());
            rc = nnc_kmeans1d_f32(c, j.w, j.n, space.data(), k, 300, 1e-4, 0, j.centers, j.centred, nullptr, nullptr, j.packed, cb, j.hist,
                                  &j.info);
        }
        if (rc == NNC_OK) {
            j.k = k;
            j.code_bits = cb;
        }
    }
    j.status = rc;
    if (rc != NNC_OK) {
        strncpy(j.error, nnc_last_error(), sizeof(j.error) - 1);
        j.error[sizeof(j.error) - 1] = 0;
    }
}

struct BatchPool {
    int device = 0;
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    // the batch in flight (guarded by mu, read by the workers after they saw the new generation)
    uint64_t gen = 0;
    nnc_tensor_job *jobs = nullptr;
    const int *order = nullptr;
    int count = 0, use_workers = 0, running = 0;
    BatchParams params{};
    cudaEvent_t start_ev = nullptr;
    std::atomic<long long> launches{0};  // kernel launches of the batch, all workers
    bool log = getenv("NNC_BATCH_LOG") != nullptr;

    void worker(int id) {
        nnc_ctx *c = nullptr;
        if (nnc_ctx_create(device, &c) != NNC_OK) c = nullptr;
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return gen != seen; });
                seen = gen;
                if (id >= use_workers) continue;  // not part of this batch
            }
            if (c) {
                cudaSetDevice(device);
                cudaStreamWaitEvent(c->stream, start_ev, 0);  // the tensors were produced on the caller's stream
            }
            // static deal, largest first, round robin: a worker sees the same sizes in every call of the same model, so its
            // workspace arena reaches its high-water mark once (a regrow is a cudaFree + cudaMalloc: a device-wide stall)
            const int64_t l0 = c ? c->total_launches : 0;
            for (int i = id; i < count; i += use_workers) {
                nnc_tensor_job &j = jobs[order[i]];
                const auto t0 = std::chrono::steady_clock::now();
                if (c) {
                    run_job(c, j, params);
                } else {
                    j.status = NNC_ERR_CUDA;
                    strncpy(j.error, "worker context could not be created", sizeof(j.error) - 1);
                }
                if (log) fprintf(stderr, "[nnc batch] worker %d tensor %d (%lld weights): %.3f ms, %d iterations\n", id, order[i], (long long)j.n,
                                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), j.info.n_iter);
            }
            if (c) launches.fetch_add(c->total_launches - l0);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--running == 0) cv_done.notify_all();
            }
        }
    }
};

std::mutex g_pools_mu;
std::map<int, BatchPool *> g_pools;  // one per device, never destroyed (its threads live as long as the process)

}  // namespace

}  // namespace nnc

using namespace nnc;

#define NNC_TRY try {
#define NNC_CATCH                                             \
    }                                                         \
    catch (const nnc::Error &e) {                             \
        cudaGetLastError();                                   \
        return e.code;                                        \
    }                                                         \
    catch (const std::exception &e) {                         \
        nnc::set_error("unexpected exception: %s", e.what()); \
        return NNC_ERR_INTERNAL;                              \
    }                                                         \
    return NNC_OK;

extern "C" int nnc_compress_many_f32(nnc_ctx *ctx, nnc_tensor_job *jobs, int count, int std_smooth, int threshold_mode, int bits, int mode,
                          int max_workers) {
    NNC_TRY
    if (!ctx || (count > 0 && !jobs) || count < 0) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_compress_many_f32: bad arguments");
    if (bits < 1 || bits > 9 || (mode != 0 && mode != 1)) NNC_FAIL(NNC_ERR_BAD_ARG, "nnc_compress_many_f32: bits = %d, mode = %d", bits, mode);
    if (count == 0) return NNC_OK;
    if (ctx->world > 1) NNC_FAIL(NNC_ERR_UNSUPPORTED, "nnc_compress_many_f32 works on whole tensors (one rank)");
    NNC_CUDA(cudaSetDevice(ctx->device));
    BatchPool *pool;
    {
        std::lock_guard<std::mutex> lk(g_pools_mu);
        auto it = g_pools.find(ctx->device);
        if (it == g_pools.end()) {
            pool = new BatchPool();
            pool->device = ctx->device;
            NNC_CUDA(cudaEventCreateWithFlags(&pool->start_ev, cudaEventDisableTiming));
            g_pools[ctx->device] = pool;
        } else {
            pool = it->second;
        }
    }
    const int want = std::max(1, std::min(std::min(max_workers > 0 ? max_workers : 8, 16), count));
    std::vector<int> order((size_t)count);
    for (int i = 0; i < count; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return jobs[a].n > jobs[b].n; });  // largest first
    {
        std::unique_lock<std::mutex> lk(pool->mu);
        pool->cv_done.wait(lk, [&] { return pool->running == 0; });  // one batch at a time per device
        while ((int)pool->threads.size() < want) {
            const int id = (int)pool->threads.size();
            pool->threads.emplace_back([pool, id] { pool->worker(id); });
            pool->threads.back().detach();
        }
        NNC_CUDA(cudaEventRecord(pool->start_ev, ctx->stream));
        pool->jobs = jobs;
        pool->order = order.data();
        pool->count = count;
        pool->use_workers = want;
        pool->running = want;
        pool->params = BatchParams{std_smooth, threshold_mode, bits, mode};
        pool->launches.store(0);
        pool->gen += 1;
        pool->cv_work.notify_all();
        pool->cv_done.wait(lk, [&] { return pool->running == 0; });
        pool->jobs = nullptr;
        ctx->total_launches += pool->launches.load();  // the workers' launches count for the calling context
    }
    // every job ends with a synchronize of its worker's stream: all results are complete here
    for (int i = 0; i < count; ++i)
        if (jobs[i].status != NNC_OK) NNC_FAIL(jobs[i].status, "tensor %d of %d: %s", i, count, jobs[i].error);
    NNC_CATCH
}
