// internal.h -- host-side plumbing shared by the .cu translation units (context, workspace arena,
// host/device pointer staging, error handling, launch accounting).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/nnc.h"

struct NcclUniqueIdBytes {  // ncclUniqueId (nccl.h: 128 opaque bytes), passed by value to ncclCommInitRank
    char internal[128];
};

namespace nnc {

void set_error(const char *fmt, ...);
bool debug_sync();  // NNC_DEBUG_SYNC: synchronize and check after every kernel launch (fault localisation)

struct Error {
    int code;
};

#define NNC_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            nnc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));   \
            throw nnc::Error{NNC_ERR_CUDA};                                                         \
        }                                                                                           \
    } while (0)

#define NNC_FAIL(code, ...)          \
    do {                             \
        nnc::set_error(__VA_ARGS__); \
        throw nnc::Error{code};      \
    } while (0)

// Scalars that flow between kernels of one call without a host round trip.  One instance lives in
// device memory (ctx->d_scal); the host reads it back once at the end of a call.
struct DevScalars {
    // numpy-exact statistics
    float tree_sum;      // result of the last pairwise tree reduction
    float mean;          // fl32(sum / n)
    float var;           // fl32(sum((x-mean)^2) / n)
    float std_;          // sqrtf(var)
    // fp64 side statistics (order dependent, estimates only)
    double sum_d, sumsq_d;
    unsigned long long n_nonfinite;
    // pruning
    double thr;          // exact threshold used by the final comparison
    double band_lo, band_hi;
    unsigned long long n_pruned;
    unsigned long long band_count;    // entries appended to the side list
    unsigned long long band_dropped;  // entries that did not fit
    int spec_failed;     // exact threshold fell outside the speculation band
    // quantization prologue
    uint32_t min_ord, max_ord;        // ordered-uint min / max over all elements
    uint32_t min_nz_ord, max_nz_ord;  // ... over non-zero elements
    uint32_t amax_bits;               // largest |x| bit pattern (>= 0x7f800000: a NaN or infinity is present)
    uint32_t amax_all;                // ... over ALL elements of the tensor the statistics pass read (VisitStats)
    uint32_t amin_nz_m1;              // smallest non-zero |x| bit pattern, minus one (0xffffffff: all zero)
    float fmin_all, fmax_all;         // float min / max over all elements (scratch; folded into min_ord / max_ord)
    unsigned long long n_nz;          // non-zero count (global once the ranks have exchanged)
    unsigned long long n_nz_local;    // ... of this rank's shard
    int pad_;
};

struct PhaseTimer {
    std::vector<cudaEvent_t> ev;
    std::vector<std::string> names;
    int used = 0;
};

}  // namespace nnc

struct nnc_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // bump arena for per-call device scratch
    char *ws = nullptr;
    size_t ws_bytes = 0;
    size_t ws_off = 0;
    size_t call_bytes = 0;      // bytes requested by the current call
    size_t high_water = 0;      // largest call so far: the main block is regrown to this at the next reset
    std::vector<void *> overflow;  // extra blocks taken when the main block was too small
    int64_t desc_n = -1;           // reduction-tree tile descriptors of an n-element tensor (reduce_np.cu): they only depend
    void *desc_ptr = nullptr;      // on n, so they are kept across calls in an allocation of their own
    size_t desc_bytes = 0;
    bool user_stream = false;
    // optional per-kernel CUDA-event timing (benchmarks): one event pair per launch, folded by kernel name
    bool ktime = false;
    std::string kfilter;  // when not empty: only launches whose kernel name contains it are timed
    std::vector<cudaEvent_t> kev;
    std::vector<const char *> knames;
    size_t kused = 0;
    std::vector<std::string> kacc_names;  // running totals since nnc_ctx_set_kernel_timing
    std::vector<double> kacc_ms;
    std::vector<long long> kacc_cnt;
    std::string ktimes;  // "name:launches:total_ms;..."
    int64_t total_launches = 0;  // kernels launched over the context's lifetime
    nnc::DevScalars *d_scal = nullptr;
    nnc::DevScalars *h_scal = nullptr;  // pinned mirror
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    nnc::PhaseTimer prof;
    std::vector<float> prof_ms;
    std::string prof_names;
    int64_t launches = 0;       // kernels launched by the current / last call
    int64_t last_launches = 0;
    // multi-GPU: shard of the current call (set by shard_setup at the entry point)
    struct {
        int64_t n_global = 0, begin = 0;
        uint32_t t0 = 0, t1 = 0;
    } sh;
    int rank = 0, world = 1;
    int64_t hint_n_global = 0;  // > 0: the sharded calls take this as the size of the whole tensor (nnc_ctx_hint_global_n)
    void *nccl_comm = nullptr;  // ncclComm_t when the library owns a communicator (nnc_ctx_init_nccl)
    // peer mailboxes for the in-kernel exchanges of the Lloyd loop (peer.cuh; nnc_peer_mailbox_create / _connect)
    bool peer_enabled = false;
    void *peer_local = nullptr;
    void *peer_mail[16] = {nullptr};
    unsigned long long peer_seq = 0;
    int *d_comm_error = nullptr;  // set by an in-kernel exchange that timed out (reduce_np.cu: np_final_peer_kernel)
    nnc_allreduce_i64_fn allreduce = nullptr;
    void *allreduce_user = nullptr;
    // kernels whose dynamic shared-memory limit has been raised ON THIS CONTEXT'S DEVICE (the attribute is per device)
    std::vector<std::pair<const void *, size_t>> func_smem;
    int fast_cluster = 0;  // CTAs per cluster of the Lloyd cluster kernel on this device (0: not decided yet)
};

namespace nnc {

// ---- arena -----------------------------------------------------------------------------------
void arena_reset(nnc_ctx *ctx);
void arena_reserve(nnc_ctx *ctx, size_t bytes);  // make sure `bytes` fit (may reallocate; only when empty)
void *arena_alloc(nnc_ctx *ctx, size_t bytes);
template <class T>
T *arena_alloc_t(nnc_ctx *ctx, size_t count) {
    return reinterpret_cast<T *>(arena_alloc(ctx, count * sizeof(T)));
}

bool is_device_ptr(const void *p);

// cudaFuncAttributeMaxDynamicSharedMemorySize >= bytes for `fn` on the context's device; set once per context (and again
// when a larger size is asked for).  The attribute is per DEVICE: a process-wide flag would leave the kernels of a second
// GPU at the 48 KB default.
void func_dyn_smem(nnc_ctx *ctx, const void *fn, size_t bytes);

// Stages a caller buffer: device pointers pass through, host pointers get an arena copy.
struct Staged {
    void *dev = nullptr;
    void *host = nullptr;  // non-null when a copy back / in is needed
    size_t bytes = 0;
};
Staged stage_in(nnc_ctx *ctx, const void *p, size_t bytes);    // read-only input
Staged stage_out(nnc_ctx *ctx, void *p, size_t bytes);         // output only
Staged stage_inout(nnc_ctx *ctx, void *p, size_t bytes);       // in place
void stage_finish(nnc_ctx *ctx, const Staged &s);              // async D2H if host-backed

// ---- phases / launches -------------------------------------------------------------------------
void prof_begin(nnc_ctx *ctx);
void prof_mark(nnc_ctx *ctx, const char *name);  // ends the phase `name` that started at the previous mark
void prof_end(nnc_ctx *ctx);

void klaunch_begin(nnc_ctx *ctx, const char *name);
void klaunch_end(nnc_ctx *ctx);

#define NNC_LAUNCH(ctx, kernel, grid, block, smem, ...)                 \
    do {                                                                \
        const bool _kt = (ctx)->ktime && ((ctx)->kfilter.empty() || strstr(#kernel, (ctx)->kfilter.c_str())); \
        if (_kt) nnc::klaunch_begin((ctx), #kernel);                    \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
        (ctx)->launches++;                                              \
        if (_kt) nnc::klaunch_end((ctx));                               \
        NNC_CUDA(cudaGetLastError());                                   \
        if (nnc::debug_sync()) {                                        \
            cudaError_t _se = cudaStreamSynchronize((ctx)->stream);     \
            if (_se != cudaSuccess) NNC_FAIL(NNC_ERR_CUDA, "kernel %s failed: %s", #kernel, cudaGetErrorString(_se)); \
        }                                                               \
    } while (0)

// the same with an explicit name for the per-kernel timing (template kernels: one name per instantiation)
#define NNC_LAUNCH_AS(ctx, name, kernel, grid, block, smem, ...)         \
    do {                                                                \
        const bool _kt = (ctx)->ktime && ((ctx)->kfilter.empty() || strstr((name), (ctx)->kfilter.c_str())); \
        if (_kt) nnc::klaunch_begin((ctx), (name));                     \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
        (ctx)->launches++;                                              \
        if (_kt) nnc::klaunch_end((ctx));                               \
        NNC_CUDA(cudaGetLastError());                                   \
    } while (0)

void read_scalars(nnc_ctx *ctx);  // D2H of DevScalars + stream sync

// In-place all-reduce of `count` int64 values in device memory over the ranks of the context (no-op for one rank).
// op: 0 sum, 1 min, 2 max.  Enqueued on the context's stream through the host's callback.
void comm_allreduce(nnc_ctx *ctx, int64_t *d_buf, int count, int op);

// ---- kernels' host entry points (one per .cu file) ----------------------------------------------
// reduce_np.cu
void np_shard_range(int64_t n, int rank, int world, int64_t *begin, int64_t *end, uint32_t *t0, uint32_t *t1);
struct NpPlan {
    int depth;
    uint32_t num_tiles;
};
NpPlan np_plan(int64_t n);
size_t np_partials_bytes(const NpPlan &p);
void np_stats(nnc_ctx *ctx, const float *d_w, int64_t n);  // fills mean/var/std_ in d_scal (2 passes)
// fuse (nnc_compress_f32): the k-means prologue of the PRUNED tensor rides on the apply pass -- survivors compacted
// into fuse->out, mean / min / max / key range / survivor count left in DevScalars exactly as quant_prologue would
// leave them.  fuse->done tells whether that happened (not for a hard threshold or when the speculation failed).
struct QuantFuse {
    float *out = nullptr;
    int64_t capacity = 0;
    bool done = false;
};
void prune_device(nnc_ctx *ctx, float *d_w, int64_t n, double q, int std_smooth, int thr_mode, uint8_t *d_mask,
                  QuantFuse *fuse = nullptr);
void mask_apply_device(nnc_ctx *ctx, float *d_w, const uint8_t *d_mask, int64_t n);
// k-means prologue: NumPy mean + min/max + non-zero count + |x| key range (DevScalars), and the non-zero elements
// themselves written densely (unordered) into d_out -- one read of the tensor.
void quant_prologue(nnc_ctx *ctx, const float *d_w, int64_t n, float *d_out, int64_t capacity);

// scan.cu : multi-CTA exclusive scans (out[n] = total)
void exclusive_scan_i64(nnc_ctx *ctx, const long long *d_in, long long n, long long *d_out);
void exclusive_scan_u32_u64(nnc_ctx *ctx, const unsigned int *d_in, long long n, unsigned long long *d_out);

// select.cu : min/max, edge histogram, ordered compaction, gather
void minmax_device(nnc_ctx *ctx, const float *d_w, int64_t n, int skip_zeros, float *mn, float *mx, int64_t *cnt);
void hist_edges_device(nnc_ctx *ctx, const float *d_w, int64_t n, const float *h_edges, int n_edges, int skip_zeros,
                       int64_t *h_counts);
int64_t compact_ordered_device(nnc_ctx *ctx, const float *d_w, int64_t n, float *d_out);
void gather_device(nnc_ctx *ctx, const float *d_w, int64_t n, const int64_t *h_idx, int m, float *h_out);

// sort.cu : onesweep LSD radix sort of float32 keys; returns pointer to the sorted buffer (d_a or d_b)
float *radix_sort_f32(nnc_ctx *ctx, float *d_a, float *d_b, int64_t n, uint32_t amin, uint32_t amax);

// khist.cu : the sorted survivors as ascending (distinct value, multiplicity) runs from a full-resolution key histogram
struct SortedRuns {
    const float *val = nullptr;
    const uint32_t *cnt = nullptr;
    int64_t n_ent = 0;
};
bool hist_sort_applicable(int64_t n, uint32_t amin, uint32_t amax);
SortedRuns hist_sort_f32(nnc_ctx *ctx, float *d_a, float *d_b, int64_t n, uint32_t amin, uint32_t amax);

// lloyd.cu
struct LloydResult {
    int n_iter, strict, n_reloc, fixed_exp;
    float tol;
};
struct LloydDevice;  // opaque device-side state
struct LloydHandle {
    LloydDevice *d_state = nullptr;
    const float *d_sorted = nullptr;    // surviving weights ascending: one entry per sample, or (d_cnt) per distinct value
    const uint32_t *d_cnt = nullptr;    // multiplicity of every entry (nullptr: one sample per entry)
    int64_t n_ent = 0;                  // entries (with d_cnt only; otherwise n_nz)
    int64_t n_nz = 0, n0 = 0, n = 0;
    int k = 0;
    float mean = 0.f;
    void *d_ptile = nullptr, *d_samp = nullptr;
    int64_t n_tiles = 0;
};
// h_hist (optional, k entries): code histogram of the final labelling, counted on the sorted survivors
LloydResult lloyd_run(nnc_ctx *ctx, LloydHandle &h, const float *h_init, int max_iter, double tol_rel,
                      float *h_centred_final, float *h_centred_emit, int64_t *h_hist);

// emit.cu : final E-step over the tensor in original order
// h_centred: centroids the labels are taken against; h_centred_final (optional): final centroids (codebook
// values and inertia); they differ only after a strict stop in which relocation fired.
// xabs / xlo / xhi: max |w - mean|, min and max of (w - mean) over the tensor when the caller already knows
// them; xabs negative to have them measured here.
void emit_device(nnc_ctx *ctx, const float *d_w, int64_t n, const float *h_centred, const float *h_centred_final, int k,
                 float mean, float xabs, float xlo, float xhi, const float *h_values, int32_t *d_labels, float *d_ris, uint8_t *d_packed, int bits,
                 int64_t *h_hist, double *h_inertia);
void pack_bits_device(nnc_ctx *ctx, const uint8_t *d_src, int64_t n, uint8_t *d_dst);
void unpack_gather_device(nnc_ctx *ctx, const uint8_t *d_packed, int64_t n, int bits, const float *h_values, int k,
                          float *d_out);
// segsum.cu
void grad_segsum_device(nnc_ctx *ctx, const float *d_grad, const void *d_codes, int64_t n, int bits, int k,
                        double *h_out);

}  // namespace nnc
