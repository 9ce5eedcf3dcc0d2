/*
 * nnc.h -- C ABI of the B200-native prune + 1-D k-means weight-compression hot path.
 *
 * This is the drop-in boundary for the helpers of the reference's
 * neural_network_compression/common/utility.py (prune_weigth :134-163,
 * get_quantized_weight :172-240, get_weight_distribution :334-392) and for the three
 * call sites in common/trainer.py (_prune_parameters :177-193, _reset_pruned_parameters
 * :195-206, quantize :42-72).  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *  - Every entry point returns an int status (NNC_OK == 0).  nnc_last_error() returns a
 *    thread-local message for the last failing call.
 *  - Data pointers (`w`, `mask`, `labels`, ...) may be HOST or DEVICE pointers; the library
 *    detects which (cudaPointerGetAttributes) and stages host buffers through its own device
 *    workspace.  Small scalar/centroid outputs (`*_out`, `centers[k]`, `hist[k]`) are HOST
 *    pointers.  The caller owns every buffer; the library keeps no pointer after returning.
 *  - All calls are synchronous with respect to the host: results are complete on return.
 *    Work is enqueued on the context's stream (nnc_ctx_set_stream to share a caller stream).
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    NNC_ERR_CUDA.
 *  - One context per host thread / rank.  A context is not thread-safe.
 */
#ifndef NNC_H
#define NNC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNC_VERSION 100 /* 0.1.0 */

enum {
    NNC_OK = 0,
    NNC_ERR_BAD_ARG = 1,      /* null pointer, n <= 0, k out of range, ... */
    NNC_ERR_CUDA = 2,         /* CUDA runtime error (message has the cudaError string) */
    NNC_ERR_NOT_ENOUGH = 3,   /* fewer samples than clusters (reference guard utility.py:202-204) */
    NNC_ERR_NONFINITE = 4,    /* NaN/Inf in k-means input (sklearn check_array raises ValueError) */
    NNC_ERR_UNSUPPORTED = 5,  /* k > NNC_KMAX, bits out of range, n too large */
    NNC_ERR_INTERNAL = 6,     /* an internal invariant failed (speculation window, lookback, ...) */
    NNC_ERR_COMM = 7          /* multi-GPU exchange failed */
};

#define NNC_KMAX 1024 /* largest cluster count (density init emits 2^bits + 1 centroids: bits <= 9) */

typedef struct nnc_ctx nnc_ctx;

/* ---- context --------------------------------------------------------------------------- */
int nnc_version(void);
const char *nnc_last_error(void);
int nnc_ctx_create(int device, nnc_ctx **out);
void nnc_ctx_destroy(nnc_ctx *ctx);
/* Run on `cuda_stream` (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream) instead of
 * the context's own stream.  NULL restores the context's stream. */
int nnc_ctx_set_stream(nnc_ctx *ctx, void *cuda_stream);
/* Pre-size the context's device workspace (optional; it otherwise grows to the largest call seen). */
int nnc_ctx_reserve(nnc_ctx *ctx, size_t bytes);
/* Device-side timing on the context's stream (CUDA events), for benchmarks. */
int nnc_timer_start(nnc_ctx *ctx);
int nnc_timer_stop(nnc_ctx *ctx, float *ms_out);
/* Per-phase device times (ms) of the last nnc_prune_f32 / nnc_kmeans1d_f32 call and the number of
 * kernels the call launched.  `names_out` receives a ';'-separated list matching ms_out order. */
int nnc_last_profile(nnc_ctx *ctx, float *ms_out, int cap, int *n_out, const char **names_out,
                     int64_t *launches_out);

/* Benchmarks: when on, every kernel launch (or only those whose kernel name contains `name_filter`, when it
 * is not NULL / empty) is bracketed by CUDA events on the context's stream;
 * nnc_last_kernel_times returns "kernel:launches:total_ms;..." accumulated since the last
 * nnc_ctx_set_kernel_timing call.  nnc_ctx_total_launches: kernels launched over the context's lifetime. */
int nnc_ctx_set_kernel_timing(nnc_ctx *ctx, int on, const char *name_filter);
int nnc_last_kernel_times(nnc_ctx *ctx, const char **out);
int nnc_ctx_total_launches(nnc_ctx *ctx, int64_t *out);

/* ---- pruning (utility.py:134-163, trainer.py:177-206) ------------------------------------ */
/* np.mean / np.var / np.std of a float32 tensor, bit-exact with NumPy's float32 pairwise
 * reduction (numpy/_core/_methods.py:117-233). */
int nnc_stats_f32(nnc_ctx *ctx, const float *w, int64_t n, float *mean_out, float *var_out, float *std_out);

/* prune_weigth(original_weigth, threshold, std_smooth): thr = np.std(w) * threshold when
 * std_smooth, mask = |w| < thr (strict), w[mask] = 0 IN PLACE.  mask is one byte per element
 * (NumPy bool layout).  threshold_mode 0: `threshold` was a Python float/int (NEP 50 weak scalar,
 * float32 arithmetic); 1: it was a float64 NumPy scalar (float64 product and comparison). */
int nnc_prune_f32(nnc_ctx *ctx, float *w, int64_t n, double threshold, int std_smooth, int threshold_mode,
                  uint8_t *mask, double *thr_out, int64_t *n_pruned_out);

/* weights[mask] = 0 (trainer.py:204-205); also the masked-gradient apply. */
int nnc_mask_apply_f32(nnc_ctx *ctx, float *w, const uint8_t *mask, int64_t n);

/* ---- weight distribution / init inputs (utility.py:334-392, trainer.py:55-60) ------------- */
/* flat[flat != 0], order preserving (trainer.py:55-59).  `out` must hold n floats. */
int nnc_compact_nonzero_f32(nnc_ctx *ctx, const float *w, int64_t n, float *out, int64_t *n_nz_out);
/* min / max (and count) of all elements, or of the non-zero elements when skip_zeros != 0. */
int nnc_minmax_f32(nnc_ctx *ctx, const float *w, int64_t n, int skip_zeros, float *min_out, float *max_out,
                   int64_t *count_out);
/* counts[b] = #{ edges[b] <= v < edges[b+1] }, b in [0, n_edges-1), edges ascending, n_edges <= 1025
 * (the 31 half-open bins of utility.py:366-372). */
int nnc_hist_edges_f32(nnc_ctx *ctx, const float *w, int64_t n, const float *edges, int n_edges, int skip_zeros,
                       int64_t *counts_out);
/* get_weight_distribution(weight_matrix) -> (xnew[300] f32, cdf[300] f64); skip_zeros != 0 fuses
 * the survivor selection of trainer.py:55-60. */
int nnc_weight_cdf_f32(nnc_ctx *ctx, const float *w, int64_t n, int skip_zeros, float *xnew300, double *cdf300);
/* out[i] = w[idx[i]] (forgy init: the host draws idx with NumPy's RNG; utility.py:224-226). */
int nnc_gather_f32(nnc_ctx *ctx, const float *w, int64_t n, const int64_t *idx, int m, float *out);

/* ---- 1-D k-means weight sharing (utility.py:237-239 -> sklearn KMeans, lloyd) -------------- */
typedef struct {
    int n_iter;           /* KMeans.n_iter_ */
    int strict;           /* stopped because the labelling did not change */
    int n_relocations;    /* empty clusters relocated over the run */
    int fixed_exp;        /* exponent E of the fixed-point image (max|x - mean| < 2^E) */
    float mean;           /* float32 mean used for centring (NumPy pairwise, bit-exact) */
    float tol;            /* absolute tolerance on sum(center_shift^2) */
    double inertia;       /* KMeans.inertia_ */
    int64_t n_nonzero;    /* survivors that were sorted */
} nnc_kmeans_info;

#define NNC_KM_INERTIA 1     /* also compute KMeans.inertia_ (costs arithmetic in the emission pass) */
#define NNC_KM_INIT_LINEAR 2 /* init = np.linspace(w.min(), w.max(), k) in float32 (utility.py:206-209); `init` may be NULL */
#define NNC_KM_MASK_BITS 4   /* nnc_compress_f32: `mask` receives ceil(n/8) bytes, bit i of byte i/8 (the 1-bit mask of the
                              * compressed-layer format) instead of one byte per weight */

/* KMeans(n_clusters=k, init=init, n_init=1, algorithm="lloyd", max_iter, tol).fit(w.reshape(-1,1)).
 * Outputs (any may be NULL): centers[k] = cluster_centers_; centred[k] = centres in sklearn's
 * mean-centred space; labels[n] int32 = labels_; ris[n] = cluster_centers_[labels_];
 * packed = n-bit codes, code i in bits [i*bits,(i+1)*bits) of a little-endian byte stream
 * (ceil(n*bits/8) bytes, bits >= ceil(log2 k)); hist[k] = code histogram. */
int nnc_kmeans1d_f32(nnc_ctx *ctx, const float *w, int64_t n, const float *init, int k, int max_iter, double tol,
                     int flags, float *centers, float *centred, int32_t *labels, float *ris, uint8_t *packed, int bits,
                     int64_t *hist, nnc_kmeans_info *info);

/* prune_weigth followed by the k-means weight sharing of the pruned tensor, keeping only the compressed form
 * (mask, codebook, packed codes, histogram): what Trainer._prune_parameters + Trainer.quantize do to one tensor
 * (trainer.py:177-193, :42-72).  A host tensor is copied to the device once; write_back == 0 skips copying the
 * pruned weights back to a HOST `w` (a device-resident `w` is always pruned in place).  Arguments as in
 * nnc_prune_f32 and nnc_kmeans1d_f32.  With std_smooth the k-means prologue of the pruned tensor (sklearn's centring
 * mean, survivor compaction, min / max) rides on the pruning pass: two sweeps of the tensor before the clustering
 * instead of three.  Results are bit-identical to nnc_prune_f32 followed by nnc_kmeans1d_f32. */
int nnc_compress_f32(nnc_ctx *ctx, float *w, int64_t n, double threshold, int std_smooth, int threshold_mode, int write_back,
                     uint8_t *mask, double *thr_out, int64_t *n_pruned_out, const float *init, int k, int max_iter, double tol,
                     int flags, float *centers, float *centred, uint8_t *packed, int bits, int64_t *hist, nnc_kmeans_info *info);

/* E-step / emission only: labels[i] = first argmin_j fl(c_j^2 + fl(-2 x'_i) c_j), x' = fl(w - mean),
 * c = centred[k] (the sklearn label rule, _k_means_lloyd.pyx:196-213).  values[k] (optional) are
 * the codebook entries written to `ris`. */
int nnc_assign_f32(nnc_ctx *ctx, const float *w, int64_t n, const float *centred, int k, float mean,
                   const float *values, int32_t *labels, float *ris, uint8_t *packed, int bits, int64_t *hist,
                   double *inertia_out);

/* Codebook de-quantisation: out[i] = values[code_i] from packed n-bit codes. */
int nnc_unpack_gather_f32(nnc_ctx *ctx, const uint8_t *packed, int64_t n, int bits, const float *values, int k,
                          float *out);

/* 0/1 bytes -> bits (bit i of byte i/8 = src[i] != 0): packs a pruning mask (the reference keeps it as a NumPy bool array,
 * common/trainer.py:25,192) for the compressed-layer format of common/storage.py. */
int nnc_pack_bits_u8(nnc_ctx *ctx, const uint8_t *src, int64_t n, uint8_t *dst_bits);

/* Trained-quantization gradient sum (papers/lat/report.tex:152): out[j] = sum_i grad[i] * [code_i == j].
 * codes: packed n-bit stream when bits > 0, int32 labels when bits == 0. */
int nnc_grad_segsum_f32(nnc_ctx *ctx, const float *grad, const void *codes, int64_t n, int bits, int k,
                        double *out);

/* ---- all tensors of a model in one call (le_net_5.py:17-34 / le_net_300.py: the trainer walks the layers, pruning
 * every kernel and bias, trainer.py:177-193, then quantising every array, trainer.py:50-70) -----------------------
 * One job per tensor; the jobs run concurrently on a pool of native worker threads, each with a context and a stream of
 * its own (LeNet tensors have 10 .. 627 200 weights: every one of them is bound by launch latency and host round trips,
 * which overlap across the workers without an interpreter in between).  Results are bit-identical to the per-tensor
 * calls: nnc_compress_f32 for mode 0 (linear init), nnc_prune_f32 + nnc_weight_cdf_f32 + the density init of
 * utility.py:210-223 + nnc_kmeans1d_f32 for mode 1 (density init: 2^bits + 1 centroids, as the reference has it).
 * A tensor with fewer than 2^bits + 1 weights is pruned only (k = 0), like get_quantized_weight (utility.py:202-204). */
typedef struct {
    float *w;             /* in: n weights, host or device; pruned in place when prune != 0 */
    int64_t n;
    double threshold;     /* pruning quality parameter of this tensor */
    int prune;            /* 0: the tensor is taken as it is */
    int pad_;
    uint8_t *mask;        /* out: n bytes (1 = pruned), same memory space as w; may be NULL when prune == 0 */
    float *centers;       /* out: k floats (host) */
    float *centred;       /* out: k floats (host) */
    uint8_t *packed;      /* out: ceil(n * code_bits / 8) bytes, same memory space as w */
    int64_t *hist;        /* out: k counts (host) */
    nnc_kmeans_info info; /* out */
    double thr;           /* out: the threshold applied */
    int64_t n_pruned;     /* out */
    int k;                /* out: clusters (0: not quantised) */
    int code_bits;        /* out: bits per packed code */
    int status;           /* out: NNC_OK or the error of this tensor */
    char error[196];      /* out: its message */
} nnc_tensor_job;

/* mode: 0 linear, 1 density.  max_workers <= 0: the default (8).  Returns NNC_OK when every job succeeded, else the first
 * failing job's status (all jobs are attempted).  ctx gives the device and the stream the tensors were produced on. */
int nnc_compress_many_f32(nnc_ctx *ctx, nnc_tensor_job *jobs, int count, int std_smooth, int threshold_mode, int bits, int mode,
                          int max_workers);

/* ---- multi-GPU (one process per GPU; contiguous shards of the flattened tensor) ------------ */
/* Multi-GPU: tells the context the element count of the WHOLE tensor the next sharded calls work on (0: ask the ranks
 * with an all-reduce at the start of every call, the default).  All ranks must give the same value.  Saves one
 * all-reduce and one host round trip per call; the reference has no counterpart (single process). */
int nnc_ctx_hint_global_n(nnc_ctx *ctx, int64_t n_global);

/* The library does not own a communicator.  The host supplies an all-reduce callback that reduces
 * `count` int64 values (DEVICE buffer, in place; op 0 sum, 1 min, 2 max) across ranks on `stream`;
 * all exchanged quantities are integers, so the result is bit-identical for any rank count.
 * With world > 1, nnc_stats_f32 / nnc_prune_f32 / nnc_kmeans1d_f32 / nnc_compress_f32 take THIS RANK'S
 * slice of the flattened tensor (`n` = its length), which must be the slice nnc_shard_range assigns:
 * shards are aligned to the tiles of NumPy's pairwise-summation tree so that the float32 mean / std
 * are bit-exact for any rank count.  Per-rank outputs (mask, codes, labels) cover the slice; scalars,
 * centroids and histograms are global and identical on every rank.  Only the seeded/explicit and
 * linear initialisations are supported on shards (density's CDF helpers stay single-rank). */
int nnc_shard_range(int64_t n, int rank, int world, int64_t *begin, int64_t *end);
/* Preferred transport: the library's own NCCL communicator (libnccl.so.2 bound at run time).  Rank 0 calls
 * nnc_comm_unique_id, the host ships the 128 bytes to every rank (any side channel), all ranks call
 * nnc_ctx_init_nccl.  All-reduces are then enqueued natively on the context's stream (no host callback). */
/* Optional, one box: peer mailboxes for the Lloyd loop.  Every rank creates one (64-byte CUDA IPC handle out),
 * the host all-gathers the handles (rank order, world x 64 bytes) and every rank connects.  The per-iteration
 * exchanges of nnc_kmeans1d_f32 / nnc_compress_f32 then happen INSIDE the update kernel over NVLink peer memory
 * instead of two NCCL calls and two extra launches per iteration. */
int nnc_peer_mailbox_create(nnc_ctx *ctx, int world, char *handle_out64);
int nnc_peer_mailbox_connect(nnc_ctx *ctx, const char *handles, int rank, int world);
int nnc_comm_unique_id(char *out128);
int nnc_ctx_init_nccl(nnc_ctx *ctx, const char *id128, int rank, int world);
typedef int (*nnc_allreduce_i64_fn)(void *user, int64_t *dev_buf, int count, int op /*0 sum,1 min,2 max*/,
                                    void *cuda_stream);
int nnc_ctx_set_comm(nnc_ctx *ctx, int rank, int world, nnc_allreduce_i64_fn fn, void *user);

#ifdef __cplusplus
}
#endif
#endif /* NNC_H */
