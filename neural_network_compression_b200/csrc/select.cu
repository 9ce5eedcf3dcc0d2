// select.cu -- survivor selection and the inputs of centroid initialisation.
//
//   minmax_device          weight.min() / weight.max()                       (utility.py:207-208, 363-364)
//   hist_edges_device      31 half-open bins [steps[i], steps[i+1])            (utility.py:366-372)
//   compact_ordered_device flat[flat != 0], order preserving                  (trainer.py:55-59)
//   gather_device          flat[idx]  (forgy: np.random.choice)                (utility.py:224-226)
//
// The reference computes the 31 counts with 31 full sweeps; here it is one sweep with warp-private
// shared-memory counters.  Compaction is a single pass (read 4 B, write 4 B per survivor) with a
// decoupled look-back over per-tile survivor counts.
#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace nnc {

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) minmax_kernel(const float *w, int64_t n, int vec_ok, int skip_zeros, uint32_t *out_min,
                                                     uint32_t *out_max, unsigned long long *out_cnt) {
    uint32_t mn = 0xffffffffu, mx = 0u;
    unsigned long long cnt = 0;
    auto one = [&](float x) {
        if (x == 0.f) {
            if (skip_zeros) return;
            x = 0.f;  // -0.0 -> +0.0
        }
        if (x != x) return;  // NaN never wins a NumPy-style min/max here; callers reject non-finite input
        uint32_t o = f2ord(x);
        mn = min(mn, o);
        mx = max(mx, o);
        cnt++;
    };
    int64_t nvec = vec_ok ? (n >> 2) : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = ld_stream_f4(w + 4 * i);
        one(x.x);
        one(x.y);
        one(x.z);
        one(x.w);
    }
    for (int64_t i = (nvec << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        one(w[i]);
    mn = warp_min_u(mn);
    mx = warp_max_u(mx);
    cnt = warp_sum_ull(cnt);
    if (lane_id() == 0) {
        atomicMin(out_min, mn);
        atomicMax(out_max, mx);
        if (cnt) atomicAdd(out_cnt, cnt);
    }
}

void minmax_device(nnc_ctx *ctx, const float *d_w, int64_t n, int skip_zeros, float *mn, float *mx, int64_t *cnt) {
    struct R {
        uint32_t mn, mx;
        unsigned long long cnt;
    };
    R *d = arena_alloc_t<R>(ctx, 1);
    R init{0xffffffffu, 0u, 0ull};
    NNC_CUDA(cudaMemcpyAsync(d, &init, sizeof(R), cudaMemcpyHostToDevice, ctx->stream));
    int vec_ok = (reinterpret_cast<uintptr_t>(d_w) & 15u) == 0;
    int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 16, (n / 4 + 255) / 256 + 1);
    NNC_LAUNCH(ctx, minmax_kernel, grid, 256, 0, d_w, n, vec_ok, skip_zeros, &d->mn, &d->mx, &d->cnt);
    R h;
    NNC_CUDA(cudaMemcpyAsync(&h, d, sizeof(R), cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    *cnt = (int64_t)h.cnt;
    *mn = h.cnt ? ord2f(h.mn) : 0.f;
    *mx = h.cnt ? ord2f(h.mx) : 0.f;
}

// ---------------------------------------------------------------------------------------------
// counts[b] = #{ edges[b] <= v < edges[b+1] }
__global__ void __launch_bounds__(256) hist_edges_kernel(const float *w, int64_t n, int vec_ok, const float *edges,
                                                         int n_edges, int skip_zeros, unsigned long long *counts) {
    extern __shared__ uint32_t sm[];
    float *s_edges = reinterpret_cast<float *>(sm);
    const int nb = n_edges - 1;
    uint32_t *s_hist = sm + n_edges + warp_id() * nb;  // warp-private counters
    for (int i = threadIdx.x; i < n_edges; i += blockDim.x) s_edges[i] = edges[i];
    for (int i = threadIdx.x; i < nb * (int)(blockDim.x >> 5); i += blockDim.x) sm[n_edges + i] = 0;
    __syncthreads();
    auto one = [&](float x, bool live) {
        int b = -1;
        if (live && !(skip_zeros && x == 0.f)) {
            // upper_bound(x) - 1: largest b with edges[b] <= x
            int lo = 0, hi = n_edges;  // first index with edges[idx] > x lies in [lo, hi]
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (s_edges[mid] <= x)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            b = lo - 1;
            if (b >= nb) b = -1;  // x >= last edge: counted nowhere (the maximum falls in no bin)
        }
        // warp-aggregated increment
        uint32_t peers = __match_any_sync(0xffffffffu, b);
        if (b >= 0 && (int)(__ffs(peers) - 1) == lane_id()) s_hist[b] += __popc(peers);
        __syncwarp();
    };
    int64_t nvec = vec_ok ? (n >> 2) : 0;
    // warp-uniform trip counts so that match_any sees the whole warp
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t start = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t iters = (nvec + stride - 1) / stride;
    for (int64_t it = 0; it < iters; ++it) {
        int64_t i = start + it * stride;
        bool live = i < nvec;
        float4 x = live ? ld_stream_f4(w + 4 * i) : make_float4(0, 0, 0, 0);
        one(x.x, live);
        one(x.y, live);
        one(x.z, live);
        one(x.w, live);
    }
    int64_t rem = n - (nvec << 2);
    iters = (rem + stride - 1) / stride;
    for (int64_t it = 0; it < iters; ++it) {
        int64_t i = (nvec << 2) + start + it * stride;
        bool live = i < n;
        one(live ? w[i] : 0.f, live);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        unsigned long long t = 0;
        for (int wp = 0; wp < (int)(blockDim.x >> 5); ++wp) t += sm[n_edges + wp * nb + b];
        if (t) atomicAdd(&counts[b], t);
    }
}

void hist_edges_device(nnc_ctx *ctx, const float *d_w, int64_t n, const float *h_edges, int n_edges, int skip_zeros,
                       int64_t *h_counts) {
    const int nb = n_edges - 1;
    float *d_edges = arena_alloc_t<float>(ctx, n_edges);
    unsigned long long *d_counts = arena_alloc_t<unsigned long long>(ctx, nb);
    NNC_CUDA(cudaMemcpyAsync(d_edges, h_edges, sizeof(float) * n_edges, cudaMemcpyHostToDevice, ctx->stream));
    NNC_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * nb, ctx->stream));
    int vec_ok = (reinterpret_cast<uintptr_t>(d_w) & 15u) == 0;
    int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 8, (n / 4 + 255) / 256 + 1);
    size_t smem = sizeof(uint32_t) * (n_edges + (size_t)nb * 8);
    NNC_LAUNCH(ctx, hist_edges_kernel, grid, 256, smem, d_w, n, vec_ok, d_edges, n_edges, skip_zeros, d_counts);
    NNC_CUDA(cudaMemcpyAsync(h_counts, d_counts, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---------------------------------------------------------------------------------------------
// order-preserving non-zero compaction, single pass, decoupled look-back
//
// A tile is 4096 consecutive elements; warp w of the CTA owns elements [512 w, 512 (w+1)) of it as 4 rows of
// 128 (one float4 per lane and row), so survivors keep their order: (row, lane, component) is element order.
constexpr int CP_THREADS = 256;
constexpr int CP_WARPS = CP_THREADS / 32;
constexpr int CP_ROWS = 4;
constexpr int CP_TILE = CP_THREADS * CP_ROWS * 4;  // 4096 elements
constexpr unsigned long long CP_FLAG_AGG = 1ull << 62;
constexpr unsigned long long CP_FLAG_PFX = 2ull << 62;
constexpr unsigned long long CP_VAL_MASK = (1ull << 62) - 1;

// exclusive prefix of the tile: sum of the survivor counts of all earlier tiles (warp-parallel look-back)
__device__ __forceinline__ unsigned long long cp_lookback(unsigned long long *state, unsigned int tile, unsigned long long tot) {
    const int lane = lane_id();
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(&state[0], CP_FLAG_PFX | tot);
        return 0;
    }
    if (lane == 0) st_volatile_u64(&state[tile], CP_FLAG_AGG | tot);
    unsigned long long excl = 0;
    long long p = (long long)tile - 1;
    for (;;) {
        const long long idx = p - lane;
        unsigned long long v = CP_FLAG_PFX;  // tiles before the first contribute 0 and end the walk
        if (idx >= 0) {
            do {
                v = ld_volatile_u64(&state[idx]);
            } while ((v >> 62) == 0);
        }
        const unsigned pfx = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int stop = pfx ? __ffs(pfx) - 1 : 32;  // nearest predecessor that already holds an inclusive prefix
        unsigned long long c = lane <= stop ? (v & CP_VAL_MASK) : 0ull;
        c = warp_sum_ull(c);
        excl += c;
        if (pfx) break;
        p -= 32;
    }
    if (lane == 0) st_volatile_u64(&state[tile], CP_FLAG_PFX | (excl + tot));
    return excl;
}

__global__ void __launch_bounds__(CP_THREADS, 5) compact_kernel(const float *w, int64_t n, int vec_ok, float *out,
                                                                unsigned long long *state, unsigned int *ticket,
                                                                unsigned long long *total_out) {
    __shared__ unsigned int s_tile;
    __shared__ int s_warp_cnt[CP_WARPS];
    __shared__ unsigned long long s_base;
    const int64_t n_tiles = (n + CP_TILE - 1) / CP_TILE;
    const int lane = lane_id(), wid = warp_id();
    const uint32_t lt = (1u << lane) - 1u;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned int tile = s_tile;
        if ((int64_t)tile >= n_tiles) break;
        const int64_t warp_base = (int64_t)tile * CP_TILE + (int64_t)wid * (128 * CP_ROWS);
        float4 x[CP_ROWS];
#pragma unroll
        for (int r = 0; r < CP_ROWS; ++r) {
            const int64_t i = warp_base + r * 128 + lane * 4;
            if (vec_ok && i + 4 <= n) {
                x[r] = ld_stream_f4(w + i);
            } else {
                x[r].x = i < n ? w[i] : 0.f;
                x[r].y = i + 1 < n ? w[i + 1] : 0.f;
                x[r].z = i + 2 < n ? w[i + 2] : 0.f;
                x[r].w = i + 3 < n ? w[i + 3] : 0.f;
            }
        }
        // per lane and row: survivors (NaN != 0 is true: NaNs survive, like numpy), exclusive offsets in the warp
        int off[CP_ROWS], wcnt = 0;
#pragma unroll
        for (int r = 0; r < CP_ROWS; ++r) {
            const int c = (x[r].x != 0.f) + (x[r].y != 0.f) + (x[r].z != 0.f) + (x[r].w != 0.f);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            off[r] = wcnt + incl - c;
            wcnt += __shfl_sync(0xffffffffu, incl, 31);
        }
        (void)lt;
        if (lane == 0) s_warp_cnt[wid] = wcnt;
        __syncthreads();
        if (wid == 0) {
            unsigned long long tot = 0;
#pragma unroll
            for (int i = 0; i < CP_WARPS; ++i) tot += s_warp_cnt[i];
            const unsigned long long excl = cp_lookback(state, tile, tot);
            if (lane == 0) {
                s_base = excl;
                if ((int64_t)tile == n_tiles - 1) *total_out = excl + tot;
            }
        }
        __syncthreads();
        unsigned long long base = s_base;
        for (int i = 0; i < wid; ++i) base += s_warp_cnt[i];
#pragma unroll
        for (int r = 0; r < CP_ROWS; ++r) {
            float *dst = out + base + off[r];
            if (x[r].x != 0.f) *dst++ = x[r].x;
            if (x[r].y != 0.f) *dst++ = x[r].y;
            if (x[r].z != 0.f) *dst++ = x[r].z;
            if (x[r].w != 0.f) *dst++ = x[r].w;
        }
        __syncthreads();  // s_tile / s_warp_cnt reuse
    }
}

int64_t compact_ordered_device(nnc_ctx *ctx, const float *d_w, int64_t n, float *d_out) {
    const int64_t n_tiles = (n + CP_TILE - 1) / CP_TILE;
    unsigned long long *state = arena_alloc_t<unsigned long long>(ctx, n_tiles + 2);
    NNC_CUDA(cudaMemsetAsync(state, 0, sizeof(unsigned long long) * (n_tiles + 2), ctx->stream));
    unsigned int *ticket = reinterpret_cast<unsigned int *>(state + n_tiles);
    unsigned long long *total = state + n_tiles + 1;
    int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 5, n_tiles);
    int vec_ok = (reinterpret_cast<uintptr_t>(d_w) & 15u) == 0;
    NNC_LAUNCH(ctx, compact_kernel, grid, CP_THREADS, 0, d_w, n, vec_ok, d_out, state, ticket, total);
    unsigned long long h = 0;
    NNC_CUDA(cudaMemcpyAsync(&h, total, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int64_t)h;
}

// ---------------------------------------------------------------------------------------------
__global__ void gather_kernel(const float *w, const long long *idx, int m, float *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = w[idx[i]];
}

void gather_device(nnc_ctx *ctx, const float *d_w, int64_t n, const int64_t *h_idx, int m, float *h_out) {
    for (int i = 0; i < m; ++i)
        if (h_idx[i] < 0 || h_idx[i] >= n) NNC_FAIL(NNC_ERR_BAD_ARG, "gather: index %lld out of range", (long long)h_idx[i]);
    long long *d_idx = arena_alloc_t<long long>(ctx, m);
    float *d_out = arena_alloc_t<float>(ctx, m);
    NNC_CUDA(cudaMemcpyAsync(d_idx, h_idx, sizeof(long long) * m, cudaMemcpyHostToDevice, ctx->stream));
    NNC_LAUNCH(ctx, gather_kernel, (m + 255) / 256, 256, 0, d_w, d_idx, m, d_out);
    NNC_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace nnc
