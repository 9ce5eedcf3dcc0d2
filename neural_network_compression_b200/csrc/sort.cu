// sort.cu -- hand-written "onesweep" least-significant-digit radix sort of float32 keys (keys only).
//
// The reference clusters with n x k distance evaluations per Lloyd iteration
// (sklearn/cluster/_k_means_lloyd.pyx:196-213).  For 1-D data every cluster is an interval of the sorted
// samples, so the B200 path sorts the surviving (non-zero) weights ONCE and turns each Lloyd iteration into
// boundary searches (lloyd.cu).  This file is that one-time sort.
//
// Algorithm: least-significant-digit radix sort with 9-bit digits over range-compressed keys (three passes
// for a pruned Gaussian layer).  Each pass is count -> base -> scatter (see "pass structure" below): 12 B of
// traffic per key and pass, and no inter-CTA dependency (a decoupled look-back "onesweep" variant was measured
// first and was bound by the look-back latency, not by HBM).  Floats are mapped to keys on the first read and
// back on the last write.
#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace nnc {

constexpr int RS_THREADS = 512;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 8192 keys
constexpr int RS_RADIX = 512;                   // bins of the widest digit (9 bits)
constexpr int RS_MAX_PASSES = 4;

// Sort key: the surviving weights are non-zero floats with |x| bit patterns in [amin, amin + range].  Instead of
// the full 32-bit order-preserving image, the key is the RANK IMAGE within that range,
//     negative x: range - (|x|_bits - amin)          positive x: range + 1 + (|x|_bits - amin),
// which needs only ceil(log2(2 range + 2)) bits -- 26 for a pruned N(0, sigma^2) layer -- so three 9-bit passes
// replace four 8-bit ones.
struct RsKeyMap {
    uint32_t amin, range;
};
__device__ __forceinline__ uint32_t rs_key(uint32_t bits, RsKeyMap km) {
    const uint32_t m = (bits & 0x7fffffffu) - km.amin;
    return (bits & 0x80000000u) ? km.range - m : km.range + 1u + m;
}
__device__ __forceinline__ uint32_t rs_unkey(uint32_t key, RsKeyMap km) {
    return key <= km.range ? ((km.range - key + km.amin) | 0x80000000u) : (key - km.range - 1u + km.amin);
}

struct RsPlan {
    int passes;
    int shift[RS_MAX_PASSES];
    int width[RS_MAX_PASSES];
};

// ---- pass structure --------------------------------------------------------------------------------------
// The keys are cut into C contiguous chunks of whole tiles (8192 keys), one per resident CTA.  Per pass:
//   count    CTA c histograms the pass digit over its chunk                          (read 4 B/key)
//   base     one CTA turns the C x radix chunk histograms into the global start of every (chunk, digit) run
//   scatter  CTA c walks its chunk tile by tile: stable ranking inside the tile (warp-level match_any
//            multisplit), tile staged in digit order in shared memory, coalesced scatter; the running
//            (chunk, digit) cursors live in shared memory                                (read + write 4 B/key)
// No CTA ever waits for another one (no decoupled look-back): every kernel is a plain streaming pass.
constexpr int RS_MAX_CHUNKS = 1024;

template <bool IN_FLOAT>
__global__ void __launch_bounds__(RS_THREADS) rs_count_kernel(const uint32_t *__restrict__ in, int64_t n, int vec_ok,
                                                              int64_t tiles_per_chunk, int shift, int width, RsKeyMap km,
                                                              uint32_t *chunk_hist /*[C][RS_RADIX]*/) {
    __shared__ uint32_t h[RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t dmask = (1u << width) - 1u;
    const int64_t lo = (int64_t)blockIdx.x * tiles_per_chunk * RS_TILE;
    const int64_t hi = min(n, lo + tiles_per_chunk * RS_TILE);
    auto one = [&](uint32_t bits) {
        const uint32_t k = IN_FLOAT ? rs_key(bits, km) : bits;
        atomicAdd(&h[(k >> shift) & dmask], 1u);
    };
    if (vec_ok) {  // lo is a multiple of 8192 keys: 16-byte aligned whenever the base is
        const int64_t nvec = (hi - lo) >> 2;
        const uint32_t *p = in + lo;
        for (int64_t i = threadIdx.x; i < nvec; i += RS_THREADS) {
            uint4 v = ld_stream_u4(p + 4 * i);
            one(v.x);
            one(v.y);
            one(v.z);
            one(v.w);
        }
        for (int64_t i = lo + (nvec << 2) + threadIdx.x; i < hi; i += RS_THREADS) one(in[i]);
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += RS_THREADS) one(in[i]);
    }
    __syncthreads();
    chunk_hist[(size_t)blockIdx.x * RS_RADIX + threadIdx.x] = h[threadIdx.x];
}

// chunk_hist[c][d] (counts)  ->  chunk_base[c][d] = keys with a smaller digit + keys with digit d in earlier chunks
__global__ void __launch_bounds__(RS_RADIX) rs_base_kernel(const uint32_t *chunk_hist, int chunks, unsigned long long *chunk_base) {
    __shared__ unsigned long long s_warp[RS_RADIX / 32];
    const int d = threadIdx.x, lane = lane_id(), w = warp_id();
    unsigned long long tot = 0;
    for (int c = 0; c < chunks; ++c) tot += chunk_hist[(size_t)c * RS_RADIX + d];
    unsigned long long incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    unsigned long long add = 0;
    for (int i = 0; i < w; ++i) add += s_warp[i];
    unsigned long long run = incl - tot + add;  // keys with a smaller digit
    for (int c = 0; c < chunks; ++c) {
        chunk_base[(size_t)c * RS_RADIX + d] = run;
        run += chunk_hist[(size_t)c * RS_RADIX + d];
    }
}

struct RsSmem {
    alignas(16) uint32_t stage[RS_TILE];  // next tile's raw keys, filled by cp.async while the current tile is processed
    uint32_t keys[RS_TILE];
    uint32_t warp_hist[RS_WARPS][RS_RADIX];
    uint32_t tile_start[RS_RADIX];
    unsigned long long run[RS_RADIX];    // next output position of every digit for this chunk
    unsigned long long gbase[RS_RADIX];  // run[d] - tile_start[d] of the tile being scattered
    uint32_t warp_tot[RS_WARPS];
};

template <bool IN_FLOAT, bool OUT_FLOAT>
__global__ void __launch_bounds__(RS_THREADS, 2) rs_scatter_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                                int64_t n, int64_t tiles_per_chunk, int shift, int width,
                                                                RsKeyMap km, const unsigned long long *__restrict__ chunk_base) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem &s = *reinterpret_cast<RsSmem *>(smem_raw);
    const int lane = lane_id(), wid = warp_id();
    const uint32_t dmask = (1u << width) - 1u;
    const int radix = 1 << width;
    const uint32_t lt = (1u << lane) - 1u;
    s.run[threadIdx.x] = chunk_base[(size_t)blockIdx.x * RS_RADIX + threadIdx.x];
    const int64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    const int64_t t0 = (int64_t)blockIdx.x * tiles_per_chunk, t1 = min(n_tiles, t0 + tiles_per_chunk);
    const bool src_aligned = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
    // Asynchronous tile fetch (cp.async, 16 B per thread and copy): the fetch of tile t + 1 is in flight while
    // tile t is ranked and scattered -- without it the pass waited on the key loads of every tile (ncu:
    // long-scoreboard stalls on the first use of the keys, 45 % of the samples).
    auto fetch = [&](int64_t tile) {
        const int64_t tile_base = tile * RS_TILE;
        if (src_aligned && tile_base + RS_TILE <= n) {
#pragma unroll
            for (int c = 0; c < RS_TILE / 4 / RS_THREADS; ++c) {
                const int chunk = c * RS_THREADS + threadIdx.x;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.stage[4 * chunk]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(in + tile_base + 4 * chunk) : "memory");
            }
        } else {  // the last, partial tile (or an unaligned source): plain loads
            for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) s.stage[i] = tile_base + i < n ? in[tile_base + i] : 0xffffffffu;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (t0 < t1) fetch(t0);
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int64_t tile_base = tile * RS_TILE;
        const int valid = (int)min((int64_t)RS_TILE, n - tile_base);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // the staged tile is complete and visible
        // ---- keys to registers (warp-striped: item i of lane l is element i*32 + l of the warp's 512-key chunk)
        uint32_t key[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const int e = wid * (32 * RS_ITEMS) + i * 32 + lane;
            uint32_t k = s.stage[e];
            if (e < valid) {
                if (IN_FLOAT) k = rs_key(k, km);
            } else {
                k = 0xffffffffu;  // padding: all digit bits set, sorts last in every pass
            }
            key[i] = k;
        }
        for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX / 4; i += RS_THREADS)
            reinterpret_cast<uint4 *>(&s.warp_hist[0][0])[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();  // stage consumed, counters cleared, previous tile's scatter finished
        if (tile + 1 < t1) fetch(tile + 1);
        // ---- stable ranking inside the warp with warp-private digit counters.
        // rank = (keys of the warp's earlier items with this digit) + (lower lanes of this item with this digit).
        // __match_any_sync would give the second term directly, but on sm_100 it costs ~60 cycles per warp
        // instruction when the 32 digits are mostly distinct (scripts/micro/matchany.cu) -- it alone made the pass
        // 15 k cycles per tile.  With 512 bins two lanes of an item rarely share a digit, so collisions are
        // DETECTED with the counter word itself -- every lane stores its lane id into the tag byte of its digit's
        // word, then reads the word back: a lane whose tag did not survive shares its digit with another lane --
        // and only those digit groups are resolved with ballots.  word = count (low 16 bits) | tag (top byte).
        uint32_t rank[RS_ITEMS];
        uint32_t *wh = s.warp_hist[wid];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const uint32_t d = (key[i] >> shift) & dmask;
            reinterpret_cast<volatile uint8_t *>(&wh[d])[3] = (uint8_t)lane;
            __syncwarp();
            const uint32_t word = reinterpret_cast<volatile uint32_t *>(wh)[d];
            uint32_t peers = 1u << lane;
            uint32_t unresolved = __ballot_sync(0xffffffffu, (word >> 24) != (uint32_t)lane);
            if (__popc(unresolved) > 6) {
                // many lanes share digits (the most significant digit of a bell-shaped layer): one ballot per digit
                // bit gives every lane its peer set at a fixed cost
                peers = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 9; ++b) {
                    if (b < width) {
                        const uint32_t v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                        peers &= ((d >> b) & 1u) ? v : ~v;
                    }
                }
            } else {
                while (unresolved) {  // one round per digit shared by several lanes (about one per item when digits are uniform)
                    const int ld = __ffs(unresolved) - 1;
                    const uint32_t dl = __shfl_sync(0xffffffffu, d, ld);
                    const uint32_t grp = __ballot_sync(0xffffffffu, d == dl);
                    if (d == dl) peers = grp;
                    unresolved &= ~grp;
                }
            }
            const uint32_t before = word & 0xffffu;
            if ((peers & lt) == 0) reinterpret_cast<volatile uint16_t *>(&wh[d])[0] = (uint16_t)(before + __popc(peers));
            __syncwarp();
            rank[i] = before + __popc(peers & lt);
        }
        __syncthreads();
        // ---- per-digit exclusive scan over warps; tile digit totals (thread d owns digit d)
        // (folding the run start into warp_hist here, to save one gather in the placement below, measured slower)
        {
            const int d = threadIdx.x;
            uint32_t tot = 0;
            if (d < radix) {
#pragma unroll
                for (int w = 0; w < RS_WARPS; ++w) {
                    uint32_t c = s.warp_hist[w][d] & 0xffffu;  // count; the top byte is the ranking's lane tag
                    s.warp_hist[w][d] = tot;
                    tot += c;
                }
            }
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s.warp_tot[wid] = incl;
            __syncthreads();
            uint32_t add = 0;
            for (int w = 0; w < wid; ++w) add += s.warp_tot[w];
            const uint32_t start = incl - tot + add;  // position of the digit's run in tile order
            s.tile_start[d] = start;
            unsigned long long real = tot;
            if (d == radix - 1) real -= (unsigned long long)(RS_TILE - valid);  // exclude padding
            const unsigned long long r = s.run[d];
            s.gbase[d] = r - start;
            s.run[d] = r + real;
        }
        __syncthreads();
        // ---- tile-order placement in shared memory
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            uint32_t d = (key[i] >> shift) & dmask;
            s.keys[s.tile_start[d] + s.warp_hist[wid][d] + rank[i]] = key[i];
        }
        __syncthreads();
        // ---- coalesced scatter
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            int p = threadIdx.x + j * RS_THREADS;
            if (p < valid) {
                uint32_t k = s.keys[p];
                uint32_t d = (k >> shift) & dmask;
                uint32_t v = OUT_FLOAT ? rs_unkey(k, km) : k;
                out[s.gbase[d] + p] = v;
            }
        }
    }
}

template <bool A, bool B>
static void launch_scatter(nnc_ctx *ctx, int chunks, const uint32_t *in, uint32_t *out, int64_t n, int64_t tiles_per_chunk,
                           int shift, int width, RsKeyMap km, const unsigned long long *chunk_base) {
    func_dyn_smem(ctx, (const void *)rs_scatter_kernel<A, B>, sizeof(RsSmem));
    NNC_LAUNCH(ctx, (rs_scatter_kernel<A, B>), chunks, RS_THREADS, sizeof(RsSmem), in, out, n, tiles_per_chunk, shift, width, km,
               chunk_base);
}

// Sorts the n non-zero floats in d_a ascending, using d_b as the other half of the ping-pong.  amin / amax: smallest
// and largest |x| bit pattern present (any enclosing range is valid).  Returns the buffer holding the result.
float *radix_sort_f32(nnc_ctx *ctx, float *d_a, float *d_b, int64_t n, uint32_t amin, uint32_t amax) {
    if (n <= 1) return d_a;
    if (n >= (1ll << 40)) NNC_FAIL(NNC_ERR_UNSUPPORTED, "radix sort: n too large");
    if (amax < amin) NNC_FAIL(NNC_ERR_INTERNAL, "radix sort: empty key range");
    RsKeyMap km{amin, amax - amin};
    const unsigned long long span = 2ull * km.range + 2ull;  // number of distinct keys
    int keybits = 1;
    while ((1ull << keybits) < span) keybits++;
    RsPlan plan;
    plan.passes = (keybits + 8) / 9;
    for (int p = 0, at = 0; p < RS_MAX_PASSES; ++p) {
        int left = plan.passes - p;
        // narrower digits first: the most significant digit of a bell-shaped layer is the skewed one, and a wider
        // digit there means fewer lanes of a warp item sharing a bin
        int wd = p < plan.passes ? (keybits - at) / left : 0;
        plan.shift[p] = at;
        plan.width[p] = wd;
        at += wd;
    }
    const int64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    const int64_t want_chunks = std::min<int64_t>(RS_MAX_CHUNKS, (int64_t)ctx->sm_count * 2);
    const int64_t tiles_per_chunk = (n_tiles + want_chunks - 1) / want_chunks;
    const int chunks = (int)((n_tiles + tiles_per_chunk - 1) / tiles_per_chunk);
    uint32_t *chunk_hist = arena_alloc_t<uint32_t>(ctx, (size_t)chunks * RS_RADIX);
    unsigned long long *chunk_base = arena_alloc_t<unsigned long long>(ctx, (size_t)chunks * RS_RADIX);
    uint32_t *a = reinterpret_cast<uint32_t *>(d_a), *b = reinterpret_cast<uint32_t *>(d_b);
    uint32_t *src = a, *dst = b;
    for (int p = 0; p < plan.passes; ++p) {
        const bool first = p == 0, last = p == plan.passes - 1;
        const int sh = plan.shift[p], wd = plan.width[p];
        const int vec_ok = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
        if (first)
            NNC_LAUNCH(ctx, rs_count_kernel<true>, chunks, RS_THREADS, 0, src, n, vec_ok, tiles_per_chunk, sh, wd, km, chunk_hist);
        else
            NNC_LAUNCH(ctx, rs_count_kernel<false>, chunks, RS_THREADS, 0, src, n, vec_ok, tiles_per_chunk, sh, wd, km, chunk_hist);
        NNC_LAUNCH(ctx, rs_base_kernel, 1, RS_RADIX, 0, chunk_hist, chunks, chunk_base);
        if (first && last)
            launch_scatter<true, true>(ctx, chunks, src, dst, n, tiles_per_chunk, sh, wd, km, chunk_base);
        else if (first)
            launch_scatter<true, false>(ctx, chunks, src, dst, n, tiles_per_chunk, sh, wd, km, chunk_base);
        else if (last)
            launch_scatter<false, true>(ctx, chunks, src, dst, n, tiles_per_chunk, sh, wd, km, chunk_base);
        else
            launch_scatter<false, false>(ctx, chunks, src, dst, n, tiles_per_chunk, sh, wd, km, chunk_base);
        std::swap(src, dst);
    }
    return reinterpret_cast<float *>(src);
}

}  // namespace nnc
