// sort.cu -- hand-written "onesweep" least-significant-digit radix sort of float32 keys (keys only).
//
// The reference clusters with n x k distance evaluations per Lloyd iteration
// (sklearn/cluster/_k_means_lloyd.pyx:196-213).  For 1-D data every cluster is an interval of the sorted
// samples, so the B200 path sorts the surviving (non-zero) weights ONCE and turns each Lloyd iteration into
// boundary searches (lloyd.cu).  This file is that one-time sort.
//
// Algorithm (Adinets & Merrill, "Onesweep"): one upfront kernel builds the four 8-bit digit histograms; each
// of the four passes then reads a tile of keys once, ranks it in shared memory (warp-level match_any
// multisplit, stable), resolves the tile's global digit offsets with a decoupled look-back over per-tile
// digit counts, and scatters the tile-sorted keys.  Traffic: 4 B read for the histograms + 4 x (4 B read +
// 4 B write) per key.  Floats are mapped to order-preserving uint32 on the first read and back on the last
// write.
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

constexpr unsigned long long RS_VAL_MASK = (1ull << 54) - 1;
__device__ __forceinline__ unsigned long long rs_pack(unsigned flag, unsigned epoch, unsigned long long v) {
    return ((unsigned long long)flag << 62) | ((unsigned long long)(epoch & 0xffu) << 54) | (v & RS_VAL_MASK);
}

// ---- upfront histograms of all digits ----------------------------------------------------------------
__global__ void __launch_bounds__(512) rs_hist_kernel(const uint32_t *in, int64_t n, int vec_ok, RsKeyMap km, RsPlan plan,
                                                      unsigned long long *ghist /*[passes][RS_RADIX]*/) {
    __shared__ uint32_t h[RS_MAX_PASSES][RS_RADIX];
    for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_RADIX; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    auto one = [&](uint32_t bits) {
        const uint32_t k = rs_key(bits, km);
#pragma unroll
        for (int p = 0; p < RS_MAX_PASSES; ++p)
            if (p < plan.passes) atomicAdd(&h[p][(k >> plan.shift[p]) & ((1u << plan.width[p]) - 1u)], 1u);
    };
    int64_t nvec = vec_ok ? (n >> 2) : 0;
    // per-CTA counts stay below 2^32: each CTA sees at most n / gridDim + slack keys (n < 2^40 / grid)
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 v = ld_stream_u4(in + 4 * i);
        one(v.x);
        one(v.y);
        one(v.z);
        one(v.w);
    }
    for (int64_t i = (nvec << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        one(in[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * RS_RADIX; i += blockDim.x) {
        uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(&ghist[i], (unsigned long long)c);
    }
}

// exclusive scan of each digit histogram -> global base offset of every digit value
__global__ void __launch_bounds__(RS_RADIX) rs_scan_kernel(unsigned long long *ghist /*[passes][RS_RADIX] in place*/, int passes) {
    __shared__ unsigned long long s[RS_RADIX];
    for (int p = 0; p < passes; ++p) {
        s[threadIdx.x] = ghist[p * RS_RADIX + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long run = 0;
            for (int i = 0; i < RS_RADIX; ++i) {
                unsigned long long c = s[i];
                s[i] = run;
                run += c;
            }
        }
        __syncthreads();
        ghist[p * RS_RADIX + threadIdx.x] = s[threadIdx.x];
        __syncthreads();
    }
}

// ---- one onesweep pass -------------------------------------------------------------------------------
struct RsSmem {
    uint32_t keys[RS_TILE];
    uint32_t warp_hist[RS_WARPS][RS_RADIX];
    uint32_t tile_start[RS_RADIX];
    unsigned long long gbase[RS_RADIX];
    uint32_t warp_tot[RS_WARPS];
    uint32_t tile;
};

template <bool IN_FLOAT, bool OUT_FLOAT>
__global__ void __launch_bounds__(RS_THREADS) rs_pass_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                             int64_t n, int shift, int width, RsKeyMap km,
                                                             const unsigned long long *__restrict__ digit_base,
                                                             unsigned long long *state, unsigned int *ticket,
                                                             unsigned epoch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem &s = *reinterpret_cast<RsSmem *>(smem_raw);
    const int lane = lane_id(), wid = warp_id();
    const uint32_t dmask = (1u << width) - 1u;
    const int radix = 1 << width;

    if (threadIdx.x == 0) s.tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&s.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s.tile;
    const int64_t tile_base = (int64_t)tile * RS_TILE;
    const int valid = (int)min((int64_t)RS_TILE, n - tile_base);

    // ---- load (warp-striped: item i of lane l is element i*32 + l of the warp's 512-key chunk)
    uint32_t key[RS_ITEMS];
    const int64_t wbase = tile_base + wid * (32 * RS_ITEMS);
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        int64_t idx = wbase + i * 32 + lane;
        uint32_t k = 0xffffffffu;  // padding: all digit bits set, sorts last in every pass
        if (idx < n) {
            k = __ldg(in + idx);
            if (IN_FLOAT) k = rs_key(k, km);
        }
        key[i] = k;
    }
    // ---- stable ranking inside the warp: match_any multisplit with warp-private digit counters
    uint32_t rank[RS_ITEMS];
    uint32_t *wh = s.warp_hist[wid];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        uint32_t d = (key[i] >> shift) & dmask;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t before = wh[d];
        __syncwarp();
        if ((peers & lt) == 0) wh[d] = before + __popc(peers);  // lowest peer lane updates the counter
        __syncwarp();
        rank[i] = before + __popc(peers & lt);
    }
    __syncthreads();
    // ---- per-digit exclusive scan over warps; tile digit totals (thread d owns digit d)
    uint32_t tot = 0;
    {
        const int d = threadIdx.x;
        if (d < radix) {
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                uint32_t c = s.warp_hist[w][d];
                s.warp_hist[w][d] = tot;
                tot += c;
            }
        }
        // exclusive scan of the digit totals -> position of each digit's run in tile order
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s.warp_tot[wid] = incl;
        s.tile_start[d] = incl - tot;  // warp-local exclusive; fixed up after the barrier
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        uint32_t add = 0;
        for (int w = 0; w < wid; ++w) add += s.warp_tot[w];
        const uint32_t start = s.tile_start[d] + add;
        if (d < radix) {
            // ---- decoupled look-back for this digit
            unsigned long long real = tot;
            if (d == radix - 1) real -= (unsigned long long)(RS_TILE - valid);  // exclude padding
            unsigned long long excl = 0;
            unsigned long long *my = state + (size_t)tile * RS_RADIX + d;
            if (tile == 0) {
                st_volatile_u64(my, rs_pack(2u, epoch, real));
            } else {
                st_volatile_u64(my, rs_pack(1u, epoch, real));
                for (int64_t p = (int64_t)tile - 1; p >= 0; --p) {
                    const unsigned long long *q = state + (size_t)p * RS_RADIX + d;
                    unsigned long long v;
                    unsigned flag;
                    do {
                        v = ld_volatile_u64(q);
                        flag = (unsigned)(v >> 62);
                        if (((v >> 54) & 0xffu) != (epoch & 0xffu)) flag = 0;  // stale word from an earlier pass
                    } while (flag == 0);
                    excl += v & RS_VAL_MASK;
                    if (flag == 2u) break;
                }
                st_volatile_u64(my, rs_pack(2u, epoch, excl + real));
            }
            s.gbase[d] = digit_base[d] + excl - start;
        }
        __syncthreads();  // every thread has read the warp-local tile_start of its digit
        s.tile_start[d] = start;
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

template <bool A, bool B>
static void launch_pass(nnc_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n, int shift, int width, RsKeyMap km,
                        const unsigned long long *digit_base, unsigned long long *state, unsigned int *ticket,
                        unsigned epoch) {
    static bool configured = false;
    if (!configured) {
        NNC_CUDA(cudaFuncSetAttribute(rs_pass_kernel<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
        configured = true;
    }
    const int64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    NNC_LAUNCH(ctx, (rs_pass_kernel<A, B>), (unsigned)n_tiles, RS_THREADS, sizeof(RsSmem), in, out, n, shift, width, km,
               digit_base, state, ticket, epoch);
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
        int wd = p < plan.passes ? (keybits - at + left - 1) / left : 0;
        plan.shift[p] = at;
        plan.width[p] = wd;
        at += wd;
    }
    const int64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    unsigned long long *ghist = arena_alloc_t<unsigned long long>(ctx, RS_MAX_PASSES * RS_RADIX + 4);
    unsigned int *tickets = reinterpret_cast<unsigned int *>(ghist + RS_MAX_PASSES * RS_RADIX);
    unsigned long long *state = arena_alloc_t<unsigned long long>(ctx, (size_t)n_tiles * RS_RADIX);
    NNC_CUDA(cudaMemsetAsync(ghist, 0, sizeof(unsigned long long) * (RS_MAX_PASSES * RS_RADIX + 4), ctx->stream));
    NNC_CUDA(cudaMemsetAsync(state, 0, sizeof(unsigned long long) * (size_t)n_tiles * RS_RADIX, ctx->stream));
    uint32_t *a = reinterpret_cast<uint32_t *>(d_a), *b = reinterpret_cast<uint32_t *>(d_b);
    int vec_ok = (reinterpret_cast<uintptr_t>(d_a) & 15u) == 0;
    int hgrid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 4, (n / 4 + 511) / 512 + 1);
    NNC_LAUNCH(ctx, rs_hist_kernel, hgrid, 512, 0, a, n, vec_ok, km, plan, ghist);
    NNC_LAUNCH(ctx, rs_scan_kernel, 1, RS_RADIX, 0, ghist, plan.passes);
    uint32_t *src = a, *dst = b;
    for (int p = 0; p < plan.passes; ++p) {
        const bool first = p == 0, last = p == plan.passes - 1;
        const unsigned long long *base = ghist + (size_t)p * RS_RADIX;
        if (first && last)
            launch_pass<true, true>(ctx, src, dst, n, plan.shift[p], plan.width[p], km, base, state, tickets + p, p + 1);
        else if (first)
            launch_pass<true, false>(ctx, src, dst, n, plan.shift[p], plan.width[p], km, base, state, tickets + p, p + 1);
        else if (last)
            launch_pass<false, true>(ctx, src, dst, n, plan.shift[p], plan.width[p], km, base, state, tickets + p, p + 1);
        else
            launch_pass<false, false>(ctx, src, dst, n, plan.shift[p], plan.width[p], km, base, state, tickets + p, p + 1);
        std::swap(src, dst);
    }
    return reinterpret_cast<float *>(src);
}

}  // namespace nnc
