// khist.cu -- the sorted survivors as (distinct value, multiplicity) runs, built WITHOUT a comparison or LSD sort.
//
// The Lloyd kernels (lloyd.cu) only need the surviving weights as an ascending sequence with multiplicities.  For a
// large pruned layer the survivors occupy a narrow |x| bit range -- 26 bits of rank-image key for a pruned
// N(0, sigma^2) layer (sort.cu: RsKeyMap) -- while there are hundreds of millions of them, so the multiset is
// smaller as a full-resolution key histogram than as a sorted array.  This file builds that histogram:
//
//   count     per-chunk histograms of the key's HIGH digit (key >> 15)                       read  4 B/key
//   scatter   NON-stable partition by the high digit (shared-memory atomics rank the keys;
//             the tile is staged in digit order and written out in coalesced runs)           read + write 4 B/key
//   hist      one CTA per bucket: 2^15-bin shared-memory histogram of the LOW 15 bits        read  4 B/key
//             -> dense H[2^keybits] (uint32)                                                 write 4 B/bin
//   compact   non-empty bins -> (value, count) entries in key order = ascending value        read H, write 8 B/entry
//
// 16 B of traffic per key + 12..16 B per histogram bin, against 36 B per key (+ 4 for the tile sums) of the
// three-pass radix sort -- and none of it needs a stable ranking, which is what made the LSD scatter pass
// shared-memory bound.  Keys-only data: equal keys are indistinguishable, so the result is exactly the sorted array
// run-length encoded.  The path is taken when the histogram is not larger than the data (see hist_sort_applicable);
// small or wide-range tensors keep the radix sort.
//
// Reference context: this replaces, together with lloyd.cu, the n x k distance evaluations of
// sklearn/cluster/_k_means_lloyd.pyx:196-213 that neural_network_compression/common/utility.py:237-238 runs.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace nnc {

constexpr int KH_LOW = 15;                // low key bits resolved in shared memory
constexpr int KH_BINS = 1 << KH_LOW;      // 32768 bins = 128 KB of shared memory
constexpr int KH_MAX_HB = 12;             // high digit: at most 4096 buckets (keybits <= 27)
constexpr int KH_THREADS = 512;
constexpr int KH_ITEMS = 16;
constexpr int KH_TILE = KH_THREADS * KH_ITEMS;  // 8192 keys per scatter tile
constexpr int KH_CHUNK = 1 << 20;               // keys per work item of the histogram kernel (a pruned Gaussian layer at 2^30: ~260 K keys per bucket, one item each)
constexpr int KH_HT = 1024;                     // bins per compaction tile

struct KhKeyMap {  // same rank image as sort.cu's RsKeyMap
    uint32_t amin, range;
};
__device__ __forceinline__ uint32_t kh_key(uint32_t bits, KhKeyMap km) {
    const uint32_t m = (bits & 0x7fffffffu) - km.amin;
    return (bits & 0x80000000u) ? km.range - m : km.range + 1u + m;
}
__device__ __forceinline__ uint32_t kh_unkey(uint32_t key, KhKeyMap km) {
    return key <= km.range ? ((km.range - key + km.amin) | 0x80000000u) : (key - km.range - 1u + km.amin);
}

// ---- count: chunk_hist[c][d] = keys of chunk c whose high digit is d ---------------------------------------
__global__ void __launch_bounds__(KH_THREADS) kh_count_kernel(const uint32_t *__restrict__ in, int64_t n, int64_t tiles_per_chunk,
                                                              int nb, KhKeyMap km, uint32_t *__restrict__ chunk_hist) {
    extern __shared__ uint32_t kh_h[];
    for (int i = threadIdx.x; i < nb; i += KH_THREADS) kh_h[i] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * tiles_per_chunk * KH_TILE;
    const int64_t hi = min(n, lo + tiles_per_chunk * KH_TILE);
    auto one = [&](uint32_t bits) { atomicAdd(&kh_h[kh_key(bits, km) >> KH_LOW], 1u); };
    if ((reinterpret_cast<uintptr_t>(in) & 15u) == 0) {  // lo is a multiple of 8192 keys
        const int64_t nvec = (hi - lo) >> 2;
        const uint32_t *p = in + lo;
        for (int64_t i = threadIdx.x; i < nvec; i += KH_THREADS) {
            uint4 v = ld_stream_u4(p + 4 * i);
            one(v.x);
            one(v.y);
            one(v.z);
            one(v.w);
        }
        for (int64_t i = lo + (nvec << 2) + threadIdx.x; i < hi; i += KH_THREADS) one(in[i]);
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += KH_THREADS) one(in[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += KH_THREADS) chunk_hist[(size_t)blockIdx.x * nb + i] = kh_h[i];
}

// ---- base: bucket starts, per-chunk write cursors, work items of the histogram kernel -----------------------
// one thread per digit: crel[c][d] = keys of digit d in the chunks before c, tot[d] = keys of digit d
__global__ void __launch_bounds__(128) kh_crel_kernel(const uint32_t *__restrict__ chunk_hist, int chunks, int nb,
                                                      uint32_t *__restrict__ crel_lo, uint32_t *__restrict__ crel_hi,
                                                      unsigned long long *__restrict__ tot) {
    const int d = blockIdx.x * 128 + threadIdx.x;
    if (d >= nb) return;
    unsigned long long r = 0;
#pragma unroll 8
    for (int c = 0; c < chunks; ++c) {
        const uint32_t v = chunk_hist[(size_t)c * nb + d];
        crel_lo[(size_t)c * nb + d] = (uint32_t)r;
        crel_hi[(size_t)c * nb + d] = (uint32_t)(r >> 32);
        r += v;
    }
    tot[d] = r;
}

// bucket_off[d] = keys with a smaller high digit (bucket_off[nb] = n); item_off[d] = work items (KH_CHUNK keys each)
// of the buckets before d
__global__ void __launch_bounds__(1024) kh_base_kernel(const unsigned long long *__restrict__ tot, int nb,
                                                       unsigned long long *__restrict__ bucket_off, uint32_t *__restrict__ item_off) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned int s_iwarp[32];
    const int per = (nb + 1023) / 1024;  // digits per thread (consecutive)
    const int d0 = threadIdx.x * per;
    unsigned long long sum = 0;
    unsigned int items = 0;
    for (int j = 0; j < per; ++j) {
        const int d = d0 + j;
        const unsigned long long t = d < nb ? tot[d] : 0ull;
        sum += t;
        items += (unsigned int)((t + KH_CHUNK - 1) / KH_CHUNK);
    }
    const int lane = lane_id(), w = warp_id();
    unsigned long long incl = sum;
    unsigned int iincl = items;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        unsigned int ti = __shfl_up_sync(0xffffffffu, iincl, o);
        if (lane >= o) {
            incl += t;
            iincl += ti;
        }
    }
    if (lane == 31) {
        s_warp[w] = incl;
        s_iwarp[w] = iincl;
    }
    __syncthreads();
    unsigned long long add = 0;
    unsigned int iadd = 0;
    for (int i = 0; i < w; ++i) {
        add += s_warp[i];
        iadd += s_iwarp[i];
    }
    unsigned long long run = incl - sum + add;
    unsigned int irun = iincl - items + iadd;
    for (int j = 0; j < per; ++j) {
        const int d = d0 + j;
        if (d < nb) {
            bucket_off[d] = run;
            item_off[d] = irun;
            const unsigned long long t = tot[d];
            run += t;
            irun += (unsigned int)((t + KH_CHUNK - 1) / KH_CHUNK);
        }
    }
    if (threadIdx.x == 1023) {
        bucket_off[nb] = run;
        item_off[nb] = irun;
    }
}

// ---- scatter: non-stable partition by the high digit --------------------------------------------------------
// Output: the LOW KH_LOW bits of every key as uint16 (inside a bucket the high digit is the bucket itself): half the
// write traffic of a 32-bit key, and half the read of the histogram kernel.
// Shared memory (dynamic): keys[KH_TILE] | hist[nb] | gbase[nb] (u64) | run[nb] (u64) -- 72 KB at 2048 buckets, three
// CTAs per SM; the tile goes from global memory straight to registers (four 128-bit loads per thread).
template <int MIN_CTAS, int THREADS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) kh_scatter_kernel(const uint32_t *__restrict__ in, uint16_t *__restrict__ out, int64_t n,
                                                                int64_t tiles_per_chunk, int nb, KhKeyMap km,
                                                                const unsigned long long *__restrict__ bucket_off,
                                                                const uint32_t *__restrict__ crel_lo, const uint32_t *__restrict__ crel_hi) {
    extern __shared__ __align__(16) unsigned char kh_smem[];
    uint32_t *keys = reinterpret_cast<uint32_t *>(kh_smem);
    unsigned long long *gbase = reinterpret_cast<unsigned long long *>(keys + KH_TILE);
    unsigned long long *run = gbase + nb;
    uint32_t *hist = reinterpret_cast<uint32_t *>(run + nb);
    __shared__ uint32_t s_warp[32];
    constexpr int ITEMS = KH_TILE / THREADS;
    const int lane = lane_id(), wid = warp_id();
    for (int d = threadIdx.x; d < nb; d += THREADS)
        run[d] = bucket_off[d] + (((unsigned long long)crel_hi[(size_t)blockIdx.x * nb + d] << 32) | crel_lo[(size_t)blockIdx.x * nb + d]);
    const int64_t n_tiles = (n + KH_TILE - 1) / KH_TILE;
    const int64_t t0 = (int64_t)blockIdx.x * tiles_per_chunk, t1 = min(n_tiles, t0 + tiles_per_chunk);
    const bool src_aligned = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
    const int per = (nb + THREADS - 1) / THREADS;  // bins per thread in the scan (consecutive bins)
    // registers: 16 keys + 8 words of packed 13-bit ranks (three CTAs of 512 threads leave 42 registers per thread).
    // The RAW words of the next tile are loaded into the key registers as soon as the current tile's keys are staged in
    // shared memory: the load latency hides behind the write-out instead of standing at the top of every tile.
    uint32_t key[ITEMS];
    auto load_raw = [&](int64_t tile) {
        const int64_t tile_base = tile * KH_TILE;
        const int valid = (int)min((int64_t)KH_TILE, n - tile_base);
        if (src_aligned && valid == KH_TILE) {
#pragma unroll
            for (int c = 0; c < ITEMS / 4; ++c) {
                const uint4 v = ld_stream_u4(in + tile_base + 4 * (c * THREADS + threadIdx.x));
                key[4 * c] = v.x, key[4 * c + 1] = v.y, key[4 * c + 2] = v.z, key[4 * c + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int e = i * THREADS + threadIdx.x;
                key[i] = e < valid ? in[tile_base + e] : 0u;
            }
        }
    };
    if (t0 < t1) load_raw(t0);
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int64_t tile_base = tile * KH_TILE;
        const int valid = (int)min((int64_t)KH_TILE, n - tile_base);
        // raw word -> key; a key beyond the end of the array is the sentinel ~0 (a real key has at most KH_LOW + KH_MAX_HB bits)
        if (src_aligned && valid == KH_TILE) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = kh_key(key[i], km);
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int e = i * THREADS + threadIdx.x;
                key[i] = e < valid ? kh_key(key[i], km) : 0xffffffffu;
            }
        }
        for (int d = threadIdx.x; d < nb; d += THREADS) hist[d] = 0;
        __syncthreads();  // counters cleared, previous tile's write-out finished
        uint32_t rank2[ITEMS / 2];  // two ranks (< KH_TILE = 2^13) per word
#pragma unroll
        for (int i = 0; i < ITEMS; i += 2) {
            const uint32_t r0 = key[i] != 0xffffffffu ? atomicAdd(&hist[key[i] >> KH_LOW], 1u) : 0u;
            const uint32_t r1 = key[i + 1] != 0xffffffffu ? atomicAdd(&hist[key[i + 1] >> KH_LOW], 1u) : 0u;
            rank2[i / 2] = r0 | (r1 << 16);
        }
        __syncthreads();  // counts complete
        // exclusive scan of the digit counts (thread t owns bins [t * per, t * per + per))
        if (per == 4) {  // 2048 buckets: 128-bit shared-memory accesses (consecutive 16-byte pieces: conflict free)
            const int b0 = threadIdx.x * 4;
            uint4 c = *reinterpret_cast<const uint4 *>(&hist[b0]);
            const uint32_t loc = c.x + c.y + c.z + c.w;
            uint32_t incl = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_warp[wid] = incl;
            ulonglong2 r01 = *reinterpret_cast<const ulonglong2 *>(&run[b0]);
            ulonglong2 r23 = *reinterpret_cast<const ulonglong2 *>(&run[b0 + 2]);
            __syncthreads();
            uint32_t add = 0;
            for (int w2 = 0; w2 < wid; ++w2) add += s_warp[w2];
            const uint32_t s0 = incl - loc + add, s1 = s0 + c.x, s2 = s1 + c.y, s3 = s2 + c.z;
            *reinterpret_cast<uint4 *>(&hist[b0]) = make_uint4(s0, s1, s2, s3);
            *reinterpret_cast<ulonglong2 *>(&gbase[b0]) = make_ulonglong2(r01.x - s0, r01.y - s1);
            *reinterpret_cast<ulonglong2 *>(&gbase[b0 + 2]) = make_ulonglong2(r23.x - s2, r23.y - s3);
            *reinterpret_cast<ulonglong2 *>(&run[b0]) = make_ulonglong2(r01.x + c.x, r01.y + c.y);
            *reinterpret_cast<ulonglong2 *>(&run[b0 + 2]) = make_ulonglong2(r23.x + c.z, r23.y + c.w);
        } else {
            const int b0 = threadIdx.x * per;
            uint32_t loc = 0;
            for (int j = 0; j < per; ++j)
                if (b0 + j < nb) loc += hist[b0 + j];
            uint32_t incl = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_warp[wid] = incl;
            __syncthreads();
            uint32_t add = 0;
            for (int w2 = 0; w2 < wid; ++w2) add += s_warp[w2];
            uint32_t start = incl - loc + add;
            for (int j = 0; j < per; ++j) {
                const int d = b0 + j;
                if (d < nb) {
                    const uint32_t c = hist[d];
                    hist[d] = start;  // position of the digit's run in the staged tile
                    const unsigned long long r = run[d];
                    gbase[d] = r - start;
                    run[d] = r + c;
                    start += c;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (key[i] != 0xffffffffu) keys[hist[key[i] >> KH_LOW] + ((rank2[i / 2] >> (16 * (i & 1))) & 0xffffu)] = key[i];
        __syncthreads();
        if (tile + 1 < t1) load_raw(tile + 1);  // (the key registers are dead: the tile is in shared memory)
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int p = threadIdx.x + j * THREADS;
            if (p < valid) {
                const uint32_t k = keys[p];
                out[gbase[k >> KH_LOW] + p] = (uint16_t)(k & (KH_BINS - 1));
            }
        }
    }
}

// ---- scatter with software write combining (opt-in: NNC_SCATTER_WC=1; measured slower, see hist_sort_f32) ------
// The partition above writes, per tile and bucket, a run of ~4 keys (8 bytes): every 32-byte sector of the output is
// written in several pieces, tiles apart, and each piece costs a read-fill of the sector from DRAM (measured: 1.4 x the
// algorithmic traffic and 22 % of the HBM rate).  Here every CTA keeps a 16-key buffer per bucket in shared memory
// (position-major: wc[slot][bucket]) and writes a bucket's keys only as complete, 32-byte ALIGNED sectors; what is left
// in the buffers goes out once, at the end of the CTA's chunk (and a partial first sector once per bucket and chunk,
// because a chunk's part of a bucket starts anywhere).
// Shared memory (dynamic): stage[KH_TILE] u16 | wc[16][nb] u16 | hist[nb] u32 | wpos[nb] u64 | fill[nb] u8 -- 106 KB at
// 2048 buckets: two CTAs per SM.  nb <= KH_WC_MAX_NB; `out` 32-byte aligned.
constexpr int KH_WC = 16;
constexpr int KH_WC_MAX_NB = 2048;
__host__ __device__ inline size_t kh_wc_smem(int nb) { return (size_t)KH_TILE * 2 + (size_t)KH_WC * nb * 2 + (size_t)nb * 4 + (size_t)nb * 8 + (size_t)nb; }

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) kh_scatter_wc_kernel(const uint32_t *__restrict__ in, uint16_t *__restrict__ out, int64_t n,
                                                                    int64_t tiles_per_chunk, int nb, KhKeyMap km,
                                                                    const unsigned long long *__restrict__ bucket_off,
                                                                    const uint32_t *__restrict__ crel_lo, const uint32_t *__restrict__ crel_hi) {
    extern __shared__ __align__(16) unsigned char kh_smem[];
    unsigned long long *wpos = reinterpret_cast<unsigned long long *>(kh_smem);          // first unwritten position of the bucket
    uint32_t *hist = reinterpret_cast<uint32_t *>(wpos + nb);                            // count, then start in the stage
    uint16_t *stage = reinterpret_cast<uint16_t *>(hist + nb);                           // the tile's low digits in bucket order
    uint16_t *wc = stage + KH_TILE;                                                      // [KH_WC][nb]
    uint8_t *fill = reinterpret_cast<uint8_t *>(wc + (size_t)KH_WC * nb);                // keys waiting in the bucket's buffer
    __shared__ uint32_t s_warp[32];
    constexpr int ITEMS = KH_TILE / THREADS;
    constexpr int PER = KH_WC_MAX_NB / THREADS;  // consecutive buckets per thread (scan and write-out)
    const int lane = lane_id(), wid = warp_id();
    for (int d = threadIdx.x; d < nb; d += THREADS) {
        wpos[d] = bucket_off[d] + (((unsigned long long)crel_hi[(size_t)blockIdx.x * nb + d] << 32) | crel_lo[(size_t)blockIdx.x * nb + d]);
        fill[d] = 0;
    }
    const int64_t n_tiles = (n + KH_TILE - 1) / KH_TILE;
    const int64_t t0 = (int64_t)blockIdx.x * tiles_per_chunk, t1 = min(n_tiles, t0 + tiles_per_chunk);
    const bool src_aligned = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
    const int b0 = threadIdx.x * PER;
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int64_t tile_base = tile * KH_TILE;
        const int valid = (int)min((int64_t)KH_TILE, n - tile_base);
        uint32_t key[ITEMS];  // a key beyond the end of the array is the sentinel ~0
        if (src_aligned && valid == KH_TILE) {
#pragma unroll
            for (int c = 0; c < ITEMS / 4; ++c) {
                const uint4 v = ld_stream_u4(in + tile_base + 4 * (c * THREADS + threadIdx.x));
                key[4 * c] = kh_key(v.x, km), key[4 * c + 1] = kh_key(v.y, km), key[4 * c + 2] = kh_key(v.z, km), key[4 * c + 3] = kh_key(v.w, km);
            }
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int e = i * THREADS + threadIdx.x;
                key[i] = e < valid ? kh_key(in[tile_base + e], km) : 0xffffffffu;
            }
        }
        for (int d = threadIdx.x; d < nb; d += THREADS) hist[d] = 0;
        __syncthreads();  // counters cleared; the previous tile's write-out has finished with the stage
        uint32_t rank2[ITEMS / 2];  // two ranks (< KH_TILE = 2^13) per word
#pragma unroll
        for (int i = 0; i < ITEMS; i += 2) {
            const uint32_t r0 = key[i] != 0xffffffffu ? atomicAdd(&hist[key[i] >> KH_LOW], 1u) : 0u;
            const uint32_t r1 = key[i + 1] != 0xffffffffu ? atomicAdd(&hist[key[i + 1] >> KH_LOW], 1u) : 0u;
            rank2[i / 2] = r0 | (r1 << 16);
        }
        __syncthreads();  // counts complete
        // exclusive scan of the bucket counts; the thread keeps the counts and starts of its PER buckets
        uint32_t cnt[PER], st[PER];
        uint32_t loc = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            cnt[j] = b0 + j < nb ? hist[b0 + j] : 0u;
            loc += cnt[j];
        }
        uint32_t incl = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t run0 = incl - loc;
        for (int w2 = 0; w2 < wid; ++w2) run0 += s_warp[w2];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            st[j] = run0;
            if (b0 + j < nb) hist[b0 + j] = run0;
            run0 += cnt[j];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (key[i] != 0xffffffffu)
                stage[hist[key[i] >> KH_LOW] + ((rank2[i / 2] >> (16 * (i & 1))) & 0xffffu)] = (uint16_t)(key[i] & (KH_BINS - 1));
        __syncthreads();
        // write-out: per bucket the pending keys are buffer ++ this tile's run; complete aligned sectors go out
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int b = b0 + j;
            if (b >= nb) break;
            const uint32_t c = cnt[j], s0 = st[j], f = fill[b];
            if (c == 0) continue;
            const unsigned long long wp = wpos[b];
            const uint32_t total = f + c;
            uint32_t v0 = 0, need = KH_WC - (uint32_t)(wp & (KH_WC - 1));
            auto get = [&](uint32_t v) -> uint32_t { return v < f ? wc[(size_t)v * nb + b] : stage[s0 + v - f]; };
            while (total - v0 >= need) {
                uint16_t *dst = out + wp + v0;
                if (need == KH_WC) {
                    uint32_t w8[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) w8[t] = get(v0 + 2 * t) | (get(v0 + 2 * t + 1) << 16);
                    reinterpret_cast<uint4 *>(dst)[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                    reinterpret_cast<uint4 *>(dst)[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
                } else {  // up to the first sector boundary of this chunk's part of the bucket (once per bucket and chunk)
                    for (uint32_t t = 0; t < need; ++t) dst[t] = (uint16_t)get(v0 + t);
                }
                v0 += need;
                need = KH_WC;
            }
            const uint32_t r = total - v0;
            if (v0 == 0) {  // nothing went out: append the run to the buffer
                for (uint32_t t = 0; t < c; ++t) wc[(size_t)(f + t) * nb + b] = stage[s0 + t];
            } else {  // (a flush always drains the old buffer: f < need) the rest of the run starts the buffer anew
                for (uint32_t t = 0; t < r; ++t) wc[(size_t)t * nb + b] = stage[s0 + (v0 - f) + t];
            }
            fill[b] = (uint8_t)r;
            wpos[b] = wp + v0;
        }
        // (the next tile's first barrier comes after its loads: nobody overwrites the stage or the counters before it)
    }
    __syncthreads();
    for (int d = threadIdx.x; d < nb; d += THREADS) {  // what is left in the buffers
        const uint32_t f = fill[d];
        const unsigned long long wp = wpos[d];
        for (uint32_t t = 0; t < f; ++t) out[wp + t] = wc[(size_t)t * nb + d];
    }
}

// ---- hist: low-digit histogram of one work item (<= KH_CHUNK keys of one bucket) in shared memory -----------
// IN_FLOAT: the input is the raw survivor array (no partition pass ran: keybits <= KH_LOW, one bucket); otherwise the
// partition pass's uint16 low digits.
template <bool IN_FLOAT>
__global__ void __launch_bounds__(1024) kh_hist_kernel(const void *__restrict__ in_raw, int nb, int low_bins, KhKeyMap km,
                                                       const unsigned long long *__restrict__ bucket_off,
                                                       const uint32_t *__restrict__ item_off, uint32_t *__restrict__ H,
                                                       uint32_t *__restrict__ tnz) {
    extern __shared__ uint32_t kh_bins[];
    __shared__ uint32_t s_tnz[KH_BINS / KH_HT];
    const uint32_t item = blockIdx.x;
    if (item >= item_off[nb]) return;
    int lo = 0, hi = nb;  // largest bucket b with item_off[b] <= item (empty buckets own no items)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (item_off[mid] <= item)
            lo = mid;
        else
            hi = mid;
    }
    const int b = lo;
    const uint32_t items_b = item_off[b + 1] - item_off[b];
    const unsigned long long beg = bucket_off[b] + (unsigned long long)(item - item_off[b]) * KH_CHUNK;
    const unsigned long long end = min(bucket_off[b + 1], beg + KH_CHUNK);
    for (int i = threadIdx.x; i < low_bins; i += 1024) kh_bins[i] = 0;
    __syncthreads();
    const uint32_t lmask = (uint32_t)low_bins - 1u;
    const unsigned long long cnt = end - beg;
    if (IN_FLOAT) {
        auto one = [&](uint32_t v) { atomicAdd(&kh_bins[kh_key(v, km) & lmask], 1u); };
        // head up to the first 16-byte boundary, vector body, tail
        const uint32_t *p = reinterpret_cast<const uint32_t *>(in_raw) + beg;
        unsigned long long head = ((16u - (unsigned)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u) >> 2;
        if (head > cnt) head = cnt;
        if (threadIdx.x < head) one(p[threadIdx.x]);
        const unsigned long long nvec = (cnt - head) >> 2;
        const uint32_t *pv = p + head;
        for (unsigned long long i = threadIdx.x; i < nvec; i += 1024) {
            const uint4 v = ld_stream_u4(pv + 4 * i);
            one(v.x);
            one(v.y);
            one(v.z);
            one(v.w);
        }
        for (unsigned long long j = head + (nvec << 2) + threadIdx.x; j < cnt; j += 1024) one(p[j]);
    } else {
        auto one = [&](uint32_t v) { atomicAdd(&kh_bins[v & lmask], 1u); };
        auto two = [&](uint32_t w) {
            one(w & 0xffffu);
            one(w >> 16);
        };
        // eight 16-bit keys per 128-bit load: head up to the first 16-byte boundary, vector body, tail
        const uint16_t *p = reinterpret_cast<const uint16_t *>(in_raw) + beg;
        unsigned long long head = ((16u - (unsigned)(reinterpret_cast<uintptr_t>(p) & 15u)) & 15u) >> 1;
        if (head > cnt) head = cnt;
        if (threadIdx.x < head) one(p[threadIdx.x]);
        const unsigned long long nvec = (cnt - head) >> 3;
        const uint32_t *pv = reinterpret_cast<const uint32_t *>(p + head);
        unsigned long long i = threadIdx.x;
        for (; i + 1024 < nvec; i += 2 * 1024) {  // two 16-byte loads in flight per thread
            const uint4 v0 = ld_stream_u4(pv + 4 * i), v1 = ld_stream_u4(pv + 4 * (i + 1024));
            two(v0.x);
            two(v0.y);
            two(v0.z);
            two(v0.w);
            two(v1.x);
            two(v1.y);
            two(v1.z);
            two(v1.w);
        }
        for (; i < nvec; i += 1024) {
            const uint4 v = ld_stream_u4(pv + 4 * i);
            two(v.x);
            two(v.y);
            two(v.z);
            two(v.w);
        }
        for (unsigned long long j = head + (nvec << 3) + threadIdx.x; j < cnt; j += 1024) one(p[j]);
    }
    __syncthreads();
    uint32_t *Hb = H + ((size_t)b << KH_LOW);
    if (items_b == 1) {
        // the bucket is complete: its bins go out together with the non-empty count of every 1024-bin compaction tile
        const int tiles_b = max(low_bins, KH_HT) / KH_HT;
        if (threadIdx.x < KH_BINS / KH_HT) s_tnz[threadIdx.x] = 0;
        __syncthreads();
        for (int j = 0; j < tiles_b; ++j) {
            const int i = j * KH_HT + threadIdx.x;
            const uint32_t c = i < low_bins ? kh_bins[i] : 0u;
            if (i < low_bins) Hb[i] = c;  // every bin, zeros included: the bucket needs no clearing beforehand
            const int nzw = __popc(__ballot_sync(0xffffffffu, c != 0));
            if (lane_id() == 0 && nzw) atomicAdd(&s_tnz[j], (uint32_t)nzw);
        }
        __syncthreads();
        if (threadIdx.x < tiles_b) tnz[(size_t)b * (KH_BINS / KH_HT) + threadIdx.x] = s_tnz[threadIdx.x];
    } else {  // a bucket larger than one work item: merge the partial histograms
        for (int i = threadIdx.x; i < low_bins; i += 1024) {
            const uint32_t c = kh_bins[i];
            if (c) atomicAdd(&Hb[i], c);
        }
    }
}

// clears the bins of the buckets that are NOT written in full by one work item (empty ones, and the ones merged with
// atomics) -- a fraction of the 4 B/bin a memset of all of H would cost
__global__ void __launch_bounds__(1024) kh_zero_kernel(uint32_t *__restrict__ H, const uint32_t *__restrict__ item_off) {
    const int b = blockIdx.x;
    if (item_off[b + 1] - item_off[b] == 1) return;
    uint4 *p = reinterpret_cast<uint4 *>(H + ((size_t)b << KH_LOW));
    for (int i = threadIdx.x; i < KH_BINS / 4; i += 1024) p[i] = make_uint4(0, 0, 0, 0);
}

// ---- compact: non-empty bins -> (value, count) entries ---------------------------------------------------------
// buckets merged from several work items (atomics): their tile counts are taken once the bucket is complete
__global__ void __launch_bounds__(KH_HT) kh_tilecount_kernel(const uint32_t *__restrict__ H, const uint32_t *__restrict__ item_off,
                                                             int tiles_per_bucket, uint32_t *__restrict__ tnz) {
    __shared__ int s_warp[KH_HT / 32];
    const int b = blockIdx.x;
    if (item_off[b + 1] - item_off[b] <= 1) return;  // empty (counts stay 0) or counted by the histogram kernel
    for (int j = 0; j < tiles_per_bucket; ++j) {
        const size_t tile = (size_t)b * tiles_per_bucket + j;
        const uint32_t c = H[tile * KH_HT + threadIdx.x];
        const int nzw = __popc(__ballot_sync(0xffffffffu, c != 0));
        if (lane_id() == 0) s_warp[warp_id()] = nzw;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < KH_HT / 32; ++w) t += s_warp[w];
            tnz[tile] = (uint32_t)t;
        }
        __syncthreads();
    }
}

// 1024-bin tiles, four bins per thread; a CTA walks tiles with a grid stride (65 536 one-tile CTAs cost more in block
// scheduling than in memory traffic)
__global__ void __launch_bounds__(KH_HT / 4) kh_compact_kernel(const uint32_t *__restrict__ H, const uint32_t *__restrict__ tnz,
                                                               const unsigned long long *__restrict__ toff, long long n_htiles, KhKeyMap km,
                                                               float *__restrict__ val, uint32_t *__restrict__ cnt) {
    __shared__ int s_warp[2][KH_HT / 128];
    int par = 0;
    for (long long tile = blockIdx.x; tile < n_htiles; tile += gridDim.x) {
        if (tnz[tile] == 0) continue;  // uniform over the CTA
        const uint32_t key0 = (uint32_t)tile * KH_HT + threadIdx.x * 4;
        const uint4 c = ld_stream_u4(H + key0);
        const int mine = (c.x != 0) + (c.y != 0) + (c.z != 0) + (c.w != 0);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane_id() >= o) incl += t;
        }
        if (lane_id() == 31) s_warp[par][warp_id()] = incl;
        __syncthreads();
        int before = incl - mine;
        for (int w = 0; w < warp_id(); ++w) before += s_warp[par][w];
        par ^= 1;  // the next tile writes the other copy: no second barrier needed
        unsigned long long at = toff[tile] + (unsigned long long)before;
        if (c.x) {
            val[at] = __uint_as_float(kh_unkey(key0, km));
            cnt[at++] = c.x;
        }
        if (c.y) {
            val[at] = __uint_as_float(kh_unkey(key0 + 1, km));
            cnt[at++] = c.y;
        }
        if (c.z) {
            val[at] = __uint_as_float(kh_unkey(key0 + 2, km));
            cnt[at++] = c.z;
        }
        if (c.w) {
            val[at] = __uint_as_float(kh_unkey(key0 + 3, km));
            cnt[at++] = c.w;
        }
    }
}

// ---- host ----------------------------------------------------------------------------------------------------------
static int kh_keybits(uint32_t amin, uint32_t amax) {
    const unsigned long long span = 2ull * (unsigned long long)(amax - amin) + 2ull;  // number of distinct keys
    int keybits = 1;
    while ((1ull << keybits) < span) keybits++;
    return keybits;
}

bool hist_sort_applicable(int64_t n, uint32_t amin, uint32_t amax) {
    if (n <= 1 || amax < amin || n >= (1ll << 40)) return false;
    const int keybits = kh_keybits(amin, amax);
    if (keybits > KH_LOW + KH_MAX_HB) return false;
    const char *env = getenv("NNC_SORT_PATH");  // tests / experiments: "hist" or "radix"
    if (env && !strcmp(env, "hist")) return true;
    if (env && !strcmp(env, "radix")) return false;
    // the histogram costs about three sweeps of 4 B per BIN on top of 16 B per key; the radix sort 40 B per key
    return n >= (1ll << keybits) / 2;
}

SortedRuns hist_sort_f32(nnc_ctx *ctx, float *d_a, float *d_b, int64_t n, uint32_t amin, uint32_t amax) {
    const int keybits = kh_keybits(amin, amax);
    if (keybits > KH_LOW + KH_MAX_HB) NNC_FAIL(NNC_ERR_INTERNAL, "histogram path: %d key bits", keybits);
    KhKeyMap km{amin, amax - amin};
    const int hb = std::max(0, keybits - KH_LOW);
    const int nb = 1 << hb;
    const int low_bins = 1 << std::min(keybits, KH_LOW);
    const size_t n_bins = (size_t)1 << keybits;
    func_dyn_smem(ctx, (const void *)kh_hist_kernel<true>, KH_BINS * 4);
    func_dyn_smem(ctx, (const void *)kh_hist_kernel<false>, KH_BINS * 4);
    func_dyn_smem(ctx, (const void *)kh_scatter_kernel<2, 512>, KH_TILE * 4 + (1 << KH_MAX_HB) * 20);
    func_dyn_smem(ctx, (const void *)kh_scatter_kernel<3, 512>, KH_TILE * 4 + (1 << KH_MAX_HB) * 20);
    func_dyn_smem(ctx, (const void *)kh_scatter_kernel<2, 1024>, KH_TILE * 4 + (1 << KH_MAX_HB) * 20);
    uint32_t *H = arena_alloc_t<uint32_t>(ctx, std::max<size_t>(n_bins, KH_HT));
    if (hb == 0) NNC_CUDA(cudaMemsetAsync(H, 0, sizeof(uint32_t) * std::max<size_t>(n_bins, KH_HT), ctx->stream));
    unsigned long long *bucket_off = arena_alloc_t<unsigned long long>(ctx, (size_t)nb + 1);
    uint32_t *item_off = arena_alloc_t<uint32_t>(ctx, (size_t)nb + 1);
    const uint32_t *a = reinterpret_cast<const uint32_t *>(d_a);
    uint32_t *b = reinterpret_cast<uint32_t *>(d_b);
    uint16_t *b16 = reinterpret_cast<uint16_t *>(d_b);  // the partition pass's output: low digits
    const int max_items = (int)std::min<int64_t>(n / KH_CHUNK + nb, (int64_t)1 << 30);
    const long long n_htiles = (long long)(std::max<size_t>(n_bins, KH_HT) / KH_HT);
    const int tiles_per_bucket = (int)(n_htiles / nb);  // 32 with a partition pass, max(low_bins, 1024) / 1024 without
    uint32_t *tnz = arena_alloc_t<uint32_t>(ctx, (size_t)n_htiles + 1);
    NNC_CUDA(cudaMemsetAsync(tnz, 0, sizeof(uint32_t) * ((size_t)n_htiles + 1), ctx->stream));
    if (hb > 0) {
        const int64_t n_tiles = (n + KH_TILE - 1) / KH_TILE;
        const char *occ_env = getenv("NNC_SCATTER_CTAS");
        const int scatter_ctas = (occ_env && atoi(occ_env) == 3) ? 3 : 2;  // resident scatter CTAs per SM: one chunk each
        const int64_t want_chunks = std::min<int64_t>(1024, (int64_t)ctx->sm_count * scatter_ctas);
        const int64_t tiles_per_chunk = (n_tiles + want_chunks - 1) / want_chunks;
        const int chunks = (int)((n_tiles + tiles_per_chunk - 1) / tiles_per_chunk);
        uint32_t *chunk_hist = arena_alloc_t<uint32_t>(ctx, (size_t)chunks * nb);
        uint32_t *crel_lo = arena_alloc_t<uint32_t>(ctx, (size_t)chunks * nb);
        uint32_t *crel_hi = arena_alloc_t<uint32_t>(ctx, (size_t)chunks * nb);
        unsigned long long *tot = arena_alloc_t<unsigned long long>(ctx, (size_t)nb);
        NNC_LAUNCH(ctx, kh_count_kernel, chunks, KH_THREADS, nb * 4, a, n, tiles_per_chunk, nb, km, chunk_hist);
        NNC_LAUNCH(ctx, kh_crel_kernel, (nb + 127) / 128, 128, 0, chunk_hist, chunks, nb, crel_lo, crel_hi, tot);
        NNC_LAUNCH(ctx, kh_base_kernel, 1, 1024, 0, tot, nb, bucket_off, item_off);
        // two CTAs per SM (64 registers) by default; three (42 registers: spills) measured slower, NNC_SCATTER_CTAS=3 selects it
        const char *thr_env = getenv("NNC_SCATTER_THREADS");
        // NNC_SCATTER_WC=1: complete 32-byte sectors out of per-bucket buffers in shared memory.  Measured and rejected: it
        // removes the partial-sector read-fills (the 1.4 x DRAM traffic of the plain partition), but the partition is bound
        // by its barriers and dependent shared-memory phases, not by DRAM, and the sector assembly adds to them
        // (1.62 ms against 1.40 ms at 2^30 weights)
        const bool wc_ok = nb <= KH_WC_MAX_NB && (reinterpret_cast<uintptr_t>(b16) & 31u) == 0 && getenv("NNC_SCATTER_WC");
        if (wc_ok) {
            func_dyn_smem(ctx, (const void *)kh_scatter_wc_kernel<512>, kh_wc_smem(KH_WC_MAX_NB));
            NNC_LAUNCH_AS(ctx, "kh_scatter_kernel", (kh_scatter_wc_kernel<512>), chunks, 512, kh_wc_smem(nb), a, b16, n, tiles_per_chunk, nb, km,
                          bucket_off, crel_lo, crel_hi);
        } else if (thr_env && atoi(thr_env) == 1024)  // 2 CTAs x 1024 threads x 8 keys: 32 registers, full occupancy
            NNC_LAUNCH_AS(ctx, "kh_scatter_kernel", (kh_scatter_kernel<2, 1024>), chunks, 1024, KH_TILE * 4 + nb * 20, a, b16, n, tiles_per_chunk,
                          nb, km, bucket_off, crel_lo, crel_hi);
        else if (scatter_ctas == 2)
            NNC_LAUNCH_AS(ctx, "kh_scatter_kernel", (kh_scatter_kernel<2, 512>), chunks, 512, KH_TILE * 4 + nb * 20, a, b16, n, tiles_per_chunk, nb,
                          km, bucket_off, crel_lo, crel_hi);
        else
            NNC_LAUNCH_AS(ctx, "kh_scatter_kernel", (kh_scatter_kernel<3, 512>), chunks, 512, KH_TILE * 4 + nb * 20, a, b16, n, tiles_per_chunk, nb,
                          km, bucket_off, crel_lo, crel_hi);
        NNC_LAUNCH(ctx, kh_zero_kernel, nb, 1024, 0, H, item_off);
        NNC_LAUNCH(ctx, kh_hist_kernel<false>, max_items, 1024, low_bins * 4, (const void *)b16, nb, low_bins, km, bucket_off, item_off, H, tnz);
    } else {
        const unsigned long long h_off[2] = {0ull, (unsigned long long)n};
        const uint32_t h_items[2] = {0u, (uint32_t)((n + KH_CHUNK - 1) / KH_CHUNK)};
        NNC_CUDA(cudaMemcpyAsync(bucket_off, h_off, sizeof(h_off), cudaMemcpyHostToDevice, ctx->stream));
        NNC_CUDA(cudaMemcpyAsync(item_off, h_items, sizeof(h_items), cudaMemcpyHostToDevice, ctx->stream));
        NNC_CUDA(cudaStreamSynchronize(ctx->stream));  // the two host arrays are on this frame
        NNC_LAUNCH(ctx, kh_hist_kernel<true>, (int)h_items[1], 1024, low_bins * 4, (const void *)a, 1, low_bins, km, bucket_off, item_off, H, tnz);
    }
    // non-empty bins -> entries; the survivor buffers are dead from here on and take the entries
    unsigned long long *toff = arena_alloc_t<unsigned long long>(ctx, (size_t)n_htiles + 2);
    NNC_LAUNCH(ctx, kh_tilecount_kernel, nb, KH_HT, 0, H, item_off, tiles_per_bucket, tnz);
    exclusive_scan_u32_u64(ctx, tnz, n_htiles, toff);
    NNC_LAUNCH(ctx, kh_compact_kernel, (int)std::min<long long>(n_htiles, (long long)ctx->sm_count * 32), KH_HT / 4, 0, H, tnz, toff, n_htiles, km, d_a, b);
    unsigned long long n_ent = 0;
    NNC_CUDA(cudaMemcpyAsync(&n_ent, toff + n_htiles, sizeof(n_ent), cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_ent == 0 || (int64_t)n_ent > n) NNC_FAIL(NNC_ERR_INTERNAL, "histogram path: %llu entries for %lld keys", n_ent, (long long)n);
    SortedRuns r;
    r.val = d_a;
    r.cnt = b;
    r.n_ent = (int64_t)n_ent;
    return r;
}

}  // namespace nnc
