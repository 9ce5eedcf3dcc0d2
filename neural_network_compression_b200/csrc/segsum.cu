// segsum.cu -- trained-quantization gradient sum: out[j] = sum_i grad[i] * [code_i == j]
// (formula: papers/lat/report.tex:152; the reference declares it "not implemented", report.tex:154-158).
//
// One streaming pass, 4 B (gradient) + bits/8 B (code) per element.  Accumulation is exact inside a tile:
// every tile of 4096 gradients is scaled by a power of two derived from the tile's own max |g| and summed as
// 64-bit integers in shared-memory bins (integer adds commute, so the shared-memory atomics are order
// independent); the tile totals are then added in float64, in tile order, into per-CTA bins, and the per-CTA
// bins are folded in CTA order by a second tiny kernel.  The result is therefore deterministic for a given
// (n, grid) and within a few float64 ulps of the exactly rounded sum.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "table.cuh"

namespace nnc {

constexpr int SG_THREADS = 256;
constexpr int SG_VEC = 4;                            // float4 groups per thread and tile
constexpr int SG_TILE = SG_THREADS * SG_VEC * 4;     // 4096 elements

// codes of the 4 consecutive elements 4*g4 .. 4*g4+3
__device__ __forceinline__ void load_codes4(const void *codes, int bits, int64_t g4, int64_t n, int (&c)[4]) {
    if (bits == 0) {
        int4 v = *reinterpret_cast<const int4 *>(reinterpret_cast<const int32_t *>(codes) + 4 * g4);
        c[0] = v.x, c[1] = v.y, c[2] = v.z, c[3] = v.w;
    } else if (bits == 8) {
        uint32_t v = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(codes) + 4 * g4);
        c[0] = v & 255u, c[1] = (v >> 8) & 255u, c[2] = (v >> 16) & 255u, c[3] = v >> 24;
    } else if (bits == 4) {
        uint32_t v = *reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(codes) + 2 * g4);
        c[0] = v & 15u, c[1] = (v >> 4) & 15u, c[2] = (v >> 8) & 15u, c[3] = (v >> 12) & 15u;
    } else {
        const uint8_t *p = reinterpret_cast<const uint8_t *>(codes);
        const int64_t total_bytes = (n * bits + 7) / 8;
        const uint32_t mask = (1u << bits) - 1u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t bit = (4 * g4 + j) * bits, byte = bit >> 3;
            uint32_t v = 0;
            for (int b = 0; b < 3; ++b)
                if (byte + b < total_bytes) v |= (uint32_t)p[byte + b] << (8 * b);
            c[j] = (int)((v >> (bit & 7)) & mask);
        }
    }
}
__device__ __forceinline__ int load_code1(const void *codes, int bits, int64_t i, int64_t n) {
    if (bits == 0) return reinterpret_cast<const int32_t *>(codes)[i];
    const uint8_t *p = reinterpret_cast<const uint8_t *>(codes);
    const int64_t total_bytes = (n * bits + 7) / 8;
    const int64_t bit = i * bits, byte = bit >> 3;
    uint32_t v = 0;
    for (int b = 0; b < 3; ++b)
        if (byte + b < total_bytes) v |= (uint32_t)p[byte + b] << (8 * b);
    return (int)((v >> (bit & 7)) & ((1u << bits) - 1u));
}

// WARP_PRIV (k <= 1024): every warp owns a private copy of the bins and adds with NATIVE 32-bit shared-memory
// atomics -- the fixed-point value q (|q| < 2^43) is split as q = hi * 2^22 + lo, lo in [0, 2^22): a warp adds at most
// 512 values per tile, so neither half can overflow 32 bits.  (A 64-bit shared-memory atomicAdd compiles to a
// compare-and-swap spin loop, ATOMS.CAST.SPIN.64, which collapses when one code dominates.)  Otherwise: one set of
// 64-bit bins per CTA.
constexpr int SG_WARPS = SG_THREADS / 32;
constexpr int SG_QBITS_PRIV = 43, SG_SPLIT = 22;

template <bool WARP_PRIV>
__global__ void __launch_bounds__(SG_THREADS) segsum_kernel(const float *__restrict__ grad, const void *__restrict__ codes,
                                                            int64_t n, int bits, int k, int vec_ok, double *partial,
                                                            unsigned long long *bad_codes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_dbl = reinterpret_cast<double *>(smem_raw);                                // [k]
    long long *s_bin = reinterpret_cast<long long *>(s_dbl + k);                          // !WARP_PRIV: [k]
    uint32_t *s_lo = reinterpret_cast<uint32_t *>(s_dbl + k);                             // WARP_PRIV: [SG_WARPS][k]
    int32_t *s_hi = reinterpret_cast<int32_t *>(s_lo + (size_t)SG_WARPS * k);             // WARP_PRIV: [SG_WARPS][k]
    __shared__ float s_max[SG_THREADS / 32];
    __shared__ float s_tile_max;
    for (int i = threadIdx.x; i < k; i += SG_THREADS) s_dbl[i] = 0.0;
    if (WARP_PRIV) {
        for (int i = threadIdx.x; i < 2 * SG_WARPS * k; i += SG_THREADS) s_lo[i] = 0;
    } else {
        for (int i = threadIdx.x; i < k; i += SG_THREADS) s_bin[i] = 0;
    }
    uint32_t *my_lo = s_lo + (size_t)warp_id() * k;
    int32_t *my_hi = s_hi + (size_t)warp_id() * k;
    constexpr int QBITS = WARP_PRIV ? SG_QBITS_PRIV : 50;
    __syncthreads();
    const int64_t n_tiles = (n + SG_TILE - 1) / SG_TILE;
    unsigned long long bad = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t base = t * SG_TILE;
        float g[SG_VEC][4];
        int c[SG_VEC][4];
        float mx = 0.f;
#pragma unroll
        for (int v = 0; v < SG_VEC; ++v) {
            const int64_t e = base + ((int64_t)v * SG_THREADS + threadIdx.x) * 4;
            if (vec_ok && e + 4 <= n) {
                float4 x = ld_stream_f4(grad + e);
                g[v][0] = x.x, g[v][1] = x.y, g[v][2] = x.z, g[v][3] = x.w;
                load_codes4(codes, bits, e >> 2, n, c[v]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool in = e + j < n;
                    g[v][j] = in ? grad[e + j] : 0.f;
                    c[v][j] = in ? load_code1(codes, bits, e + j, n) : -1;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c[v][j] >= k) {
                    bad++;
                    c[v][j] = -1;
                }
                if (c[v][j] >= 0) mx = fmaxf(mx, fabsf(g[v][j]));
            }
        }
        mx = warp_max_f(mx);
        if (lane_id() == 0) s_max[warp_id()] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = s_max[0];
            for (int i = 1; i < SG_THREADS / 32; ++i) m = fmaxf(m, s_max[i]);
            s_tile_max = m;
        }
        __syncthreads();
        const float tmax = s_tile_max;
        // |g| < 2^E ; q = rint(g * 2^(QBITS-E)), |q| < 2^QBITS
        int E = 0;
        if (tmax > 0.f) E = (int)((__float_as_uint(tmax) >> 23) & 0xffu) - 126;  // ilogb + 1 (denormals: E = -126)
        if (tmax > 0.f && tmax < 1.17549435e-38f) E = -126;
        const double scale = ldexp(1.0, QBITS - E);
#pragma unroll
        for (int v = 0; v < SG_VEC; ++v) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int code = c[v][j];
                long long q = code >= 0 ? __double2ll_rn(__dmul_rn((double)g[v][j], scale)) : 0;
                // Shared-memory integer atomics commute, so any order gives the same bins.  A code that many lanes of
                // the warp share (the cluster of the pruned zeros holds 2/3 of a pruned layer) would serialise on one
                // address: up to two such groups are summed with shuffles first (leader = lowest remaining lane),
                // the rest goes through individual atomics.  (__match_any_sync costs ~60 cycles per warp instruction
                // on sm_100 when the 32 values are mostly distinct: scripts/micro/matchany.cu.)
                unsigned remaining = __ballot_sync(0xffffffffu, code >= 0);
                bool done = code < 0;
#pragma unroll
                for (int round = 0; round < 2; ++round) {
                    if (remaining == 0) break;
                    const int leader = __ffs(remaining) - 1;
                    const int Lc = __shfl_sync(0xffffffffu, code, leader);
                    const unsigned grp = __ballot_sync(0xffffffffu, !done && code == Lc);
                    if (__popc(grp) < 4) break;
                    if (WARP_PRIV) {
                        // the two 32-bit halves of the group go through the hardware warp reduction (REDUX): at most 32
                        // values of 22 / 21 bits each, no overflow; two instructions instead of ten 64-bit shuffle steps
                        if ((grp >> lane_id()) & 1u) {
                            const uint32_t slo = __reduce_add_sync(grp, (uint32_t)(q & ((1ll << SG_SPLIT) - 1)));
                            const int32_t shi = __reduce_add_sync(grp, (int32_t)(q >> SG_SPLIT));
                            if (lane_id() == leader) {
                                atomicAdd(&my_lo[Lc], slo);
                                atomicAdd(&my_hi[Lc], shi);
                            }
                        }
                    } else {
                        const long long sgrp = warp_sum_ll((grp >> lane_id()) & 1u ? q : 0);
                        if (lane_id() == leader) atomicAdd((unsigned long long *)&s_bin[Lc], (unsigned long long)sgrp);
                    }
                    if ((grp >> lane_id()) & 1u) done = true;
                    remaining &= ~grp;
                }
                if (!done) {
                    if (WARP_PRIV) {
                        atomicAdd(&my_lo[code], (uint32_t)(q & ((1ll << SG_SPLIT) - 1)));
                        atomicAdd(&my_hi[code], (int32_t)(q >> SG_SPLIT));
                    } else {
                        atomicAdd((unsigned long long *)&s_bin[code], (unsigned long long)q);
                    }
                }
            }
        }
        __syncthreads();
        const double inv = ldexp(1.0, E - QBITS);
        for (int i = threadIdx.x; i < k; i += SG_THREADS) {
            long long b = 0;
            if (WARP_PRIV) {
#pragma unroll
                for (int w = 0; w < SG_WARPS; ++w) {
                    b += ((long long)s_hi[(size_t)w * k + i] << SG_SPLIT) + (long long)s_lo[(size_t)w * k + i];
                    s_hi[(size_t)w * k + i] = 0;
                    s_lo[(size_t)w * k + i] = 0;
                }
            } else {
                b = s_bin[i];
                s_bin[i] = 0;
            }
            if (b) s_dbl[i] = __dadd_rn(s_dbl[i], __dmul_rn((double)b, inv));
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < k; i += SG_THREADS) partial[(size_t)blockIdx.x * k + i] = s_dbl[i];
    bad = warp_sum_ull(bad);
    if (lane_id() == 0 && bad) atomicAdd(bad_codes, bad);
}

// ---------------------------------------------------------------------------------------------------------------
// k <= 1024: one warp = one independent accumulator, no block barrier anywhere in the loop.
//   * a warp takes tiles of 512 consecutive elements (16 per lane: four 128-bit gradient loads + their codes in flight);
//   * fixed point WITHOUT float64: q = rint(g * 2^(43 - E)) is a power-of-two scaling, exact in float32, then one
//     F2I.S64; 2^E bounds every |g| the warp has seen so far (running exponent: when a tile exceeds it the warp's bins are
//     flushed and E grows -- a handful of times per launch);
//   * the warp's private bins are exact 64-bit accumulators made of two NATIVE 32-bit shared-memory atomics: the low
//     word wraps, its carry (old + x < old, from the returned value) rides on the add to the high word;
//   * the most frequent code of the warp's first tile (the cluster of the pruned zeros holds 2/3 of a pruned layer) is
//     summed in a register instead of serialising 20 lanes on one address;
//   * flushes add bins * 2^(E - 43) in float64 into the WARP's own row of `partial` (one writer per row: a fixed order),
//     and a second kernel folds the rows in row order: deterministic for a given (n, grid).
// Exactness: elements below 2^-20 of the running bound lose bits below 2^(E - 44); everything else is summed exactly.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SW_QBITS = 43;
constexpr int SW_TILE = 512;           // elements per warp tile
constexpr int SW_FLUSH_TILES = 1024;   // bins are flushed at least this often: |high word| < 2^11 * 2^19 + carries < 2^31

__global__ void __launch_bounds__(SG_THREADS) segsum_warp_kernel(const float *__restrict__ grad, const void *__restrict__ codes,
                                                                 int64_t n, int bits, int k, int vec_ok, double *partial,
                                                                 unsigned long long *bad_codes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id(), wid = warp_id();
    uint32_t *my_lo = reinterpret_cast<uint32_t *>(smem_raw) + (size_t)wid * 2 * k;
    int32_t *my_hi = reinterpret_cast<int32_t *>(my_lo + k);
    for (int i = lane; i < 2 * k; i += 32) my_lo[i] = 0;
    const int64_t gwarp = (int64_t)blockIdx.x * SG_WARPS + wid, n_warps = (int64_t)gridDim.x * SG_WARPS;
    double *my_row = partial + (size_t)gwarp * k;
    for (int i = lane; i < k; i += 32) my_row[i] = 0.0;
    __syncwarp();
    const int64_t n_tiles = (n + SW_TILE - 1) / SW_TILE;
    int E = -200;          // 2^E bounds every |g| accumulated in the bins (none yet)
    float scale_f = 0.f;   // 2^(SW_QBITS - E)
    int hot = -2;          // code summed in a register (chosen at the warp's first tile)
    long long hot_acc = 0;
    int pending_tiles = 0;
    unsigned long long bad = 0;
    auto flush = [&]() {  // bins (and the hot register) -> this warp's float64 row; exact integers leave, doubles arrive
        if (pending_tiles == 0) return;
        __syncwarp();
        const double inv = ldexp(1.0, E - SW_QBITS);
        for (int i = lane; i < k; i += 32) {
            const long long b = ((long long)my_hi[i] << 32) + (long long)my_lo[i];
            if (b) {
                my_row[i] = __dadd_rn(my_row[i], __dmul_rn((double)b, inv));
                my_lo[i] = 0;
                my_hi[i] = 0;
            }
        }
        const long long h = warp_sum_ll(hot_acc);
        if (lane == 0 && h) my_row[hot] = __dadd_rn(my_row[hot], __dmul_rn((double)h, inv));
        hot_acc = 0;
        pending_tiles = 0;
        __syncwarp();
    };
    for (int64_t t = gwarp; t < n_tiles; t += n_warps) {
        const int64_t base = t * SW_TILE;
        float g[16];
        int c[16];
        if (vec_ok && base + SW_TILE <= n) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int64_t e = base + (v * 32 + lane) * 4;
                const float4 x = ld_stream_f4(grad + e);
                g[4 * v] = x.x, g[4 * v + 1] = x.y, g[4 * v + 2] = x.z, g[4 * v + 3] = x.w;
                int cc[4];
                load_codes4(codes, bits, e >> 2, n, cc);
                c[4 * v] = cc[0], c[4 * v + 1] = cc[1], c[4 * v + 2] = cc[2], c[4 * v + 3] = cc[3];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int64_t e = base + (j >> 2) * 128 + lane * 4 + (j & 3);
                const bool in = e < n;
                g[j] = in ? grad[e] : 0.f;
                c[j] = in ? load_code1(codes, bits, e, n) : -1;
            }
        }
        float mx = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (c[j] >= k) {
                bad++;
                c[j] = -1;
            }
            if (c[j] >= 0) mx = fmaxf(mx, fabsf(g[j]));
        }
        mx = warp_max_f(mx);
        if (!(mx < INFINITY)) {  // a NaN / infinity: reported like an invalid code, never accumulated
            bad += 1;
            continue;
        }
        if (hot == -2) {  // first tile of this warp: the code most lanes see in their first element
            const unsigned same = __match_any_sync(0xffffffffu, c[0]);
            int votes = c[0] >= 0 ? __popc(same) : 0;
            int best = votes, who = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int ob = __shfl_xor_sync(0xffffffffu, best, o), ow = __shfl_xor_sync(0xffffffffu, who, o);
                if (ob > best || (ob == best && ow < who)) {
                    best = ob;
                    who = ow;
                }
            }
            hot = __shfl_sync(0xffffffffu, c[0], who);
            if (best == 0) hot = -1;
        }
        if (mx > 0.f) {
            int te = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 126;  // |g| < 2^te (denormals: -126)
            if (te < -126) te = -126;
            if (te > E) {  // the bound grows: what is in the bins was scaled with the old one
                flush();
                E = te + 1;
                scale_f = __int_as_float((SW_QBITS - E + 127) << 23);  // SW_QBITS - E in [-85, 169]: clamp below
                if (SW_QBITS - E > 127) scale_f = 0.f;
            }
        }
        if (scale_f != 0.f) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int code = c[j];
                if (code < 0) continue;
                const long long q = __float2ll_rn(__fmul_rn(g[j], scale_f));
                if (code == hot) {
                    hot_acc += q;
                } else {
                    const uint32_t lo = (uint32_t)q;
                    const uint32_t old = atomicAdd(&my_lo[code], lo);
                    atomicAdd(&my_hi[code], (int32_t)(q >> 32) + (int32_t)(old + lo < old));
                }
            }
        } else {  // bound below 2^-84: the float64 scaling (tiny gradients only)
            const double sc = ldexp(1.0, SW_QBITS - E);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int code = c[j];
                if (code < 0) continue;
                const long long q = __double2ll_rn(__dmul_rn((double)g[j], sc));
                if (code == hot) {
                    hot_acc += q;
                } else {
                    const uint32_t lo = (uint32_t)q;
                    const uint32_t old = atomicAdd(&my_lo[code], lo);
                    atomicAdd(&my_hi[code], (int32_t)(q >> 32) + (int32_t)(old + lo < old));
                }
            }
        }
        if (++pending_tiles >= SW_FLUSH_TILES) flush();
    }
    flush();
    bad = warp_sum_ull(bad);
    if (lane == 0 && bad) atomicAdd(bad_codes, bad);
}

// rows of `partial` folded in row order: thread t of the bin's CTA adds rows t, t + 128, ... and the 128 partial sums are
// combined by a fixed tree -- the same additions in the same order on every run
__global__ void __launch_bounds__(128) segsum_fold_kernel(const double *__restrict__ partial, int64_t rows, int k, double *out) {
    __shared__ double s[128];
    const int j = blockIdx.x;
    double acc = 0.0;
    for (int64_t r = threadIdx.x; r < rows; r += 128) acc = __dadd_rn(acc, partial[(size_t)r * k + j]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] = __dadd_rn(s[threadIdx.x], s[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[j] = s[0];
}

__global__ void segsum_final_kernel(const double *partial, int nblocks, int k, double *out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s = __dadd_rn(s, partial[(size_t)b * k + j]);
    out[j] = s;
}

void grad_segsum_device(nnc_ctx *ctx, const float *d_grad, const void *d_codes, int64_t n, int bits, int k,
                        double *h_out) {
    if (k < 1 || k > 65536) NNC_FAIL(NNC_ERR_UNSUPPORTED, "segsum: k = %d outside [1, 65536]", k);
    if (bits < 0 || bits > 16) NNC_FAIL(NNC_ERR_BAD_ARG, "segsum: bits = %d outside [0, 16]", bits);
    const bool priv = k <= 1024;
    if (!priv && 16 * (size_t)k > 200 * 1024) NNC_FAIL(NNC_ERR_UNSUPPORTED, "segsum: k = %d does not fit shared memory", k);
    double *d_out = arena_alloc_t<double>(ctx, k);
    unsigned long long *bad = arena_alloc_t<unsigned long long>(ctx, 1);
    NNC_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), ctx->stream));
    const int vec_ok = ((reinterpret_cast<uintptr_t>(d_grad) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d_codes) & 15u) == 0);
    if (priv) {  // one independent accumulator per warp (segsum_warp_kernel)
        const int64_t n_wtiles = (n + SW_TILE - 1) / SW_TILE;
        const size_t smem = (size_t)SG_WARPS * 2 * k * sizeof(uint32_t);
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * per_sm, (n_wtiles + SG_WARPS - 1) / SG_WARPS));
        const int64_t rows = (int64_t)grid * SG_WARPS;
        double *partial = arena_alloc_t<double>(ctx, (size_t)rows * k);
        func_dyn_smem(ctx, (const void *)segsum_warp_kernel, smem);
        NNC_LAUNCH(ctx, segsum_warp_kernel, grid, SG_THREADS, smem, d_grad, d_codes, n, bits, k, vec_ok, partial, bad);
        NNC_LAUNCH(ctx, segsum_fold_kernel, k, 128, 0, partial, rows, k, d_out);
    } else {
        const int64_t n_tiles = (n + SG_TILE - 1) / SG_TILE;
        const int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 4, n_tiles);
        double *partial = arena_alloc_t<double>(ctx, (size_t)grid * k);
        const size_t smem = 16 * (size_t)k;
        func_dyn_smem(ctx, (const void *)segsum_kernel<false>, smem);
        NNC_LAUNCH(ctx, segsum_kernel<false>, grid, SG_THREADS, smem, d_grad, d_codes, n, bits, k, vec_ok, partial, bad);
        NNC_LAUNCH(ctx, segsum_final_kernel, (k + 127) / 128, 128, 0, partial, grid, k, d_out);
    }
    unsigned long long h_bad = 0;
    if (ctx->world > 1) {
        // every rank's k partial sums travel as bit patterns in its own slot (the other slots are zero, so the integer
        // all-reduce is an all-gather); they are added in rank order on the host: the same double on every rank
        const int world = ctx->world;
        double *slots = arena_alloc_t<double>(ctx, (size_t)world * k);
        NNC_CUDA(cudaMemsetAsync(slots, 0, sizeof(double) * (size_t)world * k, ctx->stream));
        NNC_CUDA(cudaMemcpyAsync(slots + (size_t)ctx->rank * k, d_out, sizeof(double) * k, cudaMemcpyDeviceToDevice, ctx->stream));
        comm_allreduce(ctx, reinterpret_cast<int64_t *>(slots), world * k, 0);
        comm_allreduce(ctx, reinterpret_cast<int64_t *>(bad), 1, 0);
        std::vector<double> all((size_t)world * k);
        NNC_CUDA(cudaMemcpyAsync(all.data(), slots, sizeof(double) * (size_t)world * k, cudaMemcpyDeviceToHost, ctx->stream));
        NNC_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(h_bad), cudaMemcpyDeviceToHost, ctx->stream));
        NNC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h_bad) NNC_FAIL(NNC_ERR_BAD_ARG, "segsum: %llu codes >= k = %d", h_bad, k);
        for (int j = 0; j < k; ++j) {
            double acc = 0.0;
            for (int r = 0; r < world; ++r) acc += all[(size_t)r * k + j];
            h_out[j] = acc;
        }
        return;
    }
    NNC_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(h_bad), cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_bad) NNC_FAIL(NNC_ERR_BAD_ARG, "segsum: %llu codes >= k = %d", h_bad, k);
}

}  // namespace nnc
