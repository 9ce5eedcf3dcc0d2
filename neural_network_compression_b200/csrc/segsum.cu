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

__global__ void __launch_bounds__(SG_THREADS) segsum_kernel(const float *__restrict__ grad, const void *__restrict__ codes,
                                                            int64_t n, int bits, int k, int vec_ok, double *partial,
                                                            unsigned long long *bad_codes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    long long *s_bin = reinterpret_cast<long long *>(smem_raw);
    double *s_dbl = reinterpret_cast<double *>(s_bin + k);
    __shared__ float s_max[SG_THREADS / 32];
    __shared__ float s_tile_max;
    for (int i = threadIdx.x; i < k; i += SG_THREADS) {
        s_bin[i] = 0;
        s_dbl[i] = 0.0;
    }
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
        // |g| < 2^E ; q = rint(g * 2^(50-E)) < 2^50 ; 4096 of them < 2^62
        int E = 0;
        if (tmax > 0.f) E = (int)((__float_as_uint(tmax) >> 23) & 0xffu) - 126;  // ilogb + 1 (denormals: E = -126)
        if (tmax > 0.f && tmax < 1.17549435e-38f) E = -126;
        const double scale = ldexp(1.0, 50 - E);
#pragma unroll
        for (int v = 0; v < SG_VEC; ++v) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int code = c[v][j];
                long long q = code >= 0 ? __double2ll_rn(__dmul_rn((double)g[v][j], scale)) : 0;
                // warp aggregation of the most populated code in the warp, individual atomics for the rest
                uint32_t peers = __match_any_sync(0xffffffffu, code);
                int sz = __popc(peers);
                int best = sz;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
                bool done = code < 0;
                if (best >= 4) {
                    unsigned cand = __ballot_sync(0xffffffffu, sz == best && code >= 0);
                    if (cand) {
                        int leader = __ffs(cand) - 1;
                        int Lc = __shfl_sync(0xffffffffu, code, leader);
                        long long s = warp_sum_ll(code == Lc ? q : 0);
                        if (lane_id() == leader) atomicAdd((unsigned long long *)&s_bin[Lc], (unsigned long long)s);
                        if (code == Lc) done = true;
                    }
                }
                if (!done) atomicAdd((unsigned long long *)&s_bin[code], (unsigned long long)q);
            }
        }
        __syncthreads();
        const double inv = ldexp(1.0, E - 50);
        for (int i = threadIdx.x; i < k; i += SG_THREADS) {
            long long b = s_bin[i];
            if (b) {
                s_dbl[i] = __dadd_rn(s_dbl[i], __dmul_rn((double)b, inv));
                s_bin[i] = 0;
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < k; i += SG_THREADS) partial[(size_t)blockIdx.x * k + i] = s_dbl[i];
    bad = warp_sum_ull(bad);
    if (lane_id() == 0 && bad) atomicAdd(bad_codes, bad);
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
    if (16 * (size_t)k > 200 * 1024) NNC_FAIL(NNC_ERR_UNSUPPORTED, "segsum: k = %d does not fit shared memory", k);
    const int64_t n_tiles = (n + SG_TILE - 1) / SG_TILE;
    const int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 4, n_tiles);
    double *partial = arena_alloc_t<double>(ctx, (size_t)grid * k);
    double *d_out = arena_alloc_t<double>(ctx, k);
    unsigned long long *bad = arena_alloc_t<unsigned long long>(ctx, 1);
    NNC_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), ctx->stream));
    const size_t smem = 16 * (size_t)k;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        NNC_CUDA(cudaFuncSetAttribute(segsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int vec_ok = ((reinterpret_cast<uintptr_t>(d_grad) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d_codes) & 15u) == 0) &&
                       (bits == 0 || bits == 4 || bits == 8 || bits == 16 || bits == 2 || bits == 1 || true);
    NNC_LAUNCH(ctx, segsum_kernel, grid, SG_THREADS, smem, d_grad, d_codes, n, bits, k, vec_ok, partial, bad);
    NNC_LAUNCH(ctx, segsum_final_kernel, (k + 127) / 128, 128, 0, partial, grid, k, d_out);
    unsigned long long h_bad = 0;
    NNC_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(h_bad), cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_bad) NNC_FAIL(NNC_ERR_BAD_ARG, "segsum: %llu codes >= k = %d", h_bad, k);
}

}  // namespace nnc
