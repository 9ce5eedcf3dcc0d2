// emit.cu -- the final E-step over the tensor in ORIGINAL order: labels_, the de-quantised tensor
// (cluster_centers_[labels_], utility.py:239), packed n-bit cluster indices and the code histogram.
//
// One streaming pass: 4 B read per weight, bits/8 B (+4 B labels, +4 B ris when requested) written.
// The label of every element is the scikit-learn float32 rule (_k_means_lloyd.pyx:196-213), evaluated
// through the region table (table.cuh): a binary search over the region starts in shared memory gives a
// SAFE region (label known) or a ZONE (evaluate the rule over the zone's few candidates).
//
// Packed layout: code i occupies bits [i*bits, (i+1)*bits) of a little-endian byte stream.  A thread owns
// 8 consecutive weights, hence exactly `bits` consecutive bytes of the stream.
#include <algorithm>

#include "common.cuh"
#include "internal.h"
#include "table.cuh"

namespace nnc {

constexpr int EM_LUT = 16384;      // buckets of the label look-up table over [x'_min, x'_max]
constexpr uint16_t EM_SLOW = 0x8000;  // LUT entry: bit 15 set -> low bits = first candidate region, resolve by search

struct EmitDevice {
    RegionTable tab;
    float c[TB_KMAX];                  // centred centroids by id the labels are taken against (table input)
    float cfin[TB_KMAX];               // final centred centroids (inertia is measured against these)
    float values[TB_KMAX];             // codebook by id
    unsigned long long hist[TB_KMAX];  // code histogram by id
    unsigned long long inertia_q;      // fixed-point sum of fl32 squared distances
    int rid[2 * TB_KMAX + 2];          // cluster id of a SAFE region, -1 for a ZONE
    int zid;                           // cluster id of the value 0.0 (the pruned weights)
    float lut_lo, lut_scale;           // bucket(x') = clamp(int((x' - lut_lo) * lut_scale), 0, EM_LUT - 1)
    alignas(16) uint32_t lut_cnt[EM_LUT];  // breakpoints per bucket (scratch of the table kernel)
    alignas(16) uint16_t lut[EM_LUT];
};

// Monotone non-decreasing in xc: float subtraction, multiplication by a non-negative scale, truncation and the
// clamp all preserve order, so "bucket(T) < b  =>  T < every x' of bucket b" and "bucket(T) > b  =>  T > ...".
__device__ __forceinline__ int lut_bucket(float xc, float lo, float scale) {
    int b = __float2int_rz(fmul(fsub(xc, lo), scale));
    return min(max(b, 0), EM_LUT - 1);
}

// region of xc: number of breakpoints rstart[1..R-1] that are <= xc, scanning upwards from region r0
__device__ __forceinline__ int region_from(const float *s_start, int R, int r0, float xc) {
    int r = r0;
    while (r + 1 < R && s_start[r + 1] <= xc) ++r;
    return r;
}

__global__ void __launch_bounds__(TB_THREADS) emit_table_kernel(EmitDevice *ed, int k, float xabs_max, float mean, float lut_lo,
                                                                float lut_scale) {
    __shared__ TableScratch S;
    __shared__ uint32_t s_scan[32];
    build_region_table(ed->c, k, xabs_max, &ed->tab, S);
    const RegionTable &T = ed->tab;
    const int R = T.R, tid = threadIdx.x;
    for (int r = tid; r < R; r += TB_THREADS) ed->rid[r] = T.rJ1[r] == T.rJ2[r] ? T.down[T.rJ1[r]] : -1;
    for (int b = tid; b < EM_LUT; b += TB_THREADS) ed->lut_cnt[b] = 0;
    __syncthreads();
    for (int r = 1 + tid; r < R; r += TB_THREADS) atomicAdd(&ed->lut_cnt[lut_bucket(T.rstart[r], lut_lo, lut_scale)], 1u);
    __syncthreads();
    // exclusive scan of the bucket counts: region of the first x' of every bucket
    constexpr int PER = EM_LUT / TB_THREADS;
    uint32_t loc[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        loc[i] = ed->lut_cnt[tid * PER + i];
        sum += loc[i];
    }
    uint32_t incl = block_scan_incl<uint32_t>(sum, [](uint32_t a, uint32_t b) { return a + b; }, s_scan);
    uint32_t run = incl - sum;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int b = tid * PER + i;
        const int rid = ed->rid[run];
        ed->lut[b] = (loc[i] == 0 && rid >= 0) ? (uint16_t)rid : (uint16_t)(EM_SLOW | run);
        run += loc[i];
    }
    if (tid == 0) {  // label of the pruned weights (value 0.0)
        const float x0 = fsub(0.f, mean);
        int r = 0;
        while (r + 1 < R && T.rstart[r + 1] <= x0) ++r;
        ed->zid = T.rJ1[r] == T.rJ2[r] ? T.down[T.rJ1[r]] : T.down[zone_argmin(x0, T.dv, T.dcn, T.down, T.rJ2[r], T.rJ1[r])];
        ed->lut_lo = lut_lo;
        ed->lut_scale = lut_scale;
    }
}

constexpr int EM_THREADS = 256;
constexpr int EM_PER = 8;  // weights per thread

struct EmitSmem {
    alignas(16) uint16_t lut[EM_LUT];
    float start[2 * TB_KMAX + 2];
    int rid[2 * TB_KMAX + 2];
    float val[TB_KMAX];
    uint32_t hist[TB_KMAX];
    unsigned long long red[EM_THREADS / 32];
};

template <bool VEC, bool INERTIA>
__global__ void __launch_bounds__(EM_THREADS) emit_kernel(const float *__restrict__ w, int64_t n, EmitDevice *ed, float mean,
                                                          double inertia_scale, int32_t *labels, float *ris,
                                                          uint8_t *packed, int bits, int want_hist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitSmem &S = *reinterpret_cast<EmitSmem *>(smem_raw);
    const RegionTable &T = ed->tab;
    const int R = T.R, k = T.k;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(ed->lut);
        uint4 *dst = reinterpret_cast<uint4 *>(S.lut);
        for (int i = threadIdx.x; i < EM_LUT * 2 / 16; i += EM_THREADS) dst[i] = src[i];
    }
    for (int i = threadIdx.x; i <= R; i += EM_THREADS) S.start[i] = T.rstart[i];
    for (int i = threadIdx.x; i < R; i += EM_THREADS) S.rid[i] = ed->rid[i];
    for (int i = threadIdx.x; i < k; i += EM_THREADS) {
        S.val[i] = ed->values[i];
        S.hist[i] = 0;
    }
    __syncthreads();
    const int zid = ed->zid;
    const float lut_lo = ed->lut_lo, lut_scale = ed->lut_scale;
    const float zval = S.val[zid];

    unsigned long long inert = 0;
    unsigned int zcount = 0;
    const int64_t n_chunks = (n + EM_PER - 1) / EM_PER;
    const int64_t total_bytes = (n * bits + 7) / 8;
    const int64_t stride = (int64_t)gridDim.x * EM_THREADS;
    for (int64_t chunk = (int64_t)blockIdx.x * EM_THREADS + threadIdx.x; chunk < n_chunks; chunk += stride) {
        const int64_t base = chunk * EM_PER;
        float x[EM_PER];
        const int cnt = (int)min((int64_t)EM_PER, n - base);
        if (VEC && cnt == EM_PER) {
            float4 a = ld_stream_f4(w + base), b = ld_stream_f4(w + base + 4);
            x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w;
            x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < EM_PER; ++j) x[j] = j < cnt ? w[base + j] : 0.f;
        }
        int id[EM_PER];
#pragma unroll
        for (int j = 0; j < EM_PER; ++j) {
            if (x[j] == 0.f) {  // pruned weight (or padding): one known label, counted in a register
                id[j] = zid;
                zcount += j < cnt;
            } else {
                const float xc = fsub(x[j], mean);
                const uint32_t e = S.lut[lut_bucket(xc, lut_lo, lut_scale)];
                int lab = (int)e;
                if (e & EM_SLOW) {
                    const int r = region_from(S.start, R, (int)(e & (EM_SLOW - 1)), xc);
                    lab = S.rid[r];
                    if (lab < 0) lab = T.down[zone_argmin(xc, T.dv, T.dcn, T.down, T.rJ2[r], T.rJ1[r])];
                }
                id[j] = lab;
                if (want_hist) atomicAdd(&S.hist[lab], 1u);
            }
            if (INERTIA && j < cnt) {
                // distance to the centroid in centred space, as sklearn's _inertia_dense computes it
                const float t = fsub(fsub(x[j], mean), ed->cfin[id[j]]);
                const float d2 = fmul(t, t);
                inert += (unsigned long long)__double2ll_rn(__dmul_rn((double)d2, inertia_scale));
            }
        }
        if (labels) {
            if (VEC && cnt == EM_PER) {
                reinterpret_cast<int4 *>(labels + base)[0] = make_int4(id[0], id[1], id[2], id[3]);
                reinterpret_cast<int4 *>(labels + base)[1] = make_int4(id[4], id[5], id[6], id[7]);
            } else {
                for (int j = 0; j < cnt; ++j) labels[base + j] = id[j];
            }
        }
        if (ris) {
            if (VEC && cnt == EM_PER) {
                st_stream_f4(ris + base, make_float4(S.val[id[0]], S.val[id[1]], S.val[id[2]], S.val[id[3]]));
                st_stream_f4(ris + base + 4, make_float4(S.val[id[4]], S.val[id[5]], S.val[id[6]], S.val[id[7]]));
            } else {
                for (int j = 0; j < cnt; ++j) ris[base + j] = S.val[id[j]];
            }
        }
        if (packed) {
            // 8 codes -> `bits` bytes, little-endian bit stream
            unsigned long long lo64 = 0, hi64 = 0;
#pragma unroll
            for (int j = 0; j < EM_PER; ++j) {
                const unsigned long long v = j < cnt ? (unsigned long long)(uint32_t)id[j] : 0ull;
                const int sh = j * bits;
                if (sh < 64) {
                    lo64 |= v << sh;
                    if (sh + bits > 64) hi64 |= v >> (64 - sh);
                } else {
                    hi64 |= v << (sh - 64);
                }
            }
            const int64_t byte0 = chunk * bits;
            uint8_t *dst = packed + byte0;
            const bool full = byte0 + bits <= total_bytes;
            if (VEC && full && bits == 8) {
                *reinterpret_cast<unsigned long long *>(dst) = lo64;
            } else if (VEC && full && bits == 4) {
                *reinterpret_cast<uint32_t *>(dst) = (uint32_t)lo64;
            } else if (VEC && full && bits == 2) {
                *reinterpret_cast<uint16_t *>(dst) = (uint16_t)lo64;
            } else if (VEC && full && bits == 16) {
                reinterpret_cast<unsigned long long *>(dst)[0] = lo64;
                reinterpret_cast<unsigned long long *>(dst)[1] = hi64;
            } else {
                for (int b = 0; b < bits; ++b) {
                    if (byte0 + b < total_bytes) dst[b] = (uint8_t)(b < 8 ? (lo64 >> (8 * b)) : (hi64 >> (8 * (b - 8))));
                }
            }
        }
    }
    (void)zval;
    if (INERTIA) {
        inert = warp_sum_ull(inert);
        if (lane_id() == 0 && inert) atomicAdd(&ed->inertia_q, inert);
    }
    if (want_hist) {
        unsigned long long z = warp_sum_ull((unsigned long long)zcount);
        if (lane_id() == 0) S.red[warp_id()] = z;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int i = 0; i < EM_THREADS / 32; ++i) t += S.red[i];
            if (t) atomicAdd(&ed->hist[zid], t);
        }
        for (int i = threadIdx.x; i < k; i += EM_THREADS)
            if (S.hist[i]) atomicAdd(&ed->hist[i], (unsigned long long)S.hist[i]);
    }
}

void emit_device(nnc_ctx *ctx, const float *d_w, int64_t n, const float *h_centred, const float *h_centred_final, int k,
                 float mean, float xabs, float xlo, float xhi, const float *h_values, int32_t *d_labels, float *d_ris,
                 uint8_t *d_packed, int bits, int64_t *h_hist, double *h_inertia) {
    if (!h_centred_final) h_centred_final = h_centred;
    if (k < 1 || k > TB_KMAX) NNC_FAIL(NNC_ERR_UNSUPPORTED, "emit: k = %d outside [1, %d]", k, TB_KMAX);
    if (d_packed) {
        int need = 0;
        while ((1 << need) < k) need++;
        if (bits < need || bits < 1 || bits > 16)
            NNC_FAIL(NNC_ERR_BAD_ARG, "emit: bits = %d cannot hold %d clusters (need %d..16)", bits, k, std::max(need, 1));
    }
    EmitDevice *ed = arena_alloc_t<EmitDevice>(ctx, 1);
    NNC_CUDA(cudaMemsetAsync(ed, 0, sizeof(EmitDevice), ctx->stream));
    NNC_CUDA(cudaMemcpyAsync(ed->c, h_centred, sizeof(float) * k, cudaMemcpyHostToDevice, ctx->stream));
    NNC_CUDA(cudaMemcpyAsync(ed->cfin, h_centred_final, sizeof(float) * k, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<float> vals(k);
    for (int j = 0; j < k; ++j) {
        volatile float v = h_centred_final[j] + mean;  // cluster_centers_ = centres + X_mean (float32)
        vals[j] = h_values ? h_values[j] : v;
    }
    NNC_CUDA(cudaMemcpyAsync(ed->values, vals.data(), sizeof(float) * k, cudaMemcpyHostToDevice, ctx->stream));
    // bound on |x'| and |c'| for the zone widths: data range from a min/max sweep
    if (xabs < 0.f) {
        float mn, mx;
        int64_t cnt;
        minmax_device(ctx, d_w, n, 0, &mn, &mx, &cnt);
        volatile float a = mn - mean, b = mx - mean;
        xabs = fmaxf(fabsf(a), fabsf(b));
        xlo = a;
        xhi = b;
    }
    float cmax = xabs;
    for (int j = 0; j < k; ++j) cmax = fmaxf(cmax, fabsf(h_centred_final[j]));
    int Ed = cmax > 0.f ? ilogbf(cmax) + 1 : 0;
    const double inertia_scale = ldexp(1.0, 31 - (2 * Ed + 2));
    // label LUT over the data range [x'_min, x'_max] (callers that do not know the range pass the symmetric bound)
    if (!(xlo <= xhi)) {
        xlo = -xabs;
        xhi = xabs;
    }
    volatile float span = xhi - xlo;
    float lut_scale = span > 0.f ? (float)((double)EM_LUT / (double)span * (1.0 - 1e-6)) : 0.f;
    if (!isfinite(lut_scale)) lut_scale = 0.f;
    NNC_LAUNCH(ctx, emit_table_kernel, 1, TB_THREADS, 0, ed, k, xabs, mean, xlo, lut_scale);
    const bool vec = ((reinterpret_cast<uintptr_t>(d_w) & 15u) == 0) &&
                     (!d_labels || (reinterpret_cast<uintptr_t>(d_labels) & 15u) == 0) &&
                     (!d_ris || (reinterpret_cast<uintptr_t>(d_ris) & 15u) == 0) &&
                     (!d_packed || (reinterpret_cast<uintptr_t>(d_packed) & 15u) == 0);
    const int64_t n_chunks = (n + EM_PER - 1) / EM_PER;
    const int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 4, (n_chunks + EM_THREADS - 1) / EM_THREADS);
    const size_t smem = sizeof(EmitSmem);
    const int want_hist = h_hist ? 1 : 0;
    static bool configured = false;
    if (!configured) {
        NNC_CUDA(cudaFuncSetAttribute(emit_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NNC_CUDA(cudaFuncSetAttribute(emit_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NNC_CUDA(cudaFuncSetAttribute(emit_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        NNC_CUDA(cudaFuncSetAttribute(emit_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    if (vec && h_inertia)
        NNC_LAUNCH(ctx, (emit_kernel<true, true>), grid, EM_THREADS, smem, d_w, n, ed, mean, inertia_scale, d_labels, d_ris,
                   d_packed, bits, want_hist);
    else if (vec)
        NNC_LAUNCH(ctx, (emit_kernel<true, false>), grid, EM_THREADS, smem, d_w, n, ed, mean, inertia_scale, d_labels, d_ris,
                   d_packed, bits, want_hist);
    else if (h_inertia)
        NNC_LAUNCH(ctx, (emit_kernel<false, true>), grid, EM_THREADS, smem, d_w, n, ed, mean, inertia_scale, d_labels, d_ris,
                   d_packed, bits, want_hist);
    else
        NNC_LAUNCH(ctx, (emit_kernel<false, false>), grid, EM_THREADS, smem, d_w, n, ed, mean, inertia_scale, d_labels, d_ris,
                   d_packed, bits, want_hist);
    if (h_hist || h_inertia) {
        std::vector<unsigned long long> hh(k + 1);
        if (h_hist) NNC_CUDA(cudaMemcpyAsync(hh.data(), ed->hist, sizeof(unsigned long long) * k, cudaMemcpyDeviceToHost, ctx->stream));
        NNC_CUDA(cudaMemcpyAsync(&hh[k], &ed->inertia_q, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        NNC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h_hist)
            for (int j = 0; j < k; ++j) h_hist[j] = (int64_t)hh[j];
        if (h_inertia) *h_inertia = (double)hh[k] / inertia_scale;
    }
}

// out[i] = values[code_i]
__global__ void __launch_bounds__(256) unpack_gather_kernel(const uint8_t *__restrict__ packed, int64_t n, int bits,
                                                            const float *__restrict__ values, int k, float *out) {
    extern __shared__ float s_v[];
    for (int i = threadIdx.x; i < k; i += blockDim.x) s_v[i] = values[i];
    __syncthreads();
    const int64_t total_bytes = (n * bits + 7) / 8;
    const uint32_t mask = (1u << bits) - 1u;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bit = i * bits;
        const int64_t byte = bit >> 3;
        uint32_t v = 0;
        for (int b = 0; b < 3; ++b)
            if (byte + b < total_bytes) v |= (uint32_t)packed[byte + b] << (8 * b);
        uint32_t code = (v >> (bit & 7)) & mask;
        out[i] = s_v[code < (uint32_t)k ? code : 0];
    }
}

void unpack_gather_device(nnc_ctx *ctx, const uint8_t *d_packed, int64_t n, int bits, const float *h_values, int k,
                          float *d_out) {
    if (bits < 1 || bits > 16) NNC_FAIL(NNC_ERR_BAD_ARG, "unpack: bits = %d outside [1, 16]", bits);
    float *d_v = arena_alloc_t<float>(ctx, k);
    NNC_CUDA(cudaMemcpyAsync(d_v, h_values, sizeof(float) * k, cudaMemcpyHostToDevice, ctx->stream));
    int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 16, (n + 255) / 256);
    NNC_LAUNCH(ctx, unpack_gather_kernel, grid, 256, sizeof(float) * k, d_packed, n, bits, d_v, k, d_out);
}

}  // namespace nnc
