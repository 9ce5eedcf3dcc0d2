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

struct EmitDevice {
    RegionTable tab;
    float c[TB_KMAX];                  // centred centroids by id the labels are taken against (table input)
    float cfin[TB_KMAX];               // final centred centroids (inertia is measured against these)
    float values[TB_KMAX];             // codebook by id
    unsigned long long hist[TB_KMAX];  // code histogram by id
    unsigned long long inertia_q;      // fixed-point sum of fl32 squared distances
};

__global__ void __launch_bounds__(TB_THREADS) emit_table_kernel(EmitDevice *ed, int k, float xabs_max) {
    __shared__ TableScratch S;
    build_region_table(ed->c, k, xabs_max, &ed->tab, S);
}

constexpr int EM_THREADS = 256;
constexpr int EM_PER = 8;  // weights per thread

template <bool VEC>
__global__ void __launch_bounds__(EM_THREADS) emit_kernel(const float *__restrict__ w, int64_t n, EmitDevice *ed, float mean,
                                                          double inertia_scale, int32_t *labels, float *ris,
                                                          uint8_t *packed, int bits, int want_hist, int want_inertia) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RegionTable &T = ed->tab;
    const int R = T.R, k = T.k;
    float *s_start = reinterpret_cast<float *>(smem_raw);                 // R + 1 region starts
    int *s_rid = reinterpret_cast<int *>(s_start + (R + 1));              // R: cluster id of a SAFE region, -1 for a ZONE
    float *s_val = reinterpret_cast<float *>(s_rid + R);                  // k codebook values
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_val + k);           // k counters
    for (int i = threadIdx.x; i <= R; i += EM_THREADS) s_start[i] = T.rstart[i];
    for (int i = threadIdx.x; i < R; i += EM_THREADS) s_rid[i] = T.rJ1[i] == T.rJ2[i] ? T.down[T.rJ1[i]] : -1;
    for (int i = threadIdx.x; i < k; i += EM_THREADS) {
        s_val[i] = ed->values[i];
        s_hist[i] = 0;
    }
    __syncthreads();

    unsigned long long inert = 0;
    const int64_t n_chunks = (n + EM_PER - 1) / EM_PER;
    const int64_t total_bytes = (n * bits + 7) / 8;
    const int64_t stride = (int64_t)gridDim.x * EM_THREADS;
    const int64_t iters = (n_chunks + stride - 1) / stride;  // uniform trip count (warp collectives below)
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t chunk = it * stride + (int64_t)blockIdx.x * EM_THREADS + threadIdx.x;
        const int64_t base = chunk * EM_PER;
        const bool live = chunk < n_chunks;
        float x[EM_PER];
        int cnt = 0;
        if (live) {
            cnt = (int)min((int64_t)EM_PER, n - base);
            if (VEC && cnt == EM_PER) {
                float4 a = ld_stream_f4(w + base), b = ld_stream_f4(w + base + 4);
                x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w;
                x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < EM_PER; ++j) x[j] = j < cnt ? w[base + j] : 0.f;
            }
        }
        int id[EM_PER];
#pragma unroll
        for (int j = 0; j < EM_PER; ++j) {
            id[j] = 0;
            if (j < cnt) {
                const float xc = fsub(x[j], mean);
                int lo = 0, hi = R;  // largest r with start[r] <= xc  (start[0] = -inf, start[R] = +inf)
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (s_start[mid] <= xc)
                        lo = mid;
                    else
                        hi = mid;
                }
                const int rid = s_rid[lo];
                if (rid < 0)
                    id[j] = T.down[zone_argmin(xc, T.dv, T.dcn, T.down, T.rJ2[lo], T.rJ1[lo])];
                else
                    id[j] = rid;
                if (want_inertia) {
                    // distance to the centroid in centred space, as sklearn's _inertia_dense computes it
                    float t = fsub(xc, ed->cfin[id[j]]);
                    float d2 = fmul(t, t);
                    inert += (unsigned long long)__double2ll_rn(__dmul_rn((double)d2, inertia_scale));
                }
            }
        }
        if (want_hist) {
#pragma unroll
            for (int j = 0; j < EM_PER; ++j) {
                int key = j < cnt ? id[j] : -1;
                uint32_t peers = __match_any_sync(0xffffffffu, key);
                if (key >= 0 && (int)(__ffs(peers) - 1) == lane_id()) atomicAdd(&s_hist[key], (uint32_t)__popc(peers));
            }
        }
        if (!live) continue;
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
                st_stream_f4(ris + base, make_float4(s_val[id[0]], s_val[id[1]], s_val[id[2]], s_val[id[3]]));
                st_stream_f4(ris + base + 4, make_float4(s_val[id[4]], s_val[id[5]], s_val[id[6]], s_val[id[7]]));
            } else {
                for (int j = 0; j < cnt; ++j) ris[base + j] = s_val[id[j]];
            }
        }
        if (packed) {
            // 8 codes -> `bits` bytes, little-endian bit stream
            unsigned long long lo64 = 0, hi64 = 0;
#pragma unroll
            for (int j = 0; j < EM_PER; ++j) {
                const unsigned long long v = (unsigned long long)(uint32_t)id[j];
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
    if (want_inertia) {
        inert = warp_sum_ull(inert);
        if (lane_id() == 0 && inert) atomicAdd(&ed->inertia_q, inert);
    }
    if (want_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < k; i += EM_THREADS)
            if (s_hist[i]) atomicAdd(&ed->hist[i], (unsigned long long)s_hist[i]);
    }
}

void emit_device(nnc_ctx *ctx, const float *d_w, int64_t n, const float *h_centred, const float *h_centred_final, int k,
                 float mean, float xabs, const float *h_values, int32_t *d_labels, float *d_ris, uint8_t *d_packed, int bits,
                 int64_t *h_hist, double *h_inertia) {
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
    }
    float cmax = xabs;
    for (int j = 0; j < k; ++j) cmax = fmaxf(cmax, fabsf(h_centred_final[j]));
    int Ed = cmax > 0.f ? ilogbf(cmax) + 1 : 0;
    const double inertia_scale = ldexp(1.0, 31 - (2 * Ed + 2));
    NNC_LAUNCH(ctx, emit_table_kernel, 1, TB_THREADS, 0, ed, k, xabs);
    const bool vec = ((reinterpret_cast<uintptr_t>(d_w) & 15u) == 0) &&
                     (!d_labels || (reinterpret_cast<uintptr_t>(d_labels) & 15u) == 0) &&
                     (!d_ris || (reinterpret_cast<uintptr_t>(d_ris) & 15u) == 0) &&
                     (!d_packed || (reinterpret_cast<uintptr_t>(d_packed) & 15u) == 0);
    const int64_t n_chunks = (n + EM_PER - 1) / EM_PER;
    const int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 8, (n_chunks + EM_THREADS - 1) / EM_THREADS);
    const size_t smem = sizeof(float) * (2 * (size_t)k + 2) + sizeof(int) * (2 * (size_t)k + 1) + sizeof(float) * k +
                        sizeof(uint32_t) * k;
    const int want_hist = h_hist ? 1 : 0, want_inertia = h_inertia ? 1 : 0;
    if (vec)
        NNC_LAUNCH(ctx, emit_kernel<true>, grid, EM_THREADS, smem, d_w, n, ed, mean, inertia_scale, d_labels, d_ris, d_packed,
                   bits, want_hist, want_inertia);
    else
        NNC_LAUNCH(ctx, emit_kernel<false>, grid, EM_THREADS, smem, d_w, n, ed, mean, inertia_scale, d_labels, d_ris,
                   d_packed, bits, want_hist, want_inertia);
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
