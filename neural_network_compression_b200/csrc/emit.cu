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
#include <type_traits>

#include "common.cuh"
#include "internal.h"
#include "table.cuh"

namespace nnc {

constexpr int EM_LUT = 16384;       // coarse buckets of the label look-up table over [x'_min, x'_max]
constexpr int EM_SUB = 32;          // fine buckets per coarse bucket (second level, only for buckets holding a boundary)
constexpr int EM_FINE = EM_LUT * EM_SUB;
constexpr int EM_DMAX = 512;        // coarse buckets that get a second-level table
// LUT entry (16 bit): 00 | cluster id        the whole bucket lies in one SAFE region
//                     01 | table index * 32  second-level table (first level only)
//                     10 | region            resolve by search starting at that region (boundary or ZONE inside)
constexpr uint32_t EM_SLOW = 0x8000, EM_L2 = 0x4000;

struct EmitDevice {
    RegionTable tab;
    float c[TB_KMAX];                  // centred centroids by id the labels are taken against (table input)
    float cfin[TB_KMAX];               // final centred centroids (inertia is measured against these)
    float values[TB_KMAX];             // codebook by id
    unsigned long long hist[TB_KMAX];  // code histogram by id
    unsigned long long inertia_q;      // fixed-point sum of fl32 squared distances
    int rid[2 * TB_KMAX + 2];          // cluster id of a SAFE region, -1 for a ZONE
    int zid;                           // cluster id of the value 0.0 (the pruned weights)
    float lut_lo, lut_scale;           // fine bucket(x') = min(int((x' - lut_lo) * lut_scale), EM_FINE - 1)
    alignas(16) uint16_t lut[EM_LUT];
    alignas(16) uint16_t lut2[EM_DMAX * EM_SUB];
};

// Monotone non-decreasing in xc: float subtraction, multiplication by a non-negative scale, truncation and the
// clamp all preserve order, so "bucket(T) < b  =>  T < every x' of bucket b" and "bucket(T) > b  =>  T > ...".
__device__ __forceinline__ int lut_fine(float xc, float lo, float scale) {
    return (int)min((unsigned)__float2int_rz(fmul(fsub(xc, lo), scale)), (unsigned)(EM_FINE - 1));
}
__device__ __forceinline__ int lut_fine_clamped(float xc, float lo, float scale) {  // for thresholds outside the data range
    const float t = fmul(fsub(xc, lo), scale);
    if (!(t > 0.f)) return 0;
    return (int)min((unsigned)__float2int_rz(t), (unsigned)(EM_FINE - 1));
}

// region of xc: number of breakpoints rstart[1..R-1] that are <= xc, scanning upwards from region r0
__device__ __forceinline__ int region_from(const float *s_start, int R, int r0, float xc) {
    int r = r0;
    while (r + 1 < R && s_start[r + 1] <= xc) ++r;
    return r;
}

// One CTA; every working array lives in shared memory (as global-memory arrays the dependent look-ups of this kernel -- a
// chain of ~50 round trips -- cost 86 us, on the critical path between the Lloyd loop and the emission pass)
struct EmitTabSmem {
    TableScratch S;
    RegionTable T;
    uint32_t scan[32];
    int rid[2 * TB_KMAX + 2];
    uint32_t cnt[EM_LUT];  // boundaries per coarse bucket; afterwards per fine bucket of the buckets with a table
    uint16_t rlo[EM_LUT];  // region of the first x' of every coarse bucket
    uint16_t did[EM_LUT];  // dense index of a bucket that holds boundaries
    float c[TB_KMAX];
};
static_assert(EM_DMAX * EM_SUB <= EM_LUT, "the fine counts reuse the coarse count array");
static_assert(sizeof(EmitTabSmem) <= 227 * 1024, "EmitTabSmem must fit the shared memory a CTA can opt into");

__global__ void __launch_bounds__(TB_THREADS) emit_table_kernel(EmitDevice *ed, int k, float xabs_max, float mean, float lut_lo,
                                                                float lut_scale) {
    extern __shared__ __align__(16) unsigned char tab_smem_raw[];
    EmitTabSmem &M = *reinterpret_cast<EmitTabSmem *>(tab_smem_raw);
    const int tid = threadIdx.x;
    for (int i = tid; i < k; i += TB_THREADS) M.c[i] = ed->c[i];
    __syncthreads();
    build_region_table(M.c, k, xabs_max, &M.T, M.S);
    const RegionTable &T = M.T;
    const int R = T.R;
    for (int r = tid; r < R; r += TB_THREADS) {
        const int id = T.rJ1[r] == T.rJ2[r] ? T.down[T.rJ1[r]] : -1;
        M.rid[r] = id;
        ed->rid[r] = id;
    }
    for (int b = tid; b < EM_LUT; b += TB_THREADS) M.cnt[b] = 0;
    __syncthreads();
    for (int r = 1 + tid; r < R; r += TB_THREADS) atomicAdd(&M.cnt[lut_fine_clamped(T.rstart[r], lut_lo, lut_scale) / EM_SUB], 1u);
    __syncthreads();
    // exclusive scans over the coarse buckets: boundaries before the bucket (-> region of its first x') and
    // buckets holding boundaries before it (-> dense index)
    constexpr int PER = EM_LUT / TB_THREADS;
    uint32_t loc[PER], sum = 0, dsum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        loc[i] = M.cnt[tid * PER + i];
        sum += loc[i];
        dsum += loc[i] != 0;
    }
    uint32_t incl = block_scan_incl<uint32_t>(sum, [](uint32_t a, uint32_t b) { return a + b; }, M.scan);
    uint32_t dincl = block_scan_incl<uint32_t>(dsum, [](uint32_t a, uint32_t b) { return a + b; }, M.scan);
    uint32_t run = incl - sum, drun = dincl - dsum;
    __syncthreads();  // every thread has its counts in registers: the count array becomes the fine counts
    alignas(16) uint16_t ent[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int b = tid * PER + i;
        const int rid = M.rid[run];
        uint32_t e;
        if (loc[i] == 0)
            e = rid >= 0 ? (uint32_t)rid : (EM_SLOW | run);
        else
            e = drun < (uint32_t)EM_DMAX ? (EM_L2 | (drun * EM_SUB)) : (EM_SLOW | run);  // (first index of its table)
        ent[i] = (uint16_t)e;
        M.rlo[b] = (uint16_t)run;
        M.did[b] = (uint16_t)(loc[i] != 0 && drun < (uint32_t)EM_DMAX ? drun : 0xffffu);
        M.cnt[b] = 0;  // (b < EM_DMAX * EM_SUB <= EM_LUT covers the fine counts)
        run += loc[i];
        drun += loc[i] != 0;
    }
    {  // this thread's PER consecutive entries in one go
        static_assert(PER * sizeof(uint16_t) % 16 == 0, "vector store of the LUT entries");
        uint4 *dst = reinterpret_cast<uint4 *>(ed->lut + tid * PER);
        const uint4 *src = reinterpret_cast<const uint4 *>(ent);
#pragma unroll
        for (int i = 0; i < (int)(PER * sizeof(uint16_t) / 16); ++i) dst[i] = src[i];
    }
    __syncthreads();
    // second level: boundaries per fine bucket of the buckets that have a table
    for (int r = 1 + tid; r < R; r += TB_THREADS) {
        const int f = lut_fine_clamped(T.rstart[r], lut_lo, lut_scale);
        const uint32_t did = M.did[f / EM_SUB];
        if (did != 0xffffu) atomicAdd(&M.cnt[did * EM_SUB + (f % EM_SUB)], 1u);
    }
    __syncthreads();
    for (int b = tid; b < EM_LUT; b += TB_THREADS) {
        const uint32_t did = M.did[b];
        if (did == 0xffffu) continue;
        uint32_t r = M.rlo[b];
        for (int f = 0; f < EM_SUB; ++f) {
            const uint32_t c = M.cnt[did * EM_SUB + f];
            const int rid = M.rid[r];
            ed->lut2[did * EM_SUB + f] = (uint16_t)((c == 0 && rid >= 0) ? (uint32_t)rid : (EM_SLOW | r));
            r += c;
        }
    }
    {  // the table itself, for the emission kernel
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&M.T);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&ed->tab);
        for (int i = tid; i < (int)(sizeof(RegionTable) / 4); i += TB_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {  // label of the pruned weights (value 0.0)
        const float x0 = fsub(0.f, mean);
        int lo = 0, hi = R;  // largest r with rstart[r] <= x0 (rstart[0] = -inf)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (T.rstart[mid] <= x0)
                lo = mid;
            else
                hi = mid;
        }
        const int r = lo;
        ed->zid = T.rJ1[r] == T.rJ2[r] ? T.down[T.rJ1[r]] : T.down[zone_argmin(x0, T.dv, T.dcn, T.down, T.rJ2[r], T.rJ1[r])];
        ed->lut_lo = lut_lo;
        ed->lut_scale = lut_scale;
    }
}

constexpr int EM_THREADS = 512;
constexpr int EM_PER = 8;  // weights per thread

struct EmitSmem {
    alignas(16) uint16_t lut[EM_LUT];
    alignas(16) uint16_t lut2[EM_DMAX * EM_SUB];
    float start[2 * TB_KMAX + 2];
    int rid[2 * TB_KMAX + 2];
    float val[TB_KMAX];
    uint32_t hist[TB_KMAX];
    unsigned long long red[EM_THREADS / 32];
};

// BITS: compile-time width of a packed code (8), or 0 for the run-time width `bits`
template <bool VEC, bool INERTIA, int BITS>
__global__ void __launch_bounds__(EM_THREADS, 2) emit_kernel(const float *__restrict__ w, int64_t n, EmitDevice *ed, float mean,
                                                          double inertia_scale, int32_t *labels, float *ris,
                                                          uint8_t *packed, int bits_rt, int want_hist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitSmem &S = *reinterpret_cast<EmitSmem *>(smem_raw);
    const RegionTable &T = ed->tab;
    const int R = T.R, k = T.k;
    const int bits = BITS ? BITS : bits_rt;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(ed->lut);
        uint4 *dst = reinterpret_cast<uint4 *>(S.lut);
        for (int i = threadIdx.x; i < EM_LUT * 2 / 16; i += EM_THREADS) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(ed->lut2);
        dst = reinterpret_cast<uint4 *>(S.lut2);
        for (int i = threadIdx.x; i < EM_DMAX * EM_SUB * 2 / 16; i += EM_THREADS) dst[i] = src[i];
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

    unsigned long long inert = 0;
    unsigned int zcount = 0;
    const int64_t n_chunks = (n + EM_PER - 1) / EM_PER;
    const int64_t total_bytes = (n * bits + 7) / 8;
    const int64_t stride = (int64_t)gridDim.x * EM_THREADS;
    const int64_t first = (int64_t)blockIdx.x * EM_THREADS + threadIdx.x;
    // complete chunks of an aligned tensor take the fast loop: two register sets in turn, the loads of the next chunk
    // (32 B per thread) issued before the current one is processed; what is left (a ragged last chunk, or everything for an
    // unaligned tensor) takes the generic loop below
    const int64_t n_full = VEC ? n / EM_PER : 0;
    auto load_full = [&](int64_t chunk, float *x) {
        const float4 a = ld_stream_f4(w + chunk * EM_PER), b = ld_stream_f4(w + chunk * EM_PER + 4);
        x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w;
        x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
    };
    // two copies of the loop: with and without the histogram bookkeeping (the fused compress path takes the code histogram
    // from the Lloyd kernel; the kernel is issue bound, every instruction per element counts)
    auto run = [&](auto hist_c) {
    constexpr bool HIST = decltype(hist_c)::value;
    auto body = [&](const float *x, int64_t chunk, auto full_c) {
        constexpr bool FULL = decltype(full_c)::value;
        const int64_t base = chunk * EM_PER;
        const int cnt = FULL ? EM_PER : (int)min((int64_t)EM_PER, n - base);
        // phase 1, branch free: LUT entry of every element (pruned weights take the known label of 0.0)
        int id[EM_PER];
        uint32_t any = 0;
#pragma unroll
        for (int j = 0; j < EM_PER; ++j) {
            const bool zero = x[j] == 0.f;
            const int f = lut_fine(fsub(x[j], mean), lut_lo, lut_scale);
            uint32_t e = S.lut[f / EM_SUB];
            if (e & EM_L2) e = S.lut2[(e & (EM_L2 - 1)) + (f % EM_SUB)];
            e = zero ? (uint32_t)zid : e;
            id[j] = (int)e;
            any |= e;
            if (HIST) zcount += zero;
        }
        uint32_t slow = 0;
        if (any & EM_SLOW) {  // which of the eight need the search (rare: their fine bucket holds a boundary or a zone)
#pragma unroll
            for (int j = 0; j < EM_PER; ++j) slow |= ((uint32_t)id[j] >> 15) << j;
        }
        // phase 2: the few elements whose fine bucket holds a region boundary or a zone (one pass per thread)
        while (slow) {
            const int j = __ffs(slow) - 1;
            slow &= slow - 1;
            float xj = x[0];
#pragma unroll
            for (int q = 1; q < EM_PER; ++q) xj = q == j ? x[q] : xj;
            int ej = id[0];
#pragma unroll
            for (int q = 1; q < EM_PER; ++q) ej = q == j ? id[q] : ej;
            const float xc = fsub(xj, mean);
            const int r = region_from(S.start, R, ej & (int)(EM_L2 - 1), xc);
            int lab = S.rid[r];
            if (lab < 0) lab = T.down[zone_argmin(xc, T.dv, T.dcn, T.down, T.rJ2[r], T.rJ1[r])];
#pragma unroll
            for (int q = 0; q < EM_PER; ++q) id[q] = q == j ? lab : id[q];
        }
        if (HIST) {
#pragma unroll
            for (int j = 0; j < EM_PER; ++j)
                if (x[j] != 0.f) atomicAdd(&S.hist[id[j]], 1u);  // NaN never gets here (rejected upstream)
        }
        if (INERTIA) {
#pragma unroll
            for (int j = 0; j < EM_PER; ++j) {
                if (j < cnt) {
                    // distance to the centroid in centred space, as sklearn's _inertia_dense computes it
                    const float t = fsub(fsub(x[j], mean), ed->cfin[id[j]]);
                    const float d2 = fmul(t, t);
                    inert += (unsigned long long)__double2ll_rn(__dmul_rn((double)d2, inertia_scale));
                }
            }
        }
        if (labels) {
            if (FULL) {
                reinterpret_cast<int4 *>(labels + base)[0] = make_int4(id[0], id[1], id[2], id[3]);
                reinterpret_cast<int4 *>(labels + base)[1] = make_int4(id[4], id[5], id[6], id[7]);
            } else {
                for (int j = 0; j < cnt; ++j) labels[base + j] = id[j];
            }
        }
        if (ris) {
            if (FULL) {
                st_stream_f4(ris + base, make_float4(S.val[id[0]], S.val[id[1]], S.val[id[2]], S.val[id[3]]));
                st_stream_f4(ris + base + 4, make_float4(S.val[id[4]], S.val[id[5]], S.val[id[6]], S.val[id[7]]));
            } else {
                for (int j = 0; j < cnt; ++j) ris[base + j] = S.val[id[j]];
            }
        }
        if (packed) {
            if (BITS == 8 && FULL) {
                const uint32_t lo = (uint32_t)id[0] | ((uint32_t)id[1] << 8) | ((uint32_t)id[2] << 16) | ((uint32_t)id[3] << 24);
                const uint32_t hi = (uint32_t)id[4] | ((uint32_t)id[5] << 8) | ((uint32_t)id[6] << 16) | ((uint32_t)id[7] << 24);
                *reinterpret_cast<uint2 *>(packed + base) = make_uint2(lo, hi);
            } else {
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
                const bool full = FULL || byte0 + bits <= total_bytes;
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
    };
    {
        float xa[EM_PER], xb[EM_PER];
        int64_t c = first;
        if (c < n_full) load_full(c, xa);
        while (c < n_full) {
            const int64_t c1 = c + stride;
            if (c1 < n_full) load_full(c1, xb);
            body(xa, c, std::true_type{});
            if (c1 >= n_full) break;
            c = c1 + stride;
            if (c < n_full) load_full(c, xa);
            body(xb, c1, std::true_type{});
        }
    }
    for (int64_t chunk = n_full + first; chunk < n_chunks; chunk += stride) {
        const int64_t base = chunk * EM_PER;
        const int cnt = (int)min((int64_t)EM_PER, n - base);
        float x[EM_PER];
#pragma unroll
        for (int j = 0; j < EM_PER; ++j) x[j] = j < cnt ? w[base + j] : 0.f;
        body(x, chunk, std::false_type{});
    }
    };
    if (want_hist)
        run(std::true_type{});
    else
        run(std::false_type{});
    if (INERTIA) {
        inert = warp_sum_ull(inert);
        if (lane_id() == 0 && inert) atomicAdd(&ed->inertia_q, inert);
    }
    if (want_hist) {
        // the thread that handled the last chunk counted its padding as zeros: take it back
        if ((n_chunks - 1 - n_full) % stride == first) zcount -= (unsigned int)(n_chunks * EM_PER - n);
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

template <bool VEC, bool INERTIA, int BITS>
static void emit_launch(nnc_ctx *ctx, int grid, const float *d_w, int64_t n, EmitDevice *ed, float mean, double inertia_scale,
                        int32_t *d_labels, float *d_ris, uint8_t *d_packed, int bits, int want_hist) {
    func_dyn_smem(ctx, (const void *)emit_kernel<VEC, INERTIA, BITS>, sizeof(EmitSmem));
    NNC_LAUNCH(ctx, (emit_kernel<VEC, INERTIA, BITS>), grid, EM_THREADS, sizeof(EmitSmem), d_w, n, ed, mean, inertia_scale,
               d_labels, d_ris, d_packed, bits, want_hist);
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
    float lut_scale = span > 0.f ? (float)((double)EM_FINE / (double)span * (1.0 - 1e-6)) : 0.f;
    if (!isfinite(lut_scale)) lut_scale = 0.f;
    func_dyn_smem(ctx, (const void *)emit_table_kernel, sizeof(EmitTabSmem));
    NNC_LAUNCH(ctx, emit_table_kernel, 1, TB_THREADS, sizeof(EmitTabSmem), ed, k, xabs, mean, xlo, lut_scale);
    const bool vec = ((reinterpret_cast<uintptr_t>(d_w) & 15u) == 0) &&
                     (!d_labels || (reinterpret_cast<uintptr_t>(d_labels) & 15u) == 0) &&
                     (!d_ris || (reinterpret_cast<uintptr_t>(d_ris) & 15u) == 0) &&
                     (!d_packed || (reinterpret_cast<uintptr_t>(d_packed) & 15u) == 0);
    const int64_t n_chunks = (n + EM_PER - 1) / EM_PER;
    const int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 2, (n_chunks + EM_THREADS - 1) / EM_THREADS);
    const int want_hist = h_hist ? 1 : 0;
    const bool b8 = bits == 8 && d_packed;
    if (vec && h_inertia)
        emit_launch<true, true, 0>(ctx, grid, d_w, n, ed, mean, inertia_scale, d_labels, d_ris, d_packed, bits, want_hist);
    else if (vec && b8)
        emit_launch<true, false, 8>(ctx, grid, d_w, n, ed, mean, inertia_scale, d_labels, d_ris, d_packed, bits, want_hist);
    else if (vec)
        emit_launch<true, false, 0>(ctx, grid, d_w, n, ed, mean, inertia_scale, d_labels, d_ris, d_packed, bits, want_hist);
    else if (h_inertia)
        emit_launch<false, true, 0>(ctx, grid, d_w, n, ed, mean, inertia_scale, d_labels, d_ris, d_packed, bits, want_hist);
    else
        emit_launch<false, false, 0>(ctx, grid, d_w, n, ed, mean, inertia_scale, d_labels, d_ris, d_packed, bits, want_hist);
    if (h_inertia) comm_allreduce(ctx, reinterpret_cast<int64_t *>(&ed->inertia_q), 1, 0);
    if (h_hist) comm_allreduce(ctx, reinterpret_cast<int64_t *>(ed->hist), k, 0);
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

// 0/1 bytes -> bits (bit i of byte i / 8 = byte i != 0): the 1-bit pruning mask of the compressed-layer format
// (common/storage.py).  A thread packs 16 mask bytes (one 128-bit load) into two output bytes.
__global__ void __launch_bounds__(256) pack_bits_kernel(const uint8_t *__restrict__ src, int64_t n, uint8_t *__restrict__ dst, int vec_ok) {
    const int64_t n16 = vec_ok ? n >> 4 : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = *reinterpret_cast<const uint4 *>(src + 16 * i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t out = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t b = w[j];
            b = (b | (b >> 1) | (b >> 2) | (b >> 3) | (b >> 4) | (b >> 5) | (b >> 6) | (b >> 7)) & 0x01010101u;  // byte != 0
            out |= (((b * 0x10204080u) >> 28) & 0xfu) << (4 * j);  // gather the four flag bits
        }
        *reinterpret_cast<uint16_t *>(dst + 2 * i) = (uint16_t)out;
    }
    // tail (and the whole array when it is not 16-byte aligned): one output byte per thread
    const int64_t b0 = n16 * 2, nb = (n + 7) >> 3;
    for (int64_t b = b0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
        uint32_t out = 0;
        for (int j = 0; j < 8; ++j) {
            const int64_t e = 8 * b + j;
            if (e < n && src[e]) out |= 1u << j;
        }
        dst[b] = (uint8_t)out;
    }
}

void pack_bits_device(nnc_ctx *ctx, const uint8_t *d_src, int64_t n, uint8_t *d_dst) {
    const int vec_ok = ((reinterpret_cast<uintptr_t>(d_src) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d_dst) & 1u) == 0);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * 16, (n / 16 + 255) / 256 + 1));
    NNC_LAUNCH(ctx, pack_bits_kernel, grid, 256, 0, d_src, n, d_dst, vec_ok);
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
