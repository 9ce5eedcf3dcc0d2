// reduce_np.cu -- float32 reductions that reproduce NumPy's pairwise summation bit for bit, with the
// pruning passes fused into them.
//
// Replaces np.std / |w| < thr / w[mask] = 0 of the reference's prune_weigth
// (neural_network_compression/common/utility.py:159-163) and X.mean(axis=0) of sklearn's KMeans.fit
// (sklearn/cluster/_kmeans.py:1486-1493).
//
// NumPy's float32 add.reduce is a fixed binary tree (numpy/_core/src/umath/loops_utils.h.src,
// @TYPE@_pairwise_sum): a node of n > 128 elements splits at n2 = (n/2) rounded down to a multiple of 8;
// a leaf (n <= 128) keeps 8 strided accumulators, combines them as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and
// then adds the n%8 tail sequentially.  Because the tree only depends on n, the GPU evaluates exactly the
// same tree: the top `depth` levels form a complete binary tree whose 2^depth subtrees ("tiles", <= 4096
// elements each) are reduced by one CTA each out of shared memory; a one-CTA kernel then folds the 2^depth
// partials pairwise.  Every addition is an explicit __fadd_rn, so the result equals NumPy's float32 value.
//
// Pruning is two such passes (13 B/weight instead of 17):
//   pass 1  tree-sum of x (-> mean), plus fp64 sum / sum of squares as an ESTIMATE of the std;
//   pass 2  tree-sum of fl((x-mean)^2) (-> NumPy's exact var/std/threshold) and, in the same read, the
//           speculative apply: elements whose |x| is outside a +-2^-16 relative band around the estimated
//           threshold are final (written as 0 / kept, mask byte emitted); the few elements inside the band
//           go to a side list and are decided by a fix-up kernel once the exact threshold is known.
#include <math.h>

#include <type_traits>

#include "common.cuh"
#include "internal.h"
#include "peer.cuh"
#include "table.cuh"

namespace nnc {

constexpr int NP_TILE_MAX = 4096;
constexpr int NP_THREADS = 256;

static inline int64_t half_down(int64_t s) {
    int64_t n2 = s / 2;
    return n2 - n2 % 8;
}

NpPlan np_plan(int64_t n) {
    int d = 0;
    int64_t s = n;
    while (s > NP_TILE_MAX) {  // follow the right (larger) children
        s = s - half_down(s);
        d++;
    }
    NpPlan p;
    p.depth = d;
    p.num_tiles = 1u << d;
    return p;
}
size_t np_partials_bytes(const NpPlan &p) { return sizeof(float) * 2 * (size_t)p.num_tiles; }

__host__ __device__ __forceinline__ void np_tile_root(int64_t n, int depth, uint32_t t, int64_t &off, int &sz) {
    int64_t o = 0, s = n;
    for (int lvl = depth - 1; lvl >= 0; --lvl) {
        int64_t n2 = s / 2;
        n2 -= n2 % 8;
        if ((t >> lvl) & 1u) {
            o += n2;
            s -= n2;
        } else {
            s = n2;
        }
    }
    off = o;
    sz = (int)s;
}

// ---------------------------------------------------------------------------------------------
// Visitors: what one pass does with each element besides producing the term that is tree-summed.
// ---------------------------------------------------------------------------------------------
struct BlockAux {  // shared scratch for visitor epilogues
    double d[2][NP_THREADS / 32];
    unsigned long long u[2][NP_THREADS / 32];
    uint32_t o[4][NP_THREADS / 32];
};

// pass 1 of pruning / nnc_stats: term = x; side: sum and sum of squares as an ESTIMATE of the std (float32 over
// the 16 elements a thread sees of a tile, float64 across tiles: relative error ~1e-7, the speculation band is
// 1.5e-5 wide).
struct VisitStats {
    static constexpr const char *kName = "np_tree_kernel<VisitStats>";
    static constexpr int kStageBufs = 3;  // staged tiles in flight + the one being worked on
    static constexpr bool kTileCount = false;
    static constexpr bool kSecondTree = false;
    static constexpr bool kCompact = false;
    DevScalars *sc;
    // The fp64 side sums feed the ESTIMATE of the std that places the speculation band of the pruning pass.  Their
    // float32 partials are taken of x - shift, shift = the mean of the first few elements: with the raw values a tensor with
    // |mean| >> std (normalisation scales, biases) lost (mean / std)^2 * 1e-7 of the variance and missed the 2^-16 band.
    // The shift is undone in float64 at the end of every tile (sum x, sum x^2 keep their meaning for the exchange).
    const float *probe = nullptr;  // this rank's elements
    int64_t probe_n = 0;
    float shift = 0.f;
    unsigned int cnt = 0;
    double s = 0.0, s2 = 0.0;
    float ts = 0.f, ts2 = 0.f;
    uint32_t amax = 0u;  // largest |x| bit pattern (the fused compress path takes its key range and NaN check from it)
    __device__ __forceinline__ void begin() {
        const int m = (int)(probe_n < 16 ? probe_n : 16);
        float a = 0.f;
        for (int i = 0; i < m; ++i) a += probe[i];  // (every thread: the same loads in the same order)
        a = m > 0 ? a / (float)m : 0.f;
        shift = (a - a == 0.f) ? a : 0.f;  // a non-finite probe: no shift
    }
    __device__ __forceinline__ float4 load4(const float *p) const { return ld_stream_f4(p); }
    __device__ __forceinline__ float load1(const float *p) const { return ld_stream_f1(p); }
    __device__ __forceinline__ float one(float x) {
        const float d = x - shift;
        ts += d;
        ts2 = fmaf(d, d, ts2);
        amax = max(amax, __float_as_uint(x) & 0x7fffffffu);
        return x;
    }
    __device__ __forceinline__ float term(float x) const { return x; }
    __device__ __forceinline__ float4 visit4(int64_t, const float *, float4 x) {
        cnt += 4;
        return make_float4(one(x.x), one(x.y), one(x.z), one(x.w));
    }
    __device__ __forceinline__ float visit1(int64_t, const float *, float x) {
        cnt += 1;
        return one(x);
    }
    __device__ __forceinline__ void end_tile() {
        const double sh = (double)shift, c = (double)cnt, t = (double)ts;
        s += t + c * sh;
        s2 += (double)ts2 + 2.0 * sh * t + c * sh * sh;
        ts = 0.f;
        ts2 = 0.f;
        cnt = 0;
    }
    __device__ void finish(BlockAux &aux) {
        double a = warp_sum_d(s), b = warp_sum_d(s2);
        const uint32_t m = warp_max_u(amax);
        if (lane_id() == 0) {
            aux.d[0][warp_id()] = a;
            aux.d[1][warp_id()] = b;
            aux.o[0][warp_id()] = m;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ta = 0, tb = 0;
            uint32_t tm = 0;
            for (int i = 0; i < NP_THREADS / 32; i++) {
                ta += aux.d[0][i];
                tb += aux.d[1][i];
                tm = max(tm, aux.o[0][i]);
            }
            atomicAdd(&sc->sum_d, ta);
            atomicAdd(&sc->sumsq_d, tb);
            atomicMax(&sc->amax_all, tm);
        }
    }
};

// plain centred squares (nnc_stats second pass, non-speculative prune fallback)
struct VisitCenSq {
    static constexpr const char *kName = "np_tree_kernel<VisitCenSq>";
    static constexpr int kStageBufs = 3;  // staged tiles in flight + the one being worked on
    static constexpr bool kTileCount = false;
    static constexpr bool kSecondTree = false;
    static constexpr bool kCompact = false;
    DevScalars *sc;
    float mean;
    __device__ __forceinline__ void begin() { mean = sc->mean; }
    __device__ __forceinline__ float4 load4(const float *p) const { return ld_stream_f4(p); }
    __device__ __forceinline__ float load1(const float *p) const { return ld_stream_f1(p); }
    __device__ __forceinline__ float one(float x) const {
        float t = fsub(x, mean);
        return fmul(t, t);
    }
    __device__ __forceinline__ float term(float x) const { return one(x); }
    __device__ __forceinline__ float4 visit4(int64_t, const float *, float4 x) {
        return make_float4(one(x.x), one(x.y), one(x.z), one(x.w));
    }
    __device__ __forceinline__ float visit1(int64_t, const float *, float x) { return one(x); }
    __device__ __forceinline__ void end_tile() {}
    __device__ void finish(BlockAux &) {}
};

// pass 2 of pruning: centred squares + speculative apply (see file header).
struct VisitCenSqApply {
    static constexpr const char *kName = "np_tree_kernel<VisitCenSqApply>";
    static constexpr int kStageBufs = 2;  // staged tiles in flight + the one being worked on
    static constexpr bool kTileCount = false;
    static constexpr bool kSecondTree = false;
    static constexpr bool kCompact = false;
    DevScalars *sc;
    float *w;            // in place
    uint8_t *mask;
    long long *side_idx;
    float *side_val;
    unsigned long long side_cap;
    int mask_vec_ok;     // mask base 4-byte aligned
    float mean, lo, hi;
    unsigned long long pruned = 0;
    __device__ __forceinline__ void begin() {
        mean = sc->mean;
        lo = (float)sc->band_lo;  // already fp32-exact thresholds (see prune_finalize1)
        hi = (float)sc->band_hi;
    }
    // the kernel writes the memory it reads: no non-coherent path here
    __device__ __forceinline__ float4 load4(const float *p) const { return *reinterpret_cast<const float4 *>(p); }
    __device__ __forceinline__ float load1(const float *p) const { return *p; }
    __device__ __forceinline__ float term(float x) const {
        const float t = fsub(x, mean);
        return fmul(t, t);
    }
    __device__ __forceinline__ float one(int64_t g, float x, float &outv, uint32_t &m) {
        float t = fsub(x, mean);
        float a = fabsf(x);
        if (a < lo) {
            outv = 0.f;
            m = 1;
            pruned++;
        } else {
            outv = x;
            m = 0;
            if (a < hi) {  // inside the band: decided later with the exact threshold
                unsigned long long slot = atomicAdd(&sc->band_count, 1ull);
                if (slot < side_cap) {
                    side_idx[slot] = g;
                    side_val[slot] = x;
                } else {
                    atomicAdd(&sc->band_dropped, 1ull);
                }
            }
        }
        return fmul(t, t);
    }
    __device__ __forceinline__ float4 visit4(int64_t g, const float *p, float4 x) {
        float4 o, r;
        uint32_t m0, m1, m2, m3;
        r.x = one(g, x.x, o.x, m0);
        r.y = one(g + 1, x.y, o.y, m1);
        r.z = one(g + 2, x.z, o.z, m2);
        r.w = one(g + 3, x.w, o.w, m3);
        if (m0 | m1 | m2 | m3) *reinterpret_cast<float4 *>(w + g) = o;
        uint32_t mm = m0 | (m1 << 8) | (m2 << 16) | (m3 << 24);
        if (mask_vec_ok) {
            *reinterpret_cast<uint32_t *>(mask + g) = mm;
        } else {
            mask[g] = (uint8_t)m0;
            mask[g + 1] = (uint8_t)m1;
            mask[g + 2] = (uint8_t)m2;
            mask[g + 3] = (uint8_t)m3;
        }
        (void)p;
        return r;
    }
    __device__ __forceinline__ float visit1(int64_t g, const float *, float x) {
        float o;
        uint32_t m;
        float r = one(g, x, o, m);
        if (m) w[g] = o;
        mask[g] = (uint8_t)m;
        return r;
    }
    __device__ __forceinline__ void end_tile() {}
    __device__ void finish(BlockAux &aux) {
        unsigned long long c = warp_sum_ull(pruned);
        if (lane_id() == 0) aux.u[0][warp_id()] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int i = 0; i < NP_THREADS / 32; i++) t += aux.u[0][i];
            if (t) atomicAdd(&sc->n_pruned, t);
        }
    }
};

// k-means prologue: term = x (-> mean); side: min / max, non-zero count, range of |x| bit patterns over the
// non-zero elements (the radix-sort key range), non-finite detection (|x| bits >= 0x7f800000).
struct VisitQuant {
    static constexpr const char *kName = "np_tree_kernel<VisitQuant>";
    static constexpr int kStageBufs = 2;  // staged tiles in flight + the one being worked on
    static constexpr bool kTileCount = false;
    static constexpr bool kSecondTree = false;
    static constexpr bool kCompact = true;  // the kernel also writes the non-zero elements of every tile to `out`
    DevScalars *sc;
    float *out;                       // survivors, dense, in no particular tile order (they are sorted next)
    unsigned long long *cursor;       // next free slot of `out`
    unsigned long long capacity;
    float mn = INFINITY, mx = -INFINITY;
    uint32_t amax = 0u, amin_m1 = 0xffffffffu;
    unsigned int nz = 0, nz_before = 0;
    __device__ __forceinline__ unsigned int take_tile_count() {
        const unsigned int c = nz - nz_before;
        nz_before = nz;
        return c;
    }
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ float4 load4(const float *p) const { return ld_stream_f4(p); }
    __device__ __forceinline__ float load1(const float *p) const { return ld_stream_f1(p); }
    __device__ __forceinline__ float term(float x) const { return x; }
    __device__ __forceinline__ float one(float x) {
        const uint32_t a = __float_as_uint(x) & 0x7fffffffu;
        amax = max(amax, a);
        amin_m1 = min(amin_m1, a - 1u);  // zero wraps to 0xffffffff and never wins
        nz += a != 0u;
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
        return x;
    }
    __device__ __forceinline__ float4 visit4(int64_t, const float *, float4 x) {
        return make_float4(one(x.x), one(x.y), one(x.z), one(x.w));
    }
    __device__ __forceinline__ float visit1(int64_t, const float *, float x) { return one(x); }
    __device__ __forceinline__ void end_tile() {}
    __device__ void finish(BlockAux &aux) {
        // a thread sees at most 2^32 / 4 elements only for absurd grids; the per-thread count fits 32 bits
        float a = warp_min_f(mn), b = warp_max_f(mx);
        uint32_t c = warp_max_u(amax), d = warp_min_u(amin_m1);
        unsigned long long e = warp_sum_ull((unsigned long long)nz);
        if (lane_id() == 0) {
            aux.o[0][warp_id()] = __float_as_uint(a);
            aux.o[1][warp_id()] = __float_as_uint(b);
            aux.o[2][warp_id()] = c;
            aux.o[3][warp_id()] = d;
            aux.u[0][warp_id()] = e;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < NP_THREADS / 32; i++) {
                a = fminf(a, __uint_as_float(aux.o[0][i]));
                b = fmaxf(b, __uint_as_float(aux.o[1][i]));
                c = max(c, aux.o[2][i]);
                d = min(d, aux.o[3][i]);
                e += aux.u[0][i];
            }
            // -0.0 -> +0.0 so that the ordered image of the extrema is canonical
            if (a == 0.f) a = 0.f;
            if (b == 0.f) b = 0.f;
            if (a <= b) {  // false only if every element seen was NaN
                atomicMin(&sc->min_ord, f2ord(a));
                atomicMax(&sc->max_ord, f2ord(b));
            }
            atomicMax(&sc->amax_bits, c);
            atomicMin(&sc->amin_nz_m1, d);
            if (e) atomicAdd(&sc->n_nz, e);
            if (c >= 0x7f800000u) atomicAdd(&sc->n_nonfinite, 1ull);
        }
    }
};

// plain term = x (re-reduction of single tiles, see VisitApplyQuant)
struct VisitPlain {
    static constexpr const char *kName = "np_tree_kernel<VisitPlain>";
    static constexpr int kStageBufs = 3;  // staged tiles in flight + the one being worked on
    static constexpr bool kTileCount = false;
    static constexpr bool kSecondTree = false;
    static constexpr bool kCompact = false;
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ float4 load4(const float *p) const { return *reinterpret_cast<const float4 *>(p); }
    __device__ __forceinline__ float load1(const float *p) const { return *p; }
    __device__ __forceinline__ float term(float x) const { return x; }
    __device__ __forceinline__ float4 visit4(int64_t, const float *, float4 x) { return x; }
    __device__ __forceinline__ float visit1(int64_t, const float *, float x) { return x; }
    __device__ __forceinline__ void end_tile() {}
    __device__ void finish(BlockAux &) {}
};

static __device__ __noinline__ void band_append(DevScalars *sc, long long *side_idx, float *side_val, unsigned long long side_cap,
                                                long long g, float x) {
    unsigned long long slot = atomicAdd(&sc->band_count, 1ull);
    if (slot < side_cap) {
        side_idx[slot] = g;
        side_val[slot] = x;
    } else {
        atomicAdd(&sc->band_dropped, 1ull);
    }
}

// pass 2 of pruning FUSED with the k-means prologue (nnc_compress_f32): besides VisitCenSqApply's work the same read
// produces what VisitQuant would get from another sweep over the PRUNED tensor --
//   second tree   NumPy's pairwise sum of the pruned values (-> X.mean of KMeans.fit, _kmeans.py:1486)
//   compaction    the survivors, written densely to `out`
//   scalars       min / max, |x| key range and count of the survivors.
// Elements inside the speculation band are not decided yet: they are left out of the compaction and the scalars
// (prune_fixup_kernel adds the ones that survive) and their tile is put on a list instead of delivering its
// second-tree partial; the few listed tiles are re-reduced from the final tensor afterwards (VisitPlain).
struct VisitApplyQuant {
    static constexpr const char *kName = "np_tree_kernel<VisitApplyQuant>";
    static constexpr int kStageBufs = 2;  // staged tiles in flight + the one being worked on
    static constexpr bool kTileCount = false;
    static constexpr bool kSecondTree = true;
    static constexpr bool kCompact = true;
    DevScalars *sc;
    float *w;  // in place
    uint8_t *mask;
    long long *side_idx;
    float *side_val;
    unsigned long long side_cap;
    int mask_vec_ok;
    float *out;  // survivors
    unsigned long long *cursor;
    unsigned long long capacity;
    float *partials2;           // second tree: one partial per tile
    uint32_t *dirty_list;       // tiles whose partial has to be recomputed
    unsigned int *dirty_count;
    float mean, lo, hi;
    unsigned int pruned = 0;  // per thread: far below 2^32
    float mn = INFINITY, mx = -INFINITY;
    bool dirty = false;
    __device__ __forceinline__ void begin() {
        mean = sc->mean;
        lo = (float)sc->band_lo;
        hi = (float)sc->band_hi;
    }
    __device__ __forceinline__ float4 load4(const float *p) const { return *reinterpret_cast<const float4 *>(p); }
    __device__ __forceinline__ float load1(const float *p) const { return *p; }
    __device__ __forceinline__ bool take_dirty() {
        const bool d = dirty;
        dirty = false;
        return d;
    }
    // tree 1: centred squares; tree 2: the pruned value.  (An element inside the band counts as kept here: its tile is on
    // the dirty list and its second-tree partial is recomputed from the final tensor, so the value does not matter.)
    __device__ __forceinline__ float term(float x) const {
        const float t = fsub(x, mean);
        return fmul(t, t);
    }
    __device__ __forceinline__ float term2(float x) const { return fabsf(x) < lo ? 0.f : x; }
    // an element inside the band: decided later with the exact threshold (rare: ~1e-5 of the elements).  A free
    // function on purpose: a non-inlined MEMBER would take `this` and push the whole visitor into local memory.
    __device__ __forceinline__ void band(int64_t g, float x) {
        dirty = true;
        band_append(sc, side_idx, side_val, side_cap, g, x);
    }
    __device__ __forceinline__ float4 visit4(int64_t g, const float *, float4 x, float4 &second) {
        const float a0 = fabsf(x.x), a1 = fabsf(x.y), a2 = fabsf(x.z), a3 = fabsf(x.w);
        const bool p0 = a0 < lo, p1 = a1 < lo, p2 = a2 < lo, p3 = a3 < lo;
        float4 o = make_float4(p0 ? 0.f : x.x, p1 ? 0.f : x.y, p2 ? 0.f : x.z, p3 ? 0.f : x.w);  // the pruned tensor, band kept
        second = o;
        // in band: lo <= |x| < hi (a NaN is in neither set: it stays, final)
        const bool b0 = !p0 && a0 < hi, b1 = !p1 && a1 < hi, b2 = !p2 && a2 < hi, b3 = !p3 && a3 < hi;
        float4 sv = o;  // values that join min / max (fminf / fmaxf skip the NaN standing for an undecided element)
        if (b0 | b1 | b2 | b3) {
            if (b0) {
                band(g, x.x);
                second.x = 0.f;
                sv.x = NAN;
            }
            if (b1) {
                band(g + 1, x.y);
                second.y = 0.f;
                sv.y = NAN;
            }
            if (b2) {
                band(g + 2, x.z);
                second.z = 0.f;
                sv.z = NAN;
            }
            if (b3) {
                band(g + 3, x.w);
                second.w = 0.f;
                sv.w = NAN;
            }
        }
        mn = fminf(fminf(mn, sv.x), fminf(sv.y, fminf(sv.z, sv.w)));
        mx = fmaxf(fmaxf(mx, sv.x), fmaxf(sv.y, fmaxf(sv.z, sv.w)));
        const uint32_t mm = (uint32_t)p0 | ((uint32_t)p1 << 8) | ((uint32_t)p2 << 16) | ((uint32_t)p3 << 24);
        pruned += (unsigned int)p0 + (unsigned int)p1 + (unsigned int)p2 + (unsigned int)p3;
        if (mm) *reinterpret_cast<float4 *>(w + g) = o;
        if (mask_vec_ok) {
            *reinterpret_cast<uint32_t *>(mask + g) = mm;
        } else {
            mask[g] = (uint8_t)p0;
            mask[g + 1] = (uint8_t)p1;
            mask[g + 2] = (uint8_t)p2;
            mask[g + 3] = (uint8_t)p3;
        }
        const float t0 = fsub(x.x, mean), t1 = fsub(x.y, mean), t2 = fsub(x.z, mean), t3 = fsub(x.w, mean);
        return make_float4(fmul(t0, t0), fmul(t1, t1), fmul(t2, t2), fmul(t3, t3));
    }
    __device__ __forceinline__ float visit1(int64_t g, const float *, float x, float &second) {
        const float a = fabsf(x);
        const bool p = a < lo;
        const float o = p ? 0.f : x;
        second = o;
        if (!p && a < hi) {
            band(g, x);
            second = 0.f;
        } else {
            mn = fminf(mn, o);
            mx = fmaxf(mx, o);
        }
        if (p) {
            w[g] = 0.f;
            pruned++;
        }
        mask[g] = (uint8_t)p;
        const float t = fsub(x, mean);
        return fmul(t, t);
    }
    __device__ __forceinline__ void end_tile() {}
    __device__ void finish(BlockAux &aux) {
        float a = warp_min_f(mn), b = warp_max_f(mx);
        unsigned long long pc = warp_sum_ull((unsigned long long)pruned);
        if (lane_id() == 0) {
            aux.o[0][warp_id()] = __float_as_uint(a);
            aux.o[1][warp_id()] = __float_as_uint(b);
            aux.u[1][warp_id()] = pc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long pt = pc;
            for (int i = 1; i < NP_THREADS / 32; i++) {
                a = fminf(a, __uint_as_float(aux.o[0][i]));
                b = fmaxf(b, __uint_as_float(aux.o[1][i]));
                pt += aux.u[1][i];
            }
            if (a == 0.f) a = 0.f;  // -0.0 -> +0.0: canonical ordered image
            if (b == 0.f) b = 0.f;
            if (a <= b) {
                atomicMin(&sc->min_ord, f2ord(a));
                atomicMax(&sc->max_ord, f2ord(b));
            }
            if (pt) atomicAdd(&sc->n_pruned, pt);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// The tree kernel: one tile (= one depth-`depth` subtree of NumPy's recursion, 2041..4096 elements) per loop
// iteration.
//   stage   the RAW tile arrives in shared memory by ONE bulk-asynchronous copy (cp.async.bulk + mbarrier), double
//           buffered: the copy of tile i+1 is in flight while tile i is visited, summed and folded.  The visitor reads
//           the staged values for its side effects (apply, mask, statistics, compaction)
//   leaves  8 lanes per depth-5 node of the tile's sub-tree: NumPy's 8 strided accumulators per leaf (a node
//           that is still > 128 elements splits once more into two leaves); the tree TERM of an element (x, (x-mean)^2,
//           the pruned value) is computed from the staged raw value on the fly
//   fold    warp 0 folds the <= 32 node values up the sub-tree while the other warps stage the next tile
// ---------------------------------------------------------------------------------------------
// sum of the leaf [o, o + s) of the staged tile in NumPy's order; called by the 8 lanes j = 0..7 of a group,
// every lane returns the leaf's value.  s <= 128, o a multiple of 8.  f: the tree term of a raw element.
template <class F>
__device__ __forceinline__ float np_leaf_sum(const float *tile, int o, int s, int j, F f) {
    float r = 0.f;
    const int rows = s >> 3;  // full rows of 8 consecutive elements
    const float *pA = tile + o + j;
    if (rows > 0) {
        r = f(pA[0]);
#pragma unroll
        for (int i = 1; i < 16; ++i)
            if (i < rows) r = fadd(r, f(pA[8 * i]));
    }
    // the 8 lanes of the group only: groups of one warp may be in different branches
    const unsigned gmask = 0xffu << (threadIdx.x & 24);
    r = fadd(r, __shfl_xor_sync(gmask, r, 1));
    r = fadd(r, __shfl_xor_sync(gmask, r, 2));
    r = fadd(r, __shfl_xor_sync(gmask, r, 4));
    if (rows > 0) {
        for (int i = rows << 3; i < s; ++i) r = fadd(r, f(tile[o + i]));
    } else {  // n < 8: plain sequential sum starting from 0
        r = 0.f;
        for (int i = 0; i < s; ++i) r = fadd(r, f(tile[o + i]));
    }
    return r;
}

// Per-tile description of the tree (it only depends on n), filled once per reduction by one thread per tile:
// offset and size of the tile, for each of the 32 depth-5 path groups the node it sums
//   desc = o | s << 13 | h << 21 | mine << 27     (o, s: offset / size inside the tile; h: heap slot of the value;
//                                                   mine: this group is the one that stores it)
// and the bit mask of the internal heap nodes h < 32 (size > 128: value = left + right).
struct NpTileDesc {
    long long off;
    int sz;
    uint32_t internal;
    uint32_t grp[32];
};

__global__ void np_tiles_kernel(int64_t n, int depth, NpTileDesc *desc) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (1u << depth)) return;
    int64_t off;
    int sz;
    np_tile_root(n, depth, t, off, sz);
    NpTileDesc d;
    d.off = off;
    d.sz = sz;
    d.internal = 0;
    for (int g = 0; g < 32; ++g) {
        int o = 0, s = sz, h = 1;
        bool mine = true, leaf = false;
        for (int lvl = 4; lvl >= 0; --lvl) {
            if (!leaf && s > 128) {
                d.internal |= 1u << h;
                int n2 = s >> 1;
                n2 -= n2 & 7;
                if ((g >> lvl) & 1) {
                    o += n2;
                    s -= n2;
                    h = 2 * h + 1;
                } else {
                    s = n2;
                    h = 2 * h;
                }
            } else {
                leaf = true;  // h stays the leaf's own heap slot; it belongs to its leftmost descendant group
                if ((g >> lvl) & 1) mine = false;
            }
        }
        d.grp[g] = (uint32_t)o | ((uint32_t)s << 13) | ((uint32_t)h << 21) | ((mine ? 1u : 0u) << 27);
    }
    desc[t] = d;
}

// value of the node `gd` of the staged tile `tl` (NumPy order); called by the 8 lanes j = 0..7 of a group.
// The tile lies in shared memory as in global memory (no padding).  A full leaf is 16 rows of 8 floats, lane j adds row
// after row into accumulator j; the four groups of a warp work on leaves 128 floats apart, i.e. on the SAME eight banks
// in every row.  They are therefore skewed by one row each: in step s group g reads row s - g (19 steps instead of 16), so
// that the four groups always touch four different bank octets -- conflict free without padding, which lets the tile
// arrive by a single bulk copy.  The additions of an accumulator happen in row order as in NumPy.
template <class F>
__device__ __forceinline__ float np_node_value(const float *tl, uint32_t gd, int j, F f) {
    const int o = gd & 8191, s = (gd >> 13) & 255;
    float val;
    if (s == 128 && (o & 127) == 0) {  // the common case: a full, aligned leaf
        const int g = (threadIdx.x >> 3) & 3;
        const float *p = tl + o + j - 8 * g;
        // row i = st - g: steps 0..2 start the groups one after the other, steps 4..15 are unconditional for every group
        // (1 <= i <= 15), steps 16..18 let the later groups finish
        val = 0.f;
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            if (st >= g) {
                const float x = f(p[8 * st]);
                val = st == g ? x : fadd(val, x);
            }
        }
#pragma unroll
        for (int st = 4; st < 16; ++st) val = fadd(val, f(p[8 * st]));
#pragma unroll
        for (int st = 16; st < 19; ++st)
            if (st - g < 16) val = fadd(val, f(p[8 * st]));
        const unsigned gmask = 0xffu << (threadIdx.x & 24);
        val = fadd(val, __shfl_xor_sync(gmask, val, 1));
        val = fadd(val, __shfl_xor_sync(gmask, val, 2));
        val = fadd(val, __shfl_xor_sync(gmask, val, 4));
    } else if (s > 128) {  // a depth-5 node that splits once more: two leaves
        int n2 = s >> 1;
        n2 -= n2 & 7;
        const float l = np_leaf_sum(tl, o, n2, j, f);
        const float r = np_leaf_sum(tl, o + n2, s - n2, j, f);
        val = fadd(l, r);
    } else {
        val = np_leaf_sum(tl, o, s, j, f);
    }
    return val;
}

// both trees of the fused pass from ONE read of the staged values (full leaves; other shapes take the two single passes)
template <class F1, class F2>
__device__ __forceinline__ void np_node_value2(const float *tl, uint32_t gd, int j, F1 f1, F2 f2, float &v1, float &v2) {
    const int o = gd & 8191, s = (gd >> 13) & 255;
    if (s == 128 && (o & 127) == 0) {
        const int g = (threadIdx.x >> 3) & 3;
        const float *p = tl + o + j - 8 * g;
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            if (st >= g) {
                const float x = p[8 * st];
                const float x1 = f1(x), x2 = f2(x);
                a = st == g ? x1 : fadd(a, x1);
                b = st == g ? x2 : fadd(b, x2);
            }
        }
#pragma unroll
        for (int st = 4; st < 16; ++st) {
            const float x = p[8 * st];
            a = fadd(a, f1(x));
            b = fadd(b, f2(x));
        }
#pragma unroll
        for (int st = 16; st < 19; ++st) {
            if (st - g < 16) {
                const float x = p[8 * st];
                a = fadd(a, f1(x));
                b = fadd(b, f2(x));
            }
        }
        const unsigned gmask = 0xffu << (threadIdx.x & 24);
        a = fadd(a, __shfl_xor_sync(gmask, a, 1));
        b = fadd(b, __shfl_xor_sync(gmask, b, 1));
        a = fadd(a, __shfl_xor_sync(gmask, a, 2));
        b = fadd(b, __shfl_xor_sync(gmask, b, 2));
        a = fadd(a, __shfl_xor_sync(gmask, a, 4));
        b = fadd(b, __shfl_xor_sync(gmask, b, 4));
        v1 = a;
        v2 = b;
    } else {
        v1 = np_node_value(tl, gd, j, f1);
        v2 = np_node_value(tl, gd, j, f2);
    }
}

// ---- bulk-asynchronous copy + mbarrier (PTX; SASS: UBLKCP / SYNCS) --------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    const uint32_t addr = smem_addr(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(addr), "r"(parity)
                     : "memory");
    } while (!done);
}

// dynamic shared memory of np_tree_kernel<V>: float raw[2][NP_TILE_MAX] (the double-buffered staged tile, as it lies in global memory), then for
// kCompact float s_out[2][NP_TILE_MAX]
template <class V>
constexpr size_t np_tree_smem() {
    return sizeof(float) * ((size_t)V::kStageBufs * NP_TILE_MAX + (V::kCompact ? 2 * (size_t)NP_TILE_MAX : 0));
}

// a: this rank's shard (elements [shard_begin, ...) of the flattened tensor); tiles [t0, t1) belong to it -- or, with a
// tile list, the tiles tile_list[0, *list_count).
template <class V>
__global__ void __launch_bounds__(NP_THREADS, V::kSecondTree ? 3 : 4) np_tree_kernel(const float *a, uint32_t t0, uint32_t t1, int64_t shard_begin,
                                                             int vec_ok, const NpTileDesc *__restrict__ desc, float *partials,
                                                             const uint32_t *__restrict__ tile_list, const unsigned int *list_count,
                                                             V v) {
    __shared__ float heap_val[2][64];
    __shared__ float heap_val2[V::kSecondTree ? 2 : 1][64];
    __shared__ unsigned int s_cnt[2];   // survivors staged so far in s_out[b]
    __shared__ unsigned int s_fcnt[2];  // ... of the completed tile in s_out[b], waiting to be flushed to s_base[b]
    __shared__ unsigned long long s_base[2];
    __shared__ int s_dirty[2];
    constexpr int NBUF = V::kStageBufs;
    __shared__ __align__(8) unsigned long long mbar[NBUF];  // "tile staged" of raw[b]
    extern __shared__ __align__(16) unsigned char np_dyn_smem[];
    float(*raw)[NP_TILE_MAX] = reinterpret_cast<float(*)[NP_TILE_MAX]>(np_dyn_smem);
    float(*s_out)[NP_TILE_MAX] = reinterpret_cast<float(*)[NP_TILE_MAX]>(np_dyn_smem + NBUF * sizeof(float) * NP_TILE_MAX);
    __shared__ BlockAux aux;

    v.begin();
    if (threadIdx.x < 2) {
        s_cnt[threadIdx.x] = 0;
        s_fcnt[threadIdx.x] = 0;
        s_base[threadIdx.x] = 0;
        s_dirty[threadIdx.x] = 0;
    }
    if (threadIdx.x < NBUF) mbar_init(&mbar[threadIdx.x], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int grp = threadIdx.x >> 3, j = threadIdx.x & 7;
    const uint32_t n_mine = tile_list ? *list_count : t1 - t0;
    auto tile_id = [&](uint32_t it) { return tile_list ? tile_list[it] : t0 + it; };
    // a tile travels by bulk copies when the tensor is 16-byte aligned and the tile is a whole number of 16-byte pieces
    // (tile offsets are multiples of 8 elements; only the last tile of a tensor can be ragged)
    auto bulk_ok = [&](int sz) { return vec_ok && (sz & 3) == 0; };
    // ONE bulk copy per tile (<= 16 KB; the copy engine's rate is per operation: 32 copies of 512 B per tile, landing in a
    // padded layout, ran 35 % slower than ordinary loads), issued by one thread into raw[b]
    auto issue = [&](uint32_t it, int b) {
        const uint32_t t = tile_id(it);
        const int sz = desc[t].sz;
        if (!bulk_ok(sz) || lane_id() != 0) return;
        const float *src = a + (desc[t].off - shard_begin);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer's previous readers were ordinary loads
        mbar_expect_tx(&mbar[b], (uint32_t)sz * 4u);
        bulk_copy_g2s(&raw[b][0], src, (uint32_t)sz * 4u, &mbar[b]);
    };
    // NBUF - 1 tiles ahead: the copies of the next tiles are in flight while this one is visited, summed and folded
    if (warp_id() == 1)
        for (int a0 = 0; a0 < NBUF - 1; ++a0)
            if (blockIdx.x + (uint32_t)a0 * gridDim.x < n_mine) issue(blockIdx.x + (uint32_t)a0 * gridDim.x, a0);
    int buf = 0;              // parity of the tile: node values, survivor stage, dirty flag
    int rb = 0;               // staging buffer of the tile
    uint32_t phase_bits = 0;  // bit b: parity of the next completed bulk staging of raw[b]
    for (uint32_t it = blockIdx.x; it < n_mine; it += gridDim.x, buf ^= 1, rb = rb + 1 == NBUF ? 0 : rb + 1) {
        const uint32_t t = tile_id(it);
        const int64_t off = desc[t].off - shard_begin;  // offset inside the shard
        const int sz = desc[t].sz;
        const uint32_t gd = desc[t].grp[grp];
        const float *src = a + off;
        // ---- the next tile's copies go out first (its buffer was last read before the barrier that ended the previous
        // iteration), then wait for this tile
        {
            const uint32_t ahead = it + (uint32_t)(NBUF - 1) * gridDim.x;  // goes into the buffer the previous tile used
            if (ahead < n_mine && warp_id() == 1) issue(ahead, rb == 0 ? NBUF - 1 : rb - 1);
        }
        float *tl = raw[rb];
        if (bulk_ok(sz)) {
            mbar_wait(&mbar[rb], (phase_bits >> rb) & 1u);
            phase_bits ^= 1u << rb;
        } else {  // unaligned tensor or the ragged last tile: ordinary loads into the same padded layout
            for (int i = threadIdx.x; i < sz; i += NP_THREADS) tl[i] = v.load1(src + i);
            __syncthreads();
        }
        // ---- visit: the visitor's side effects from the staged values (apply / mask / statistics / compaction)
        if (vec_ok) {
            const int nvec = sz >> 2;
            constexpr int NV = NP_TILE_MAX / 4 / NP_THREADS;
            constexpr int RSTRIDE = 4 * NP_THREADS;
            const int sidx0 = 4 * (int)threadIdx.x;
            if constexpr (V::kSecondTree) {
                float4 y2[NV];
                int cnt = 0;
                auto stage = [&](auto full_c) {
                    constexpr bool FULL = decltype(full_c)::value;  // a complete 4096-element tile: no bounds checks
#pragma unroll
                    for (int r = 0; r < NV; ++r) {
                        const int i = r * NP_THREADS + threadIdx.x;
                        y2[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (FULL || i < nvec) {
                            const float4 x = *reinterpret_cast<const float4 *>(&tl[sidx0 + r * RSTRIDE]);
                            (void)v.visit4(off + 4 * i, src + 4 * i, x, y2[r]);
                        }
                        cnt += (y2[r].x != 0.f) + (y2[r].y != 0.f) + (y2[r].z != 0.f) + (y2[r].w != 0.f);
                    }
                };
                if (sz == NP_TILE_MAX)
                    stage(std::true_type{});
                else
                    stage(std::false_type{});
                // survivors straight from the registers into the shared-memory stage s_out[buf] (any order: they are
                // sorted next): one warp scan and one shared-memory atomic per warp and tile
                const int lane = lane_id();
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int tt = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += tt;
                }
                unsigned int wb = 0;
                if (lane == 31 && incl) wb = atomicAdd(&s_cnt[buf], (unsigned int)incl);
                wb = __shfl_sync(0xffffffffu, wb, 31);
                float *dst = s_out[buf] + wb + (incl - cnt);
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    if (y2[r].x != 0.f) *dst++ = y2[r].x;
                    if (y2[r].y != 0.f) *dst++ = y2[r].y;
                    if (y2[r].z != 0.f) *dst++ = y2[r].z;
                    if (y2[r].w != 0.f) *dst++ = y2[r].w;
                }
                for (int i = (nvec << 2) + threadIdx.x; i < sz; i += NP_THREADS) {  // ragged end of the last tile
                    float y2s;
                    (void)v.visit1(off + i, src + i, tl[i], y2s);
                    if (y2s != 0.f) s_out[buf][atomicAdd(&s_cnt[buf], 1u)] = y2s;
                }
            } else {
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    const int i = r * NP_THREADS + threadIdx.x;
                    if (i < nvec) {
                        const float4 x = *reinterpret_cast<const float4 *>(&tl[sidx0 + r * RSTRIDE]);
                        (void)v.visit4(off + 4 * i, src + 4 * i, x);
                    }
                }
                for (int i = (nvec << 2) + threadIdx.x; i < sz; i += NP_THREADS) (void)v.visit1(off + i, src + i, tl[i]);
            }
        } else {
            for (int i = threadIdx.x; i < sz; i += NP_THREADS) {
                if constexpr (V::kSecondTree) {
                    float y2s;
                    (void)v.visit1(off + i, src + i, tl[i], y2s);
                    if (y2s != 0.f) s_out[buf][atomicAdd(&s_cnt[buf], 1u)] = y2s;
                } else {
                    (void)v.visit1(off + i, src + i, tl[i]);
                }
            }
        }
        v.end_tile();
        if constexpr (V::kTileCount) {
            const unsigned int c = (unsigned int)warp_sum_i((int)v.take_tile_count());
            if (lane_id() == 0 && c) atomicAdd(&s_cnt[buf], c);
        }
        if constexpr (V::kSecondTree) {
            if (v.take_dirty()) s_dirty[buf] = 1;
        }
        // ---- leaves: the 8 lanes of group `grp` sum the node described by gd, the terms computed from the staged values
        {
            const int h = (gd >> 21) & 63;
            if constexpr (V::kSecondTree) {
                float val, val2;
                np_node_value2(tl, gd, j, [&](float x) { return v.term(x); }, [&](float x) { return v.term2(x); }, val, val2);
                if (j == 0 && (gd >> 27)) {
                    heap_val[buf][h] = val;
                    heap_val2[buf][h] = val2;
                }
            } else {
                const float val = np_node_value(tl, gd, j, [&](float x) { return v.term(x); });
                if (j == 0 && (gd >> 27)) heap_val[buf][h] = val;
            }
        }
        if constexpr (V::kCompact && !V::kSecondTree) {
            // survivors of the tile, straight from the staged values, compacted into the shared-memory stage s_out[buf]: a
            // warp takes the 512 elements [512 w, 512 (w + 1)), counts, and reserves its slots of the stage with one
            // shared-memory atomic (warp order inside the tile is arbitrary -- the consumer is a sort).
            const float *vals = tl;
            const int lane = lane_id(), wbase_e = warp_id() * 512;
            float4 xs[4];
            int cnt = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int e = wbase_e + r * 128 + lane * 4;
                xs[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e < sz) xs[r] = *reinterpret_cast<const float4 *>(&vals[e]);
                if (e + 3 >= sz) {  // tail of the (globally last) tile: mask what lies beyond it
                    if (e + 1 >= sz) xs[r].y = 0.f;
                    if (e + 2 >= sz) xs[r].z = 0.f;
                    if (e + 3 >= sz) xs[r].w = 0.f;
                    if (e >= sz) xs[r].x = 0.f;
                }
                cnt += (xs[r].x != 0.f) + (xs[r].y != 0.f) + (xs[r].z != 0.f) + (xs[r].w != 0.f);
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tt = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += tt;
            }
            const int wtot = __shfl_sync(0xffffffffu, incl, 31);
            unsigned int wb = 0;
            if (lane == 0 && wtot) wb = atomicAdd(&s_cnt[buf], (unsigned int)wtot);
            wb = __shfl_sync(0xffffffffu, wb, 0);
            float *dst = s_out[buf] + wb + (incl - cnt);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (xs[r].x != 0.f) *dst++ = xs[r].x;
                if (xs[r].y != 0.f) *dst++ = xs[r].y;
                if (xs[r].z != 0.f) *dst++ = xs[r].z;
                if (xs[r].w != 0.f) *dst++ = xs[r].w;
            }
        }
        __syncthreads();  // node values, survivor stage and dirty flag of this tile complete; the previous tile's staging buffer has no readers left
        if constexpr (V::kSecondTree) {
            // the other buffer's flag: warp 0 read it (fold of the previous tile) before arriving at the barrier above, and
            // the next tile's visit sets it again only after the barrier below
            if (threadIdx.x == 96) s_dirty[buf ^ 1] = 0;
        }
        if constexpr (V::kCompact) {
            // flush the previous tile's survivors: its stage, count and global base are complete since the barrier
            const int pb = buf ^ 1;
            const unsigned int c = s_fcnt[pb];
            if (c) {
                const unsigned long long b = s_base[pb];
                if (b + c <= v.capacity)
                    for (unsigned int i = threadIdx.x; i < c; i += NP_THREADS) v.out[b + i] = s_out[pb][i];
            }
        }
        if constexpr (V::kCompact || V::kSecondTree) __syncthreads();  // the stage / flag of the other buffer may be written again
        // ---- fold (warp 0): internal nodes take left + right, level by level, while the other warps visit the next tile
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const uint32_t internal = desc[t].internal;
#pragma unroll
            for (int lvl = 4; lvl >= 0; --lvl) {
                const int h = (1 << lvl) + lane;
                if (lane < (1 << lvl) && ((internal >> h) & 1u)) {
                    heap_val[buf][h] = fadd(heap_val[buf][2 * h], heap_val[buf][2 * h + 1]);
                    if constexpr (V::kSecondTree) heap_val2[buf][h] = fadd(heap_val2[buf][2 * h], heap_val2[buf][2 * h + 1]);
                }
                __syncwarp();
            }
            if (lane == 0) {
                partials[t] = heap_val[buf][1];
                if constexpr (V::kSecondTree) {
                    if (s_dirty[buf])  // holds undecided elements: re-reduced from the final tensor later
                        v.dirty_list[atomicAdd(v.dirty_count, 1u)] = t;
                    else
                        v.partials2[t] = heap_val2[buf][1];
                }
            }
        }
        if constexpr (V::kCompact) {
            // one global atomic per tile reserves the output slots; the copy stage -> global happens after the next
            // tile's barrier, so the atomic's latency is off the critical path
            if (threadIdx.x == 64) {
                const unsigned int c = s_cnt[buf];
                s_base[buf] = c ? atomicAdd(v.cursor, (unsigned long long)c) : 0ull;
                s_fcnt[buf] = c;  // read by the flush after the next tile's barrier
                s_cnt[buf] = 0;   // next written by the tile after next, which starts after the next tile's barrier
            }
        }
    }
    __syncthreads();
    if constexpr (V::kCompact) {  // the last tile of this CTA
        const int pb = buf ^ 1;
        const unsigned int c = s_fcnt[pb];
        const unsigned long long b = s_base[pb];
        if (b + c <= v.capacity)
            for (unsigned int i = threadIdx.x; i < c; i += NP_THREADS) v.out[b + i] = s_out[pb][i];
    }
    v.finish(aux);
}

// Fold the 2^depth partials pairwise (complete binary tree) and run a scalar epilogue.
enum { FIN_NONE = 0, FIN_MEAN = 1, FIN_VAR = 2, FIN_PRUNE1 = 3, FIN_PRUNE2 = 4 };

struct FinArgs {
    int mode;
    int64_t n;
    double q;
    int thr_mode;   // 0: float32 product/compare, 1: float64
    int std_smooth;
};

__device__ __forceinline__ float f32_ceil_of(double d) { return __double2float_ru(d); }

// First stage of the fold for big trees: every CTA folds 1024 consecutive partials -- a complete sub-tree of the
// (complete, binary) fold, so the additions are exactly the ones the one-CTA fold below would do.
__global__ void __launch_bounds__(512) np_fold1024_kernel(const float *__restrict__ partials, float *__restrict__ out) {
    __shared__ float a[1024], b[512];
    const float *src = partials + (size_t)blockIdx.x * 1024;
    a[threadIdx.x] = src[threadIdx.x];
    a[threadIdx.x + 512] = src[threadIdx.x + 512];
    __syncthreads();
    float *s = a, *d = b;
    for (int len = 1024; len > 1; len >>= 1) {
        const int half = len >> 1;
        if ((int)threadIdx.x < half) d[threadIdx.x] = fadd(s[2 * threadIdx.x], s[2 * threadIdx.x + 1]);
        __syncthreads();
        float *t = s;
        s = d;
        d = t;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

// scalar epilogue of a finished tree sum `total` (one thread)
__device__ void np_final_epilogue(float total, DevScalars *sc, const FinArgs &fa) {
    sc->tree_sum = total;
    double nd = (double)fa.n;
    if (fa.mode == FIN_MEAN || fa.mode == FIN_PRUNE1) {
        sc->mean = (float)((double)total / nd);  // np: true_divide(sum, intp count) in float64, cast back
    }
    if (fa.mode == FIN_PRUNE1) {
        // fp64 estimate of the std -> speculation band around the threshold.  All comparisons against a
        // double bound B are done in float32 as |x| < ceil32(B), which is equivalent for float32 |x|.
        double m = sc->sum_d / nd;
        double var = sc->sumsq_d / nd - m * m;
        if (var < 0) var = 0;
        double sd = sqrt(var);
        double thr = fa.thr_mode == 0 ? (double)((float)sd * (float)fa.q) : sd * fa.q;
        double lo = thr * (1.0 - 1.0 / 65536.0), hi = thr * (1.0 + 1.0 / 65536.0);
        if (!(thr > 0)) {  // zero, negative or NaN threshold: nothing can be pruned speculatively
            lo = 0.0;
            hi = 0.0;
        }
        sc->band_lo = (double)f32_ceil_of(lo);
        sc->band_hi = (double)f32_ceil_of(hi);
    }
    if (fa.mode == FIN_VAR || fa.mode == FIN_PRUNE2) {
        float var = (float)((double)total / nd);
        sc->var = var;
        sc->std_ = sqrtf(var);
    }
    if (fa.mode == FIN_PRUNE2) {
        double thr = fa.thr_mode == 0 ? (double)fmul(sc->std_, (float)fa.q) : (double)sc->std_ * fa.q;
        sc->thr = thr;
        float thr_f = f32_ceil_of(thr);  // |x| < thr  <=>  |x| < thr_f for float32 |x|
        float lo = (float)sc->band_lo, hi = (float)sc->band_hi;
        // the band [lo, hi) must contain the exact threshold, otherwise speculative decisions may be wrong
        bool ok = (thr_f >= lo && thr_f <= hi) || (thr != thr) || (lo == 0.f && hi == 0.f && !(thr > 0));
        sc->spec_failed = ok ? 0 : 1;
    }
}

// folds a[0, count) (count a power of two) pairwise in place (ping-pong with the `count` floats behind it); result in
// a[0] or a[count] -- the returned pointer.  All threads of one CTA.
__device__ float *np_fold_pow2(float *a, uint32_t count) {
    float *src = a, *dst = a + count;
    for (uint32_t len = count; len > 1; len >>= 1) {
        const uint32_t half = len >> 1;
        for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) dst[i] = fadd(src[2 * i], src[2 * i + 1]);
        __syncthreads();
        float *tmp = src;
        src = dst;
        dst = tmp;
    }
    return src;
}

__global__ void __launch_bounds__(1024) np_final_kernel(float *partials, uint32_t count, DevScalars *sc, FinArgs fa) {
    float *src = np_fold_pow2(partials, count);
    if (threadIdx.x != 0) return;
    np_final_epilogue(src[0], sc, fa);
}

// Decide the band elements with the exact threshold.  With `out` (fused k-means prologue, VisitApplyQuant) the decided
// values also join the prologue's statistics and the survivors are appended to the compacted array.
__global__ void prune_fixup_kernel(float *w, uint8_t *mask, const long long *side_idx, const float *side_val,
                                   unsigned long long cap, DevScalars *sc, float *out, unsigned long long *cursor,
                                   unsigned long long capacity) {
    unsigned long long cnt = sc->band_count;
    if (cnt > cap) cnt = cap;
    float thr_f = f32_ceil_of(sc->thr);
    unsigned long long pruned = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cnt;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        long long g = side_idx[i];
        float x = side_val[i];
        const bool p = fabsf(x) < thr_f;
        if (p) {
            w[g] = 0.f;
            mask[g] = 1;
            pruned++;
            x = 0.f;
        }
        if (out) {
            if (x == 0.f) x = 0.f;  // canonical zero
            atomicMin(&sc->min_ord, f2ord(x));
            atomicMax(&sc->max_ord, f2ord(x));
            const uint32_t ab = __float_as_uint(x) & 0x7fffffffu;
            if (ab != 0u) {
                atomicMax(&sc->amax_bits, ab);
                atomicMin(&sc->amin_nz_m1, ab - 1u);
                atomicAdd(&sc->n_nz, 1ull);
                const unsigned long long at = atomicAdd(cursor, 1ull);
                if (at < capacity) out[at] = x;
            }
        }
    }
    pruned = warp_sum_ull(pruned);
    if (lane_id() == 0 && pruned) atomicAdd(&sc->n_pruned, pruned);
}

// Fused k-means prologue (VisitApplyQuant): the scalars the apply pass did not track per element.  Every survivor has
// |x| >= band_lo and |x| <= the largest |x| of the original tensor (statistics pass), which is all the sort needs (any
// enclosing key range is valid); the survivor count is the compaction cursor.
__global__ void quant_scalars_kernel(DevScalars *sc, const unsigned long long *cursor) {
    sc->n_nz = *cursor;
    const float lo = (float)sc->band_lo;
    const uint32_t lo_m1 = lo > 0.f ? __float_as_uint(lo) - 1u : 0u;
    sc->amin_nz_m1 = min(sc->amin_nz_m1, lo_m1);
    sc->amax_bits = max(sc->amax_bits, sc->amax_all);
    if (sc->amax_all >= 0x7f800000u) sc->n_nonfinite = 1ull;
}

// Plain elementwise apply with a known threshold (std_smooth = False, and the safe fallback).
// mode 0: fresh apply over original data.  mode 1: resolve pass after a speculative pass -- elements that are
// already zero keep their mask byte, everything else is decided with the exact threshold.
__global__ void __launch_bounds__(256) prune_apply_kernel(float *w, uint8_t *mask, int64_t n, DevScalars *sc, int vec_ok,
                                                          int mode) {
    float thr_f = f32_ceil_of(sc->thr);
    unsigned long long pruned = 0;
    int64_t nvec = vec_ok ? (n >> 2) : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = *reinterpret_cast<const float4 *>(w + 4 * i);
        uint32_t old = mode ? *reinterpret_cast<const uint32_t *>(mask + 4 * i) : 0u;
        uint32_t m0 = fabsf(x.x) < thr_f, m1 = fabsf(x.y) < thr_f, m2 = fabsf(x.z) < thr_f, m3 = fabsf(x.w) < thr_f;
        uint32_t mm = m0 | (m1 << 8) | (m2 << 16) | (m3 << 24);
        if (mode) {
            uint32_t fresh = mm & ~old;  // newly pruned by this pass
            pruned += __popc(fresh);
            mm |= old;
        } else {
            pruned += m0 + m1 + m2 + m3;
        }
        if (m0 | m1 | m2 | m3) {
            x.x = m0 ? 0.f : x.x;
            x.y = m1 ? 0.f : x.y;
            x.z = m2 ? 0.f : x.z;
            x.w = m3 ? 0.f : x.w;
            *reinterpret_cast<float4 *>(w + 4 * i) = x;
        }
        *reinterpret_cast<uint32_t *>(mask + 4 * i) = mm;
    }
    for (int64_t i = (nvec << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float x = w[i];
        uint32_t old = mode ? mask[i] : 0u;
        uint32_t m = fabsf(x) < thr_f;
        if (m) w[i] = 0.f;
        if (mode) {
            pruned += (m & ~old) & 1u;
            m |= old;
        } else {
            pruned += m;
        }
        mask[i] = (uint8_t)m;
    }
    pruned = warp_sum_ull(pruned);
    if (lane_id() == 0 && pruned) atomicAdd(&sc->n_pruned, pruned);
}

__global__ void set_thr_kernel(DevScalars *sc, double thr) { sc->thr = thr; }

__global__ void __launch_bounds__(256) mask_apply_kernel(float *w, const uint8_t *mask, int64_t n, int vec_ok) {
    int64_t nvec = vec_ok ? (n >> 2) : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t mm = *reinterpret_cast<const uint32_t *>(mask + 4 * i);
        if (mm) {
            float4 x = *reinterpret_cast<const float4 *>(w + 4 * i);
            if (mm & 0xffu) x.x = 0.f;
            if (mm & 0xff00u) x.y = 0.f;
            if (mm & 0xff0000u) x.z = 0.f;
            if (mm & 0xff000000u) x.w = 0.f;
            *reinterpret_cast<float4 *>(w + 4 * i) = x;
        }
    }
    for (int64_t i = (nvec << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        if (mask[i]) w[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned4(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// ---- sharding: a rank owns a contiguous range of the reduction tree's tiles ----------------------------------
void np_shard_range(int64_t n, int rank, int world, int64_t *begin, int64_t *end, uint32_t *t0_out, uint32_t *t1_out) {
    const NpPlan p = np_plan(n);
    uint32_t t0 = 0, t1 = p.num_tiles;
    if (world > 1) {
        if (p.num_tiles >= (uint32_t)world) {
            t0 = (uint32_t)((uint64_t)p.num_tiles * rank / world);
            t1 = (uint32_t)((uint64_t)p.num_tiles * (rank + 1) / world);
        } else {  // too small to cut: rank 0 owns everything
            t0 = rank == 0 ? 0 : p.num_tiles;
            t1 = p.num_tiles;
        }
    }
    int64_t off0 = n, off1 = n;
    int sz;
    if (t0 < p.num_tiles) np_tile_root(n, p.depth, t0, off0, sz);
    if (t1 < p.num_tiles) np_tile_root(n, p.depth, t1, off1, sz);
    *begin = off0;
    *end = off1;
    if (t0_out) *t0_out = t0;
    if (t1_out) *t1_out = t1;
}

// ---- scalar exchange between ranks (all integers on the wire; doubles travel as per-rank slots) -------------------
enum { EX_NONE = 0, EX_STATS = 1, EX_PRUNE = 2, EX_QUANT = 3 };

__global__ void scal_pack_kernel(DevScalars *sc, long long *buf, int mode, int rank, int world) {
    if (mode == EX_STATS) {
        for (int r = 0; r < world; ++r) {
            buf[r] = r == rank ? __double_as_longlong(sc->sum_d) : 0;
            buf[world + r] = r == rank ? __double_as_longlong(sc->sumsq_d) : 0;
        }
    } else if (mode == EX_PRUNE) {
        buf[0] = (long long)sc->n_pruned;
        buf[1] = (long long)sc->band_dropped;
    } else if (mode == EX_QUANT) {
        sc->n_nz_local = sc->n_nz;
        buf[0] = (long long)sc->n_nz;
        buf[1] = (long long)sc->n_nonfinite;
        // second buffer (max): minima travel complemented
        buf[8] = (long long)sc->max_ord;
        buf[9] = (long long)(0xffffffffu - sc->min_ord);
        buf[10] = (long long)sc->amax_bits;
        buf[11] = (long long)(0xffffffffu - sc->amin_nz_m1);
    }
}
__global__ void scal_unpack_kernel(DevScalars *sc, const long long *buf, int mode, int world) {
    if (mode == EX_STATS) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < world; ++r) {  // rank order: every rank gets the same double
            a += __longlong_as_double(buf[r]);
            b += __longlong_as_double(buf[world + r]);
        }
        sc->sum_d = a;
        sc->sumsq_d = b;
    } else if (mode == EX_PRUNE) {
        sc->n_pruned = (unsigned long long)buf[0];
        sc->band_dropped = (unsigned long long)buf[1];
    } else if (mode == EX_QUANT) {
        sc->n_nz = (unsigned long long)buf[0];
        sc->n_nonfinite = (unsigned long long)buf[1];
        sc->max_ord = (uint32_t)buf[8];
        sc->min_ord = 0xffffffffu - (uint32_t)buf[9];
        sc->amax_bits = (uint32_t)buf[10];
        sc->amin_nz_m1 = 0xffffffffu - (uint32_t)buf[11];
    }
}

static void exchange_scalars(nnc_ctx *ctx, int mode) {
    if (ctx->world <= 1 || mode == EX_NONE) return;
    const int world = ctx->world;
    long long *buf = arena_alloc_t<long long>(ctx, std::max(16, 2 * world));
    NNC_LAUNCH(ctx, scal_pack_kernel, 1, 1, 0, ctx->d_scal, buf, mode, ctx->rank, world);
    if (mode == EX_STATS) {
        comm_allreduce(ctx, reinterpret_cast<int64_t *>(buf), 2 * world, 0);
    } else if (mode == EX_PRUNE) {
        comm_allreduce(ctx, reinterpret_cast<int64_t *>(buf), 2, 0);
    } else {
        comm_allreduce(ctx, reinterpret_cast<int64_t *>(buf), 2, 0);
        comm_allreduce(ctx, reinterpret_cast<int64_t *>(buf + 8), 4, 2);
    }
    NNC_LAUNCH(ctx, scal_unpack_kernel, 1, 1, 0, ctx->d_scal, buf, mode, world);
}

// fold of the tile partials (+ scalar epilogue): two stages for big trees
static void launch_final(nnc_ctx *ctx, float *partials, uint32_t count, const FinArgs &fa) {
    if (count >= 4096) {  // count is a power of two
        float *stage = partials + count;
        NNC_LAUNCH(ctx, np_fold1024_kernel, count / 1024, 512, 0, partials, stage);
        NNC_LAUNCH(ctx, np_final_kernel, 1, 1024, 0, stage, count / 1024, ctx->d_scal, fa);
    } else {
        NNC_LAUNCH(ctx, np_final_kernel, 1, 1024, 0, partials, count, ctx->d_scal, fa);
    }
}

// ---- the same over NVLink peer memory (one box, peer mailboxes connected, power-of-two rank count) -----------------------
// A rank's tiles are a complete sub-tree of the fold when the rank count is a power of two that divides the tile count:
// the rank folds ITS partials to one float, and ONE in-kernel exchange (peer.cuh: stores into every rank's mailbox, flags,
// local polling) carries that float together with the scalars the pass produced.  Every rank then folds the `world` floats
// up the remaining levels -- the same additions as the single-rank fold -- and combines the scalars in rank order.
// Replaces, per pass: a memset + an NCCL all-reduce of all tile partials (1 MB at 2^30 weights) + up to three small NCCL
// all-reduces with their pack / unpack kernels.
enum { EXW_PARTIAL = 0, EXW_SUM = 1, EXW_SUMSQ = 2, EXW_NPRUNED = 3, EXW_DROPPED = 4, EXW_NNZ = 5, EXW_NONFINITE = 6, EXW_MAXORD = 7,
       EXW_MINORD = 8, EXW_AMAX = 9, EXW_AMIN = 10, EXW_COUNT = 11 };

__global__ void __launch_bounds__(1024) np_final_peer_kernel(float *local, uint32_t count_local, DevScalars *sc, FinArgs fa, int modes,
                                                             PeerComm pc, unsigned long long *gathered, int *comm_error) {
    __shared__ unsigned long long words[EXW_COUNT];
    __shared__ float top[2 * PEER_MAX_WORLD];
    float *src = np_fold_pow2(local, count_local);
    if (threadIdx.x == 0) {
        words[EXW_PARTIAL] = (unsigned long long)__float_as_uint(src[0]);
        words[EXW_SUM] = (unsigned long long)__double_as_longlong(sc->sum_d);
        words[EXW_SUMSQ] = (unsigned long long)__double_as_longlong(sc->sumsq_d);
        words[EXW_NPRUNED] = sc->n_pruned;
        words[EXW_DROPPED] = sc->band_dropped;
        words[EXW_NNZ] = sc->n_nz;
        words[EXW_NONFINITE] = sc->n_nonfinite;
        words[EXW_MAXORD] = sc->max_ord;
        words[EXW_MINORD] = sc->min_ord;
        words[EXW_AMAX] = sc->amax_bits;
        words[EXW_AMIN] = sc->amin_nz_m1;
    }
    __syncthreads();
    unsigned long long xseq = *peer_counter(pc);
    const bool ok = peer_allgather_tagged(pc, words, EXW_COUNT, gathered, EXW_COUNT, ++xseq);
    if (threadIdx.x == 0) {
        *peer_counter(pc) = xseq;
        if (!ok) *comm_error = 1;
    }
    if (threadIdx.x < (unsigned)pc.world) top[threadIdx.x] = __uint_as_float((uint32_t)gathered[(size_t)threadIdx.x * EXW_COUNT + EXW_PARTIAL]);
    __syncthreads();
    float *t = np_fold_pow2(top, (uint32_t)pc.world);
    if (threadIdx.x != 0) return;
    const int world = pc.world;
    if (modes & (1 << EX_STATS)) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < world; ++r) {  // rank order: every rank gets the same double
            a += __longlong_as_double((long long)gathered[(size_t)r * EXW_COUNT + EXW_SUM]);
            b += __longlong_as_double((long long)gathered[(size_t)r * EXW_COUNT + EXW_SUMSQ]);
        }
        sc->sum_d = a;
        sc->sumsq_d = b;
    }
    if (modes & (1 << EX_PRUNE)) {
        unsigned long long a = 0, b = 0;
        for (int r = 0; r < world; ++r) {
            a += gathered[(size_t)r * EXW_COUNT + EXW_NPRUNED];
            b += gathered[(size_t)r * EXW_COUNT + EXW_DROPPED];
        }
        sc->n_pruned = a;
        sc->band_dropped = b;
    }
    if (modes & (1 << EX_QUANT)) {
        sc->n_nz_local = sc->n_nz;
        unsigned long long a = 0, b = 0;
        uint32_t mx = 0, mn = 0xffffffffu, am = 0, an = 0xffffffffu;
        for (int r = 0; r < world; ++r) {
            const unsigned long long *g = gathered + (size_t)r * EXW_COUNT;
            a += g[EXW_NNZ];
            b += g[EXW_NONFINITE];
            mx = max(mx, (uint32_t)g[EXW_MAXORD]);
            mn = min(mn, (uint32_t)g[EXW_MINORD]);
            am = max(am, (uint32_t)g[EXW_AMAX]);
            an = min(an, (uint32_t)g[EXW_AMIN]);
        }
        sc->n_nz = a;
        sc->n_nonfinite = b;
        sc->max_ord = mx;
        sc->min_ord = mn;
        sc->amax_bits = am;
        sc->amin_nz_m1 = an;
    }
    np_final_epilogue(t[0], sc, fa);
}

static bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }

// true when the fold + exchange of a pass can run over the peer mailboxes (see np_final_peer_kernel)
static bool peer_final_ok(nnc_ctx *ctx, uint32_t num_tiles) {
    return ctx->world > 1 && ctx->peer_enabled && ctx->world <= PEER_MAX_WORLD && is_pow2((uint32_t)ctx->world) && num_tiles >= (uint32_t)ctx->world &&
           num_tiles % (uint32_t)ctx->world == 0 && !getenv("NNC_NO_PEER_FINAL");
}

// partials: the 2 * num_tiles + 2 floats of a tree (this rank's tiles filled); modes: bit mask of (1 << EX_*)
static void launch_final_peer(nnc_ctx *ctx, float *partials, uint32_t num_tiles, const FinArgs &fa, int modes) {
    const uint32_t L = num_tiles / (uint32_t)ctx->world;
    float *mine = partials + ctx->sh.t0;  // [t0, t0 + L): a complete sub-tree
    float *scratch = arena_alloc_t<float>(ctx, 2 * (size_t)std::max<uint32_t>(L, 2048) + 16);
    float *local = scratch;
    uint32_t count = L;
    if (L >= 4096) {  // first stage of the local fold: 1024 partials per CTA (complete sub-trees again)
        NNC_LAUNCH(ctx, np_fold1024_kernel, L / 1024, 512, 0, mine, scratch);
        count = L / 1024;
    } else {
        NNC_CUDA(cudaMemcpyAsync(scratch, mine, sizeof(float) * L, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    PeerComm pc;
    memset(&pc, 0, sizeof(pc));
    pc.enabled = 1;
    pc.rank = ctx->rank;
    pc.world = ctx->world;
    for (int r = 0; r < ctx->world; ++r) pc.mail[r] = static_cast<unsigned long long *>(ctx->peer_mail[r]);
    unsigned long long *gathered = arena_alloc_t<unsigned long long>(ctx, (size_t)ctx->world * EXW_COUNT);
    if (!ctx->d_comm_error) {
        NNC_CUDA(cudaMalloc(&ctx->d_comm_error, sizeof(int)));
        NNC_CUDA(cudaMemsetAsync(ctx->d_comm_error, 0, sizeof(int), ctx->stream));
    }
    NNC_LAUNCH(ctx, np_final_peer_kernel, 1, 1024, 0, local, count, ctx->d_scal, fa, modes, pc, gathered, ctx->d_comm_error);
}

static int tree_grid(nnc_ctx *ctx, uint32_t tiles) {
    int64_t g = (int64_t)ctx->sm_count * 8;
    if ((int64_t)tiles < g) g = tiles;
    return (int)g;
}

// One reduction over the whole (possibly sharded) tensor: tile descriptors for the GLOBAL n, this rank's tiles
// reduced here, the float32 tile partials all-gathered bit for bit (sum of int64 words, the other ranks' slots are
// zero), and the fold done redundantly -- and therefore identically -- on every rank.
template <class V>
static NpTileDesc *run_tree(nnc_ctx *ctx, const float *d_w, V v, const FinArgs &fa, int exchange_mode) {
    const int64_t n = ctx->sh.n_global;
    NpPlan p = np_plan(n);
    float *partials = arena_alloc_t<float>(ctx, 2 * (size_t)p.num_tiles + 2);
    NpTileDesc *desc = static_cast<NpTileDesc *>(ctx->desc_ptr);
    if (ctx->desc_n != n || !desc) {  // the descriptors only depend on n: built once, kept for the following calls
        const size_t need = sizeof(NpTileDesc) * (size_t)p.num_tiles;
        if (need > ctx->desc_bytes) {
            if (ctx->desc_ptr) {
                NNC_CUDA(cudaStreamSynchronize(ctx->stream));
                cudaFree(ctx->desc_ptr);
                ctx->desc_ptr = nullptr;
                ctx->desc_bytes = 0;
            }
            NNC_CUDA(cudaMalloc(&ctx->desc_ptr, need));
            ctx->desc_bytes = need;
        }
        desc = static_cast<NpTileDesc *>(ctx->desc_ptr);
        NNC_LAUNCH(ctx, np_tiles_kernel, (p.num_tiles + 127) / 128, 128, 0, n, p.depth, desc);
        ctx->desc_n = n;
    }
    const bool peer_final = peer_final_ok(ctx, p.num_tiles);
    if (ctx->world > 1 && !peer_final) NNC_CUDA(cudaMemsetAsync(partials, 0, sizeof(float) * (2 * (size_t)p.num_tiles + 2), ctx->stream));
    const uint32_t t0 = ctx->sh.t0, t1 = ctx->sh.t1;
    const size_t dyn = np_tree_smem<V>();
    func_dyn_smem(ctx, (const void *)np_tree_kernel<V>, dyn);
    if (t1 > t0)
        NNC_LAUNCH_AS(ctx, V::kName, np_tree_kernel<V>, tree_grid(ctx, t1 - t0), NP_THREADS, dyn, d_w, t0, t1, ctx->sh.begin,
                   aligned16(d_w) ? 1 : 0, desc, partials, (const uint32_t *)nullptr, (const unsigned int *)nullptr, v);
    if (peer_final) {  // local fold + one in-kernel exchange (partial and scalars together) + the top of the fold
        launch_final_peer(ctx, partials, p.num_tiles, fa, exchange_mode == EX_NONE ? 0 : (1 << exchange_mode));
        return desc;
    }
    if (ctx->world > 1) comm_allreduce(ctx, reinterpret_cast<int64_t *>(partials), (int)((p.num_tiles + 1) / 2), 0);
    exchange_scalars(ctx, exchange_mode);
    launch_final(ctx, partials, p.num_tiles, fa);
    return desc;
}

static void clear_scalars(nnc_ctx *ctx) {
    DevScalars z;
    memset(&z, 0, sizeof(z));
    z.min_ord = 0xffffffffu;
    z.min_nz_ord = 0xffffffffu;
    z.amin_nz_m1 = 0xffffffffu;
    *ctx->h_scal = z;
    NNC_CUDA(cudaMemcpyAsync(ctx->d_scal, ctx->h_scal, sizeof(DevScalars), cudaMemcpyHostToDevice, ctx->stream));
}

void np_stats(nnc_ctx *ctx, const float *d_w, int64_t n) {
    clear_scalars(ctx);
    VisitStats v1;
    v1.sc = ctx->d_scal;
    v1.probe = d_w;
    v1.probe_n = n;
    const int64_t ng = ctx->sh.n_global;
    run_tree(ctx, d_w, v1, FinArgs{FIN_MEAN, ng, 0.0, 0, 1}, EX_NONE);
    VisitCenSq v2;
    v2.sc = ctx->d_scal;
    v2.mean = 0.f;
    run_tree(ctx, d_w, v2, FinArgs{FIN_VAR, ng, 0.0, 0, 1}, EX_NONE);
}

// mean (NumPy tree), min / max, key range, non-zero count -- and the survivors themselves, compacted into d_out
// (capacity elements) in the same read of the tensor.
void quant_prologue(nnc_ctx *ctx, const float *d_w, int64_t n, float *d_out, int64_t capacity) {
    clear_scalars(ctx);
    (void)n;
    const int64_t ng = ctx->sh.n_global;
    unsigned long long *cursor = arena_alloc_t<unsigned long long>(ctx, 1);
    NNC_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), ctx->stream));
    VisitQuant v;
    v.sc = ctx->d_scal;
    v.out = d_out;
    v.cursor = cursor;
    v.capacity = (unsigned long long)capacity;
    run_tree(ctx, d_w, v, FinArgs{FIN_MEAN, ng, 0.0, 0, 1}, EX_QUANT);
}

void prune_device(nnc_ctx *ctx, float *d_w, int64_t n, double q, int std_smooth, int thr_mode, uint8_t *d_mask, QuantFuse *fuse) {
    clear_scalars(ctx);
    if (fuse) fuse->done = false;
    const int vec_ok = aligned16(d_w) && aligned4(d_mask);
    const int ew_grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 16, (n / 4 + 255) / 256 + 1);
    if (!std_smooth) {
        double thr = thr_mode == 0 ? (double)(float)q : q;
        NNC_LAUNCH(ctx, set_thr_kernel, 1, 1, 0, ctx->d_scal, thr);
        NNC_LAUNCH(ctx, prune_apply_kernel, ew_grid, 256, 0, d_w, d_mask, n, ctx->d_scal, vec_ok, 0);
        exchange_scalars(ctx, EX_PRUNE);
        prof_mark(ctx, "apply");
        return;
    }
    // pass 1: NumPy mean + fp64 estimate of the std -> speculation band
    VisitStats v1;
    v1.sc = ctx->d_scal;
    v1.probe = d_w;
    v1.probe_n = n;
    const int64_t ng = ctx->sh.n_global;
    run_tree(ctx, d_w, v1, FinArgs{FIN_PRUNE1, ng, q, thr_mode, 1}, EX_STATS);
    prof_mark(ctx, "mean");
    // pass 2: NumPy var/std/threshold + speculative apply
    unsigned long long cap = (unsigned long long)std::max<int64_t>(65536, n / 256);
    long long *side_idx = arena_alloc_t<long long>(ctx, cap);
    float *side_val = arena_alloc_t<float>(ctx, cap);
    unsigned long long *cursor = nullptr;
    if (fuse && fuse->out) {
        // pass 2 fused with the k-means prologue of the pruned tensor (VisitApplyQuant)
        const NpPlan p = np_plan(ng);
        float *partials2 = arena_alloc_t<float>(ctx, 2 * (size_t)p.num_tiles + 2);
        uint32_t *dirty_list = arena_alloc_t<uint32_t>(ctx, (size_t)p.num_tiles);
        unsigned long long *ctr = arena_alloc_t<unsigned long long>(ctx, 2);  // survivor cursor, dirty-tile count
        NNC_CUDA(cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), ctx->stream));
        const bool peer_final2 = peer_final_ok(ctx, p.num_tiles);
        if (ctx->world > 1 && !peer_final2) NNC_CUDA(cudaMemsetAsync(partials2, 0, sizeof(float) * (2 * (size_t)p.num_tiles + 2), ctx->stream));
        cursor = ctr;
        unsigned int *dirty_count = reinterpret_cast<unsigned int *>(ctr + 1);
        VisitApplyQuant v2;
        v2.sc = ctx->d_scal;
        v2.w = d_w;
        v2.mask = d_mask;
        v2.side_idx = side_idx;
        v2.side_val = side_val;
        v2.side_cap = cap;
        v2.mask_vec_ok = aligned4(d_mask) ? 1 : 0;
        v2.out = fuse->out;
        v2.cursor = cursor;
        v2.capacity = (unsigned long long)fuse->capacity;
        v2.partials2 = partials2;
        v2.dirty_list = dirty_list;
        v2.dirty_count = dirty_count;
        v2.mean = v2.lo = v2.hi = 0.f;
        NpTileDesc *desc = run_tree(ctx, d_w, v2, FinArgs{FIN_PRUNE2, ng, q, thr_mode, 1}, EX_NONE);
        prof_mark(ctx, "var+apply");
        NNC_LAUNCH(ctx, prune_fixup_kernel, 64, 256, 0, d_w, d_mask, side_idx, side_val, cap, ctx->d_scal, fuse->out, cursor,
                   (unsigned long long)fuse->capacity);
        NNC_LAUNCH(ctx, quant_scalars_kernel, 1, 1, 0, ctx->d_scal, cursor);
        // tiles that held undecided elements: their partial of the second tree from the final tensor
        func_dyn_smem(ctx, (const void *)np_tree_kernel<VisitPlain>, np_tree_smem<VisitPlain>());
        NNC_LAUNCH_AS(ctx, VisitPlain::kName, np_tree_kernel<VisitPlain>, std::max(1, std::min<int>(ctx->sm_count * 4, (int)std::min<uint32_t>(p.num_tiles, 1u << 20))),
                   NP_THREADS, np_tree_smem<VisitPlain>(), d_w, 0u, 0u, ctx->sh.begin, aligned16(d_w) ? 1 : 0, desc, partials2, (const uint32_t *)dirty_list,
                   (const unsigned int *)dirty_count, VisitPlain{});
        if (peer_final2) {
            launch_final_peer(ctx, partials2, p.num_tiles, FinArgs{FIN_MEAN, ng, 0.0, 0, 1}, (1 << EX_PRUNE) | (1 << EX_QUANT));
        } else {
            exchange_scalars(ctx, EX_PRUNE);
            if (ctx->world > 1) comm_allreduce(ctx, reinterpret_cast<int64_t *>(partials2), (int)((p.num_tiles + 1) / 2), 0);
            exchange_scalars(ctx, EX_QUANT);
            launch_final(ctx, partials2, p.num_tiles, FinArgs{FIN_MEAN, ng, 0.0, 0, 1});
        }
        prof_mark(ctx, "fixup");
    } else {
    VisitCenSqApply v2;
    v2.sc = ctx->d_scal;
    v2.w = d_w;
    v2.mask = d_mask;
    v2.side_idx = side_idx;
    v2.side_val = side_val;
    v2.side_cap = cap;
    v2.mask_vec_ok = aligned4(d_mask) ? 1 : 0;
    v2.mean = v2.lo = v2.hi = 0.f;
    run_tree(ctx, d_w, v2, FinArgs{FIN_PRUNE2, ng, q, thr_mode, 1}, EX_NONE);
    prof_mark(ctx, "var+apply");
    NNC_LAUNCH(ctx, prune_fixup_kernel, 64, 256, 0, d_w, d_mask, side_idx, side_val, cap, ctx->d_scal, (float *)nullptr,
               (unsigned long long *)nullptr, 0ull);
    exchange_scalars(ctx, EX_PRUNE);
    prof_mark(ctx, "fixup");
    }
    read_scalars(ctx);
    const DevScalars &s = *ctx->h_scal;
    if (fuse && fuse->out) fuse->done = !(s.spec_failed || s.band_dropped);
    if (s.spec_failed || s.band_dropped) {
        // Either the exact threshold left the speculation band (non-finite or badly scaled data) or the side
        // list overflowed.  Everything that is still non-zero is re-decided with the exact threshold.  This is
        // only wrong if the band was ABOVE the exact threshold (elements were zeroed that should have stayed),
        // which cannot be undone in place: report it.
        if (s.spec_failed && s.thr < s.band_lo && s.n_pruned > 0)
            NNC_FAIL(NNC_ERR_INTERNAL, "prune: exact threshold %.9g below speculation band [%.9g, %.9g]", s.thr, s.band_lo,
                     s.band_hi);
        if (ctx->world > 1)  // n_pruned is already a global count: the resolve pass would add local counts to it
            NNC_FAIL(NNC_ERR_INTERNAL, "prune: speculation band missed on a sharded tensor (thr %.9g, band [%.9g, %.9g])", s.thr,
                     s.band_lo, s.band_hi);
        NNC_LAUNCH(ctx, prune_apply_kernel, ew_grid, 256, 0, d_w, d_mask, n, ctx->d_scal, vec_ok, 1);
        prof_mark(ctx, "resolve");
        read_scalars(ctx);
    }
}

void mask_apply_device(nnc_ctx *ctx, float *d_w, const uint8_t *d_mask, int64_t n) {
    const int vec_ok = aligned16(d_w) && aligned4(d_mask);
    const int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 16, (n / 4 + 255) / 256 + 1);
    NNC_LAUNCH(ctx, mask_apply_kernel, grid, 256, 0, d_w, d_mask, n, vec_ok);
}

}  // namespace nnc
