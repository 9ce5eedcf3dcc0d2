// table.cuh -- the "region table": an exact piecewise description of scikit-learn's float32 label rule
// on the real line, built from the current centroids.
//
// sklearn labels a centred sample x' with the FIRST j minimising d_j = fl(fl(c_j c_j) + fl(fl(-2 x') c_j))
// (sklearn/cluster/_k_means_lloyd.pyx:196-213).  In exact arithmetic that is the nearest centroid; in
// float32 the comparison of two neighbouring centroids a < b can go either way only in a narrow window
// around their midpoint.  With u = 2^-24 and M >= max(|x'|, |c|) every computed d_j is within
// 2u(1+4u)(c_j^2 + 2|x' c_j|) + eta <= E1 := 6u(1+4u)M^2 + eta of its exact value, and the exact margin
// between neighbours is 2(b-a)|mid - x'|.  So outside the "zone" |x' - mid| <= E2/(2(b-a)), E2 = 2 E1, the
// float32 comparison agrees with exact arithmetic, and (margins to farther centroids being larger) the label is
// the exact nearest distinct centroid, lowest id among exact duplicates.  Overlapping zones are merged into
// groups.  The line is thereby cut into alternating regions
//     SAFE_0 | ZONE_0 | SAFE_1 | ZONE_1 | ... | ZONE_{G-1} | SAFE_G
// SAFE_s has one label; inside ZONE_g the label is found by evaluating the float32 rule over the group's
// candidates (distinct centroids gp_lo[g] .. gp_hi[g]+1).  Both are bit-exact restatements of the rule.
#pragma once
#include <float.h>

#include "common.cuh"

namespace nnc {

constexpr int TB_KMAX = 1024;
constexpr int TB_THREADS = 1024;

struct RegionTable {
    int k;   // clusters
    int m;   // distinct centroid values
    int G;   // zone groups; regions = 2G + 1
    int pad_;
    float dv[TB_KMAX];      // distinct centroid values, ascending
    float dcn[TB_KMAX];     // fl(dv * dv)
    int down[TB_KMAX];      // owner id = lowest cluster id with that value
    int gp_lo[TB_KMAX];     // group g spans adjacent pairs gp_lo..gp_hi, i.e. distinct indices gp_lo..gp_hi+1
    int gp_hi[TB_KMAX];
    float rstart[2 * TB_KMAX + 2];  // region r covers x' in [rstart[r], rstart[r+1]);  r even: SAFE, odd: ZONE
};

__device__ __forceinline__ int safe_distinct_index(const int *gp_hi, int s) { return s == 0 ? 0 : gp_hi[s - 1] + 1; }

__device__ __forceinline__ float f32_nextup(float f) {
    if (f != f || f == INFINITY) return f;
    if (f == 0.f) return __uint_as_float(1u);
    uint32_t u = __float_as_uint(f);
    return __uint_as_float(f > 0.f ? u + 1u : u - 1u);
}

// The float32 label rule over the distinct candidates [lo, hi] (inclusive); ties -> lowest owner id.
__device__ __forceinline__ int zone_argmin(float xc, const float *dv, const float *dcn, const int *down, int lo, int hi) {
    const float m2x = fmul(-2.0f, xc);
    float best = skl_dist(m2x, dv[lo], dcn[lo]);
    int bi = lo, bid = down[lo];
    for (int j = lo + 1; j <= hi; ++j) {
        float d = skl_dist(m2x, dv[j], dcn[j]);
        int id = down[j];
        if (d < best || (d == best && id < bid)) {
            best = d;
            bi = j;
            bid = id;
        }
    }
    return bi;
}

template <class T, class Op>
__device__ __forceinline__ T block_scan_incl(T v, Op op, T *s_warp /*[32]*/) {
    const int lane = lane_id(), w = warp_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = op(t, v);
    }
    if (lane == 31) s_warp[w] = v;
    __syncthreads();
    if (w == 0) {
        T x = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x = op(t, x);
        }
        s_warp[lane] = x;
    }
    __syncthreads();
    if (w > 0) v = op(s_warp[w - 1], v);
    __syncthreads();
    return v;
}

struct TableScratch {
    unsigned long long keys[TB_KMAX];
    double a[TB_KMAX];
    double b[TB_KMAX];
    double warp_d[32];
    int warp_i[32];
    float warp_f[32];
    int flag[TB_KMAX];
};

// Build the table for centroids c[0..k) (centred space).  Must be called by all TB_THREADS threads of a CTA.
// xabs_max: max |x'| over the data.
static __device__ void build_region_table(const float *c, int k, float xabs_max, RegionTable *T, TableScratch &S) {
    const int tid = threadIdx.x;
    // ---- 1. sort (value, id)
    int P = 32;
    while (P < k) P <<= 1;
    float cv = 0.f;
    if (tid < k) {
        cv = c[tid];
        if (cv == 0.f) cv = 0.f;  // -0.0 -> +0.0 (identical distances)
    }
    if (tid < P) S.keys[tid] = tid < k ? (((unsigned long long)f2ord(cv) << 32) | (unsigned)tid) : ~0ull;
    // max |c|
    float cm = tid < k ? fabsf(cv) : 0.f;
    cm = warp_max_f(cm);
    if (lane_id() == 0) S.warp_f[warp_id()] = cm;
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (tid < P) {
                int partner = tid ^ stride;
                if (partner > tid) {
                    bool asc = (tid & size) == 0;
                    unsigned long long x = S.keys[tid], y = S.keys[partner];
                    if ((x > y) == asc) {
                        S.keys[tid] = y;
                        S.keys[partner] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
    float M = xabs_max;
    for (int i = 0; i < TB_THREADS / 32; ++i) M = fmaxf(M, S.warp_f[i]);
    // ---- 2. distinct values
    uint32_t myord = 0, prevord = 0;
    int myid = 0;
    if (tid < k) {
        myord = (uint32_t)(S.keys[tid] >> 32);
        myid = (int)(S.keys[tid] & 0xffffffffu);
        prevord = tid > 0 ? (uint32_t)(S.keys[tid - 1] >> 32) : 0;
    }
    int first = (tid < k) && (tid == 0 || myord != prevord);
    int incl = block_scan_incl<int>(first, [](int x, int y) { return x + y; }, S.warp_i);
    if (first) {
        int di = incl - 1;
        float v = ord2f(myord);
        T->dv[di] = v;
        T->dcn[di] = fmul(v, v);
        T->down[di] = myid;
        S.a[di] = (double)v;  // stage values for the pair computation
    }
    if (tid == k - 1) {
        T->m = incl;
        T->k = k;
        S.flag[0] = incl;
    }
    __syncthreads();
    const int m = S.flag[0];
    __syncthreads();
    // ---- 3. zones of adjacent pairs
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double Md = (double)M;
    const double E2 = 12.0 * u * (1.0 + 4.0 * u) * Md * Md + 1e-42;
    double L = DBL_MAX, R = -DBL_MAX;
    const int npairs = m - 1;
    if (tid < npairs) {
        double a = S.a[tid], b = S.a[tid + 1];
        double mid = 0.5 * (a + b);
        double delta = E2 / (2.0 * (b - a));
        L = mid - delta;
        R = mid + delta;
    }
    __syncthreads();
    // prefix max of R, suffix min of L
    double pmaxR = block_scan_incl<double>(R, [](double x, double y) { return x > y ? x : y; }, S.warp_d);
    if (tid < npairs) S.b[tid] = L;
    __syncthreads();
    double Lrev = tid < npairs ? S.b[npairs - 1 - tid] : DBL_MAX;
    double sminL_rev = block_scan_incl<double>(Lrev, [](double x, double y) { return x < y ? x : y; }, S.warp_d);
    if (tid < npairs) S.a[npairs - 1 - tid] = sminL_rev;  // S.a[i] = min_{j>=i} L_j
    __syncthreads();
    // ---- 4. groups
    int brk = 0;  // a new group starts at pair tid+1
    if (tid + 1 < npairs) brk = pmaxR < S.a[tid + 1];
    S.flag[tid] = brk;
    int brk_incl = block_scan_incl<int>(brk, [](int x, int y) { return x + y; }, S.warp_i);
    if (tid < npairs) {
        int gid = brk_incl - brk;  // breaks strictly before this pair
        bool is_start = tid == 0 || S.flag[tid - 1];
        bool is_end = tid == npairs - 1 || brk;
        if (is_start) {
            T->gp_lo[gid] = tid;
            T->rstart[2 * gid + 1] = __double2float_ru(S.a[tid]);  // smallest float >= group's min L
        }
        if (is_end) {
            T->gp_hi[gid] = tid;
            T->rstart[2 * gid + 2] = f32_nextup(__double2float_rd(pmaxR));  // smallest float > group's max R
        }
        if (tid == npairs - 1) {
            T->G = gid + 1;
            T->rstart[2 * (gid + 1) + 1] = INFINITY;
        }
    }
    if (tid == 0) {
        T->rstart[0] = -INFINITY;
        if (npairs <= 0) {
            T->G = 0;
            T->rstart[1] = INFINITY;
        }
    }
    __syncthreads();
}

}  // namespace nnc
