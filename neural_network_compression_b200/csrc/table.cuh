// table.cuh -- the "region table": an exact piecewise description of scikit-learn's float32 label rule
// on the real line, built from the current centroids.
//
// sklearn labels a centred sample x' with the FIRST j minimising d_j = fl(fl(c_j c_j) + fl(fl(-2 x') c_j))
// (sklearn/cluster/_k_means_lloyd.pyx:196-213).  In exact arithmetic that is the nearest centroid; in float32
// the comparison of two centroids can go either way when x' is close to their midpoint.  With u = 2^-24,
// u' = u(1+4u), every computed d_j is within err_j(x') = 2u'(c_j^2 + 2|x' c_j|) + eta of its exact value.
//
// For every pair of ADJACENT distinct centroids a < b (mid = (a+b)/2) a half-width delta is computed such that
// 2(b-a)|x' - mid| > err_a(x') + err_b(x') for all data x' with |x' - mid| > delta, i.e. the float32 comparison
// of a and b is decided exactly as in real arithmetic outside [A, B] = [mid - delta, mid + delta].
// Lemma: for x' < A the left centroid a beats EVERY centroid right of it in float32, and
// for x' > B the right centroid b beats every centroid left of it.  Hence centroid j can only be the float32
// argmin for  lo_j <= x' <= hi_j,  lo_j = max_{pairs left of j} A,  hi_j = min_{pairs right of j} B.  Both
// sequences are non-decreasing in j, so the candidates at x' are the contiguous range [J2(x'), J1(x')],
// J1 = #{j >= 1 : lo_j <= x'},  J2 = #{j <= m-2 : hi_j < x'}.  Sorting the 2(m-1) breakpoints cuts the line into
// 2m-1 regions with constant (J2, J1): J1 == J2 is a SAFE region (one label, no arithmetic); otherwise the
// float32 rule is evaluated over the candidates J2..J1 (a ZONE).  Near-duplicate centroids (delta huge) simply
// never eliminate each other and stay joint candidates inside their common cell; they do not widen anyone
// else's zone.  Both cases are bit-exact restatements of the rule.
#pragma once
#include <float.h>

#include "common.cuh"

namespace nnc {

constexpr int TB_KMAX = 1024;
constexpr int TB_THREADS = 1024;

template <int KM>
struct RegionTableT {
    int k;   // clusters
    int m;   // distinct centroid values
    int R;   // regions = 2m - 1 (region r covers x' in [rstart[r], rstart[r+1]))
    int pad_;
    float dv[KM];      // distinct centroid values, ascending
    float dcn[KM];     // fl(dv * dv)
    int down[KM];      // owner id = lowest cluster id with that value
    float rstart[2 * KM + 2];  // rstart[0] = -inf, rstart[R] = +inf
    short rJ2[2 * KM + 2];     // candidates of region r: distinct indices rJ2[r] .. rJ1[r]
    short rJ1[2 * KM + 2];
};
using RegionTable = RegionTableT<TB_KMAX>;

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

// Exclusive scan of in[0, n) into out[0, n] (out[n] = total) by ONE CTA of 1024 threads: rounds of 8192 consecutive
// elements (8 per thread, coalesced), a block scan per round, the carry kept in a register.
template <class TIn, class TOut>
__device__ __forceinline__ TOut cta_exclusive_scan(const TIn *__restrict__ in, long long n, TOut *__restrict__ out, TOut *s_warp /*[32]*/) {
    constexpr int PER = 8;
    TOut carry = 0;
    for (long long base = 0; base < n; base += (long long)blockDim.x * PER) {
        const long long i0 = base + (long long)threadIdx.x * PER;
        TOut v[PER], sum = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            v[j] = i0 + j < n ? (TOut)in[i0 + j] : (TOut)0;
            sum += v[j];
        }
        TOut incl = block_scan_incl<TOut>(sum, [](TOut a, TOut b) { return a + b; }, s_warp);
        TOut run = carry + incl - sum;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (i0 + j < n) out[i0 + j] = run;
            run += v[j];
        }
        // total of the round = inclusive value of the last thread
        __shared__ TOut s_total;
        if (threadIdx.x == blockDim.x - 1) s_total = incl;
        __syncthreads();
        carry += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
    return carry;
}

// ---- multi-CTA exclusive scan: chunk sums -> per-chunk rescan; every CTA of the second kernel adds up the sums of the
// chunks before it itself (at most 1024 of them: one per thread), which saves a launch of a one-CTA kernel in between.
// out[0, n) = exclusive prefix, out[n] = total.  chunk_sum: gridDim.x entries of scratch.
template <class TIn, class TOut>
__global__ void __launch_bounds__(1024) scan_chunk_sum_kernel(const TIn *__restrict__ in, long long n, long long chunk, TOut *chunk_sum) {
    __shared__ TOut s_warp[32];
    const long long lo = (long long)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    TOut sum = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) sum += (TOut)in[i];
    const TOut incl = block_scan_incl<TOut>(sum, [](TOut a, TOut b) { return a + b; }, s_warp);
    if (threadIdx.x == blockDim.x - 1) chunk_sum[blockIdx.x] = incl;
}
template <class TIn, class TOut>
__global__ void __launch_bounds__(1024) scan_chunk_apply_kernel(const TIn *__restrict__ in, long long n, long long chunk,
                                                                const TOut *__restrict__ chunk_sum, TOut *__restrict__ out) {
    __shared__ TOut s_warp[32];
    __shared__ TOut s_total;
    constexpr int PER = 4;
    const long long lo = (long long)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    TOut carry;
    {  // offset of this chunk (gridDim.x <= blockDim.x)
        const TOut mine = threadIdx.x < blockIdx.x ? chunk_sum[threadIdx.x] : (TOut)0;
        const TOut incl = block_scan_incl<TOut>(mine, [](TOut a, TOut b) { return a + b; }, s_warp);
        if (threadIdx.x == blockDim.x - 1) s_total = incl;
        __syncthreads();
        carry = s_total;
        __syncthreads();
    }
    for (long long base = lo; base < hi; base += (long long)blockDim.x * PER) {
        const long long i0 = base + (long long)threadIdx.x * PER;
        TOut v[PER], sum = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            v[j] = i0 + j < hi ? (TOut)in[i0 + j] : (TOut)0;
            sum += v[j];
        }
        const TOut incl = block_scan_incl<TOut>(sum, [](TOut a, TOut b) { return a + b; }, s_warp);
        TOut run = carry + incl - sum;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (i0 + j < hi) out[i0 + j] = run;
            run += v[j];
        }
        if (threadIdx.x == blockDim.x - 1) s_total = incl;
        __syncthreads();
        carry += s_total;
        __syncthreads();
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = carry;  // the total
}

template <int KM>
struct TableScratchT {
    unsigned long long keys[KM];
    double a[KM];
    double b[KM];
    double warp_d[32];
    int warp_i[32];
    float warp_f[32];
    int flag[KM];
    float tlo[KM];
    float thi[KM];
};
using TableScratch = TableScratchT<TB_KMAX>;

// Build the table for centroids c[0..k) (centred space).  Must be called by all TB_THREADS threads of a CTA.
// xabs_max: max |x'| over the data.
// perm (optional, k ints in global memory, zero-initialised): the sorted order of the previous call.  Centroids
// rarely change their order between Lloyd iterations: when the previous order still sorts them the bitonic network
// (36 barriers for 256 centroids) is skipped.
template <class RT, class TS>
static __device__ void build_region_table(const float *c, int k, float xabs_max, RT *T, TS &S, int *perm = nullptr) {
    const int tid = threadIdx.x;
    // ---- 1. sort (value, id)
    int P = 32;
    while (P < k) P <<= 1;
    float cv = 0.f;
    if (tid < k) {
        cv = c[tid];
        if (cv == 0.f) cv = 0.f;  // -0.0 -> +0.0 (identical distances)
    }
    // max |c|
    float cm = tid < k ? fabsf(cv) : 0.f;
    cm = warp_max_f(cm);
    if (lane_id() == 0) S.warp_f[warp_id()] = cm;
    bool sorted = false;
    if (perm) {
        unsigned long long key = ~0ull;
        if (tid < k) {
            const int pi = perm[tid];
            float pv = c[pi];
            if (pv == 0.f) pv = 0.f;
            key = ((unsigned long long)f2ord(pv) << 32) | (unsigned)pi;
        }
        if (tid < P) S.keys[tid] = key;
        __syncthreads();
        // strictly increasing (value, id) keys over distinct ids = the sorted order
        const bool ok = tid + 1 < k ? S.keys[tid] < S.keys[tid + 1] : true;
        sorted = __syncthreads_and(ok) != 0;
    }
    if (!sorted) {
    if (tid < P) S.keys[tid] = tid < k ? (((unsigned long long)f2ord(cv) << 32) | (unsigned)tid) : ~0ull;
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
    if (perm && tid < k) perm[tid] = (int)(S.keys[tid] & 0xffffffffu);
    }
    float M = xabs_max;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) M = fmaxf(M, S.warp_f[i]);
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
    // ---- 3. decision intervals [A, B] of adjacent pairs
    // Two valid half-widths, the smaller one is used:
    //   global:  |x'|, |c| <= M, so err_a + err_b <= 12u'M^2 =: E2 and delta = E2 / (2(b-a));
    //   pair:    err_a + err_b <= alpha + beta|x'|, alpha = 2u'(a^2+b^2), beta = 4u'(|a|+|b|); with
    //            |x'| <= |mid| + t the pair is decided once 2(b-a)t > alpha + beta(|mid| + t), i.e.
    //            t > (alpha + beta|mid|) / (2(b-a) - beta).
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double up = u * (1.0 + 4.0 * u);
    const double Md = (double)M;
    const double eta = 1e-42;  // absorbs float32 underflow in the products
    const double E2 = 12.0 * up * Md * Md + 2.0 * eta;
    double A = -DBL_MAX, B = DBL_MAX;
    const int npairs = m - 1;
    if (tid < npairs) {
        double a = S.a[tid], b = S.a[tid + 1];
        double mid = 0.5 * (a + b);
        double g2 = 2.0 * (b - a);
        double delta = E2 / g2 * (1.0 + 1e-9);
        double alpha = 2.0 * up * (a * a + b * b) + 2.0 * eta, beta = 4.0 * up * (fabs(a) + fabs(b));
        if (g2 > 2.0 * beta) {  // keep the denominator well conditioned
            double dp = (alpha + beta * fabs(mid)) / (g2 - beta) * (1.0 + 1e-9);
            if (dp < delta) delta = dp;
        }
        A = mid - delta;
        B = mid + delta;
    }
    __syncthreads();
    // lo_{j} = max_{i < j} A_i (j = 1..m-1), hi_j = min_{i >= j} B_i (j = 0..m-2)
    double pmaxA = block_scan_incl<double>(A, [](double x, double y) { return x > y ? x : y; }, S.warp_d);
    if (tid < npairs) S.b[tid] = B;
    __syncthreads();
    double Brev = tid < npairs ? S.b[npairs - 1 - tid] : DBL_MAX;
    double sminB_rev = block_scan_incl<double>(Brev, [](double x, double y) { return x < y ? x : y; }, S.warp_d);
    // breakpoints as float thresholds T with the predicate "x' >= T is right of it":
    //   T_lo(j) = largest float <= lo_j              (j - 1 = tid)   -> S.tlo[tid]
    //   T_hi(j) = smallest float >  ceil32(hi_j)     (j = tid)       -> S.thi[tid]
    if (tid < npairs) {
        S.tlo[tid] = __double2float_rd(pmaxA);
        S.thi[npairs - 1 - tid] = f32_nextup(__double2float_ru(sminB_rev));
    }
    __syncthreads();
    // ---- 4. merge the two sorted breakpoint lists (lo first on ties); region r starts at merged breakpoint r-1
    if (tid < npairs) {
        {   // lo breakpoint of distinct index j = tid + 1
            const float t = S.tlo[tid];
            int lo = 0, hi = npairs;  // #{i : thi[i] < t}
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (S.thi[mid] < t)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            const int p = tid + lo;  // merged position
            T->rstart[p + 1] = t;
            T->rJ1[p + 1] = (short)(tid + 1);
            T->rJ2[p + 1] = (short)(p + 1 - (tid + 1));
        }
        {   // hi breakpoint of distinct index j = tid
            const float t = S.thi[tid];
            int lo = 0, hi = npairs;  // #{j : tlo[j] <= t}
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (S.tlo[mid] <= t)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            const int p = tid + lo;
            T->rstart[p + 1] = t;
            T->rJ2[p + 1] = (short)(tid + 1);
            T->rJ1[p + 1] = (short)(p - tid);
        }
    }
    if (tid == 0) {
        const int R = npairs > 0 ? 2 * npairs + 1 : 1;
        T->R = R;
        T->rstart[0] = -INFINITY;
        T->rJ1[0] = 0;
        T->rJ2[0] = 0;
        T->rstart[R] = INFINITY;
    }
    __syncthreads();
}

}  // namespace nnc
