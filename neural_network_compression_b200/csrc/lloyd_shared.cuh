// lloyd_shared.cuh -- device-side state and helpers shared by the Lloyd kernels (lloyd.cu: per-phase kernels and the
// cooperative one-launch loop; lloyd_fast.cu: the thread-block-cluster loop with its state in shared memory).
#pragma once
#include <stddef.h>

#include "common.cuh"
#include "internal.h"
#include "peer.cuh"
#include "table.cuh"

namespace nnc {

constexpr int LL_TS = 512;   // sorted-array tile (one round of 16 loads per lane in the boundary search)
constexpr int LL_TOP = 2048; // entries of the shared-memory top level of the tile-sample index (loop kernel)
constexpr int LL_LOG = 304;  // per-iteration diagnostics kept for the first LL_LOG iterations

__host__ __device__ __forceinline__ long long llmin2(long long a, long long b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ long long llmax2(long long a, long long b) { return a > b ? a : b; }

struct LloydHeader {
    int k, max_iter, fixed_exp, pad0;
    int rank, world;
    unsigned long long *cand;  // relocation candidates [world][k][2] (see ll_update_kernel phase 1)
    long long n, n_nz, n0, n_tiles;  // n_nz: surviving ELEMENTS of this shard (sum of the entry counts)
    long long n_ent;                 // entries of the sorted array (= n_nz when every entry counts once)
    const unsigned int *cnt;         // multiplicity of every entry (nullptr: all ones)
    const long long *ctile;          // exclusive prefix of the entry counts at tile granularity (with cnt only)
    float mean, tol, xabs_max, pad1;
    double scale, tol_rel;
};
struct LloydDevice : LloydHeader {
    // exact moments of q over all n samples (for the tolerance)
    long long s1;
    unsigned long long s2_lo, s2_hi;
    long long total_q;  // sum of q over the sorted survivors
    // centroids (centred space) by cluster id
    float c[TB_KMAX];
    float c_emit[TB_KMAX];
    float c_save[TB_KMAX];
    long long hist[TB_KMAX];  // code histogram of the final labelling (ll_count_kernel)
    RegionTable tab;
    int perm[TB_KMAX];                // sorted order of the centroids at the previous table build
    long long rpos[2 * TB_KMAX + 2];  // entries before every region boundary
    long long rcnt[2 * TB_KMAX + 2];  // elements before every region boundary (= rpos without multiplicities)
    long long rsum[2 * TB_KMAX + 2];
    // zone partials per distinct index
    long long zW[TB_KMAX], zS[TB_KMAX], zmin[TB_KMAX], zmax[TB_KMAX];
    // per-id partials (pre-relocation) of the previous iteration: label-equality proxy
    long long Wprev[TB_KMAX], Sprev[TB_KMAX];
    // state handed between the phases of the update kernel when they run as separate launches (multi-GPU: the
    // per-cluster partials and the relocation candidates are all-reduced between the phases)
    long long gW[TB_KMAX], gS[TB_KMAX];        // per distinct index: count, fixed-point sum (contiguous: one all-reduce)
    long long gfirst[TB_KMAX], glast[TB_KMAX]; // local member cursors per distinct index
    long long idW[TB_KMAX], idS[TB_KMAX];      // per cluster id, before relocation
    int empt_s[TB_KMAX];
    int zdi_s, n_empty_s, same_s, comm_error;
    long long xbuf[2 * TB_KMAX];  // staging of the in-kernel peer exchange
    // control
    int iter, done, strict, n_reloc, n_iter, pad2;
    unsigned int gbar;  // arrival counter of the grid barrier (ll_loop_kernel)
    int gbail;          // a CTA gave up waiting at the grid barrier
    // per-iteration log (diagnostics): zone elements, zone groups, distinct centroids, empty clusters
    long long logZ[LL_LOG];
    int logG[LL_LOG], logM[LL_LOG], logE[LL_LOG];
    unsigned int logT[LL_LOG][4];  // loop kernel: ns spent in search / zone / update / table (+ their barriers) per iteration
};


// entries / elements of the sorted survivors with fl(x - mean) < t, and the sum of q over them; one warp per boundary.
// top (optional, shared memory): top[i] = samp[i * top_step], i < top_n -- the first, coarse level of the search without a
// trip to L2.
struct SearchConst {  // per-launch constants of the search (hoisted out of the iterations by the loop kernel)
    const float *ks;
    const unsigned int *cnt;
    const float *samp;
    const long long *ptile, *ctile;
    long long n_ent, n_tiles;
    float mean;
    double scale;
    const float *top;
    long long top_step;
    int top_n;
    int vec_ok;     // ks / cnt 16-byte aligned: the tile scan uses 128-bit loads
    float scale_f;  // 2^(30-E) as a float (0: not representable, use `scale`)
};

// fixed-point image without float64 arithmetic: x' * 2^s is exact in float32 (a power-of-two scaling below 2^30), so
// rint of the float product equals rint of the double product.  scale_f = 2^s as a float, or 0 when 2^s is not a normal
// float (absurd data ranges): then the float64 expression is used.
__device__ __forceinline__ long long fixed_qf(float xc, float scale_f, double scale) {
    return scale_f != 0.f ? (long long)__float2int_rn(__fmul_rn(xc, scale_f)) : fixed_q(xc, scale);
}
__device__ __forceinline__ float4 ld_vol_f4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_vol_u4(const unsigned int *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// first tile whose first key FAILS the predicate fl(key - mean) < t (n_tiles when none fails): 32-ary search over the
// shared-memory top level, then over the tile samples
__device__ __forceinline__ long long warp_locate_tile(const SearchConst &C, float t) {
    const int lane = lane_id();
    const float mean = C.mean;
    long long lo = 0, hi = C.n_tiles;  // the answer lies in [lo, hi]
    if (C.top) {  // coarse level: number of top samples that satisfy the predicate (they are a prefix)
        int a = 0, b = C.top_n;  // first top index failing lies in [a, b]
        while (a < b) {
            const int span = b - a, step = (span + 31) >> 5;
            const int cs = a + lane * step;
            const int last = min(cs + step, b) - 1;
            const bool p = cs < b ? (fsub(C.top[last], mean) < t) : false;
            const int c = __popc(__ballot_sync(0xffffffffu, p));
            const int na = min(a + c * step, b);
            if (na >= b) {
                a = b;
                break;
            }
            b = min(na + step, b) - 1;
            a = na;
        }
        // top samples 0 .. a-1 satisfy the predicate, sample a (if any) fails
        if (a == 0) {
            lo = 0;
            hi = 0;
        } else {
            lo = (long long)(a - 1) * C.top_step + 1;  // tile (a-1)*step satisfies: the first failing tile is after it
            hi = a < C.top_n ? (long long)a * C.top_step : C.n_tiles;
        }
    }
    while (lo < hi) {
        long long span = hi - lo;
        long long step = (span + 31) >> 5;
        long long cs = lo + (long long)lane * step;  // chunk start
        bool inr = cs < hi;
        long long last = llmin2(cs + step, hi) - 1;
        bool p = inr ? (fsub(C.samp[last], mean) < t) : false;
        unsigned b = __ballot_sync(0xffffffffu, p);
        int c = __popc(b);
        long long nlo = llmin2(lo + (long long)c * step, hi);
        long long nhi = llmin2(nlo + step, hi) - 1;
        if (nlo >= hi) {
            lo = hi;
            break;
        }
        lo = nlo;
        hi = nhi;
    }
    return lo;
}

__device__ __forceinline__ float ld_vol_f1(const float *p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ long long ld_vol_s64(const long long *p) {
    long long v;
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// entries / elements of the sorted survivors with fl(x - mean) < t, and the sum of q over them; one warp per boundary.
// hint (optional, one per boundary slot of the calling warp, -1 initially): the answer of warp_locate_tile of the previous
// call for this boundary.  Boundaries move little between Lloyd iterations: the hinted tile is loaded SPECULATIVELY
// together with the two tile samples that bracket it; when they still bracket t (almost always) the search is ONE memory
// round trip instead of four.  All loads are volatile asm so that they are issued before the first use (the compiler
// otherwise sinks every load to its use: 16 serialised round trips).
__device__ __forceinline__ void warp_boundary_search(const SearchConst &C, float t, long long &pos_out, long long &cnt_out,
                                                     long long &sum_out, long long *hint = nullptr) {
    const int lane = lane_id();
    const float mean = C.mean;
    float4 a0, a1, a2, a3;
    uint4 c0 = make_uint4(1u, 1u, 1u, 1u), c1 = c0, c2 = c0, c3 = c0;
    long long psum = 0, pcnt = 0;
    auto full_tile = [&](long long tile) { return C.vec_ok && tile >= 0 && (tile + 1) * LL_TS <= C.n_ent; };
    auto load_tile = [&](long long tile) {  // a complete, 16-byte aligned tile: 128-bit loads, all in flight
        const long long base = tile * LL_TS;
        const float *pk = C.ks + base + 4 * lane;
        a0 = ld_vol_f4(pk);
        a1 = ld_vol_f4(pk + 128);
        a2 = ld_vol_f4(pk + 256);
        a3 = ld_vol_f4(pk + 384);
        if (C.cnt) {
            const unsigned int *pc = C.cnt + base + 4 * lane;
            c0 = ld_vol_u4(pc);
            c1 = ld_vol_u4(pc + 128);
            c2 = ld_vol_u4(pc + 256);
            c3 = ld_vol_u4(pc + 384);
            pcnt = ld_vol_s64(C.ctile + tile);
        }
        psum = ld_vol_s64(C.ptile + tile);
    };
    long long lo = -1;
    bool loaded = false;
    if (hint) {
        const long long g = *hint;
        if (g >= 0 && g <= C.n_tiles) {
            const bool spec = g >= 1 && full_tile(g - 1);
            if (spec) load_tile(g - 1);
            const float below = g > 0 ? ld_vol_f1(C.samp + (g - 1)) : -INFINITY;
            const float at = g < C.n_tiles ? ld_vol_f1(C.samp + g) : INFINITY;
            const bool ok = (g == 0 || fsub(below, mean) < t) && (g == C.n_tiles || !(fsub(at, mean) < t));
            if (ok) {
                lo = g;
                loaded = spec;
            }
        }
    }
    if (lo < 0) lo = warp_locate_tile(C, t);
    if (hint) *hint = lo;
    if (lo == 0) {
        pos_out = 0;
        cnt_out = 0;
        sum_out = 0;
        return;
    }
    const long long tile = lo - 1;
    const long long base = tile * LL_TS;
    int npos = 0;
    long long acc = 0, cacc = 0;
    if (!loaded && full_tile(tile)) {
        load_tile(tile);
        loaded = true;
    }
    if (loaded) {
        auto one = [&](float x, unsigned int c) {
            const float xc = fsub(x, mean);
            if (xc < t) {
                acc += fixed_qf(xc, C.scale_f, C.scale) * (long long)c;
                cacc += c;
                npos += 1;
            }
        };
        one(a0.x, c0.x); one(a0.y, c0.y); one(a0.z, c0.z); one(a0.w, c0.w);
        one(a1.x, c1.x); one(a1.y, c1.y); one(a1.z, c1.z); one(a1.w, c1.w);
        one(a2.x, c2.x); one(a2.y, c2.y); one(a2.z, c2.z); one(a2.w, c2.w);
        one(a3.x, c3.x); one(a3.y, c3.y); one(a3.z, c3.z); one(a3.w, c3.w);
    } else {  // the last (partial) tile, or an unaligned array
        psum = C.ptile[tile];
        if (C.cnt) pcnt = C.ctile[tile];
        for (int j = 0; j < LL_TS / 32; ++j) {
            const long long i = base + j * 32 + lane;
            if (i < C.n_ent) {
                const float xc = fsub(C.ks[i], mean);
                if (xc < t) {
                    const unsigned int c = C.cnt ? C.cnt[i] : 1u;
                    acc += fixed_qf(xc, C.scale_f, C.scale) * (long long)c;
                    cacc += c;
                    npos += 1;
                }
            }
        }
    }
    npos = warp_sum_i(npos);
    acc = warp_sum_ll(acc);
    pos_out = base + npos;
    cnt_out = C.cnt ? pcnt + warp_sum_ll(cacc) : pos_out;
    sum_out = psum + acc;
}

__device__ __forceinline__ SearchConst search_const(const LloydDevice *st, const float *ks, const float *samp, const long long *ptile) {
    SearchConst C;
    C.ks = ks;
    C.cnt = st->cnt;
    C.samp = samp;
    C.ptile = ptile;
    C.ctile = st->ctile;
    C.n_ent = st->n_ent;
    C.n_tiles = st->n_tiles;
    C.mean = st->mean;
    C.scale = st->scale;
    C.top = nullptr;
    C.top_step = 1;
    C.top_n = 0;
    C.vec_ok = ((reinterpret_cast<uintptr_t>(ks) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(st->cnt) & 15u) == 0);
    const int sh = 30 - st->fixed_exp;
    C.scale_f = (sh >= -126 && sh <= 127) ? __int_as_float((sh + 127) << 23) : 0.f;
    return C;
}
// NumPy pairwise sum of a small float32 array (numpy/_core/src/umath/loops_utils.h.src), single thread.
static __device__ float np_pairwise_small(const float *a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = fadd(r, a[i]);
        return r;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = fadd(r[j], a[i + j]);
        float res = fadd(fadd(fadd(r[0], r[1]), fadd(r[2], r[3])), fadd(fadd(r[4], r[5]), fadd(r[6], r[7])));
        for (; i < n; ++i) res = fadd(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return fadd(np_pairwise_small(a, n2), np_pairwise_small(a + n2, n - n2));
}

struct FarKey {  // priority of a relocation candidate: larger is farther
    uint32_t d2, gap, ordx;
};
__device__ __forceinline__ bool far_before(const FarKey &a, const FarKey &b) {
    if (a.d2 != b.d2) return a.d2 > b.d2;
    if (a.gap != b.gap) return a.gap > b.gap;
    return a.ordx > b.ordx;
}
__device__ __forceinline__ FarKey far_key(float xc, float c) {
    float t = fsub(xc, c);
    float d2 = fmul(t, t);
    FarKey k;
    k.d2 = __float_as_uint(d2);
    uint32_t ox = f2ord(xc), oc = f2ord(c);
    k.gap = ox > oc ? ox - oc : oc - ox;
    k.ordx = ox;
    return k;
}

// lloyd_fast.cu: the loop as one thread-block cluster with its state in shared memory (k <= 512)
bool lloyd_fast_applicable(int k);
void lloyd_fast_launch(nnc_ctx *ctx, LloydDevice *st, const float *d_sorted, const float *samp, const long long *ptile,
                       const float *d_init, int want_hist, const PeerComm &pc, int k);

}  // namespace nnc
