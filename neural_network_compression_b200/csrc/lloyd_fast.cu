// lloyd_fast.cu -- the Lloyd loop as ONE thread-block cluster with its whole state in shared memory.
//
// Same arithmetic as lloyd.cu (which restates sklearn's _kmeans_single_lloyd / lloyd_iter_chunked_dense /
// _relocate_empty_clusters_dense / _average_centers / _center_shift, sklearn/cluster/_kmeans.py:630-758,
// _k_means_lloyd.pyx:23-218, _k_means_common.pyx:167-311, reached through KMeans(...).fit in
// neural_network_compression/common/utility.py:237-238) -- bit-identical results, checked against the cooperative
// loop kernel and the oracle.  What changes is where the iteration state lives and how the CTAs synchronise.
//
// An iteration on the sorted survivors is latency, not bandwidth: <= m - 1 pairs of boundary searches, ~0.01 % of the
// entries evaluated with the float32 label rule, and a k-element update.  The cooperative kernel (ll_loop_kernel) spends
// ~43 us per iteration on that: the region table, the searched positions and the per-cluster partials live in global
// memory (every phase starts with dependent L2 round trips), CTA 0 runs the serial phases, and three software grid
// barriers of 1.6 us separate the phases.  Here:
//   * one cluster of up to 16 CTAs x 512 threads (a non-portable cluster size; 8 x 1024 is the fallback), one CTA per four
//     boundary pairs: 16 CTAs for an 8-bit codebook, one for a 2-bit one;
//   * every CTA builds the region table REDUNDANTLY in its own shared memory from its copy of the centroids
//     (deterministic: identical tables, no broadcast of 14 KB);
//   * one warp per PAIR of region boundaries locates both, loads the tile(s) under them once, takes both prefix triples
//     and evaluates the zone between them from the same registers (warp_pair_search); results go to CTA 0's shared memory
//     by plain distributed-shared-memory stores, one writer per slot (remote atomics lost updates: never used);
//   * what the pair search cannot evaluate (multi-candidate zones, very wide zones) goes to a chunk pass over all warps;
//     CTA 0 publishes whether there is anything left, and usually there is not;
//   * phases are separated by the hardware cluster barrier instead of a spinning grid barrier;
//   * CTA 0 updates the centroids out of shared memory only -- empty clusters are refilled by a block-wide tournament over
//     preloaded candidate streams; on several GPUs the exact (count, sum) pairs are all-reduced with tagged words over
//     NVLink peer memory (peer.cuh) -- and the other CTAs pull the k new centroids over DSMEM.
// k <= LF_KMAX (512: 8-bit codebooks and the 257-centroid density init); larger k keeps ll_loop_kernel.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "lloyd_shared.cuh"

namespace cg = cooperative_groups;

namespace nnc {

constexpr int LF_KMAX = 512;
constexpr int LF_CL_MAX = 16;  // CTAs per cluster: 16 (non-portable size, one GPC) x 512 threads, or 8 x 1024
constexpr int LF_R = 2 * LF_KMAX + 2;

// scratch of np_pairwise_warp
struct NpWarpScratch {
    short off[16], len[16];
    float val[32];              // [0, 16): the leaves' sums, [16, 32): the internal nodes of the fold
    short fl[16], fr[16];       // fold program: node t = val[fl[t]] + val[fr[t]] (built once: the tree depends only on n)
    int n_ops, root;
};

// relocation: every distinct index has two candidate streams -- its members below its centroid walked from the left end
// (stream 2 i) and its members at or above it walked from the right end (stream 2 i + 1); along each stream the FarKey
// decreases strictly.  The first LF_POP_D members of every stream are PRELOADED by all threads at once (one memory round
// trip), so that the serial pop chain runs out of shared memory; a stream that is drained deeper refills itself (rare).
constexpr int LF_POP_D = 6;      // preloaded members per stream for m <= LF_KMAX / 2 distinct centroids (3 above)
constexpr int LF_POP_ENT = 2560;  // preloaded members behind the heads, all streams: [stream * (D - 1) + j] = member j + 1 (key,
                                  // multiplicity); they live in the chunk-slot array, which is dead during the relocation
struct PopState {
    uint4 hk[2 * LF_KMAX];    // head of the stream: FarKey (d2, gap, ordx) and the samples left in it (0: stream empty)
    int next_off[2 * LF_KMAX];  // entries already walked from the stream's end (where a refill continues); -1: nothing left
    unsigned char head[2 * LF_KMAX], depth[2 * LF_KMAX];  // preloaded members: the one in hk, how many there are
};

struct FastUpdate {  // scratch of the update step (CTA 0)
    long long Wd[LF_KMAX], Sd[LF_KMAX];       // per distinct index
    long long W[LF_KMAX], S[LF_KMAX];         // per cluster id
    long long first[LF_KMAX], last[LF_KMAX];  // member cursors per distinct index
    // two-candidate zones gathered from the chunk slots: what a distinct index gets as the LEFT candidate (a) of the zone to
    // its right and as the RIGHT candidate (b) of the zone to its left (a pair of candidates occurs in one region only)
    union {  // the gather scratch is dead when relocation starts
        struct {
            long long aW[LF_KMAX], aS[LF_KMAX], af[LF_KMAX], al[LF_KMAX];
            long long bW[LF_KMAX], bS[LF_KMAX], bf[LF_KMAX], bl[LF_KMAX];
        };
        PopState pop;
    };
    float raw[LF_KMAX], cnew[LF_KMAX], sq[LF_KMAX];
    int empt[LF_KMAX];
    float far_x[LF_KMAX];
    int far_old[LF_KMAX];
    float red_d[32];
    int red_i[32], red_id[32];
    uint32_t rk_d2[32], rk_gap[32], rk_ord[32];
    int rk_who[32];
    unsigned long long red_w[32];
    int n_empty, zdi, same, winner, stop, strict;
    long long zero_left;
};
constexpr int LF_CH = 128;      // entries per zone chunk (one warp: four rounds of 32 with all loads in flight)
constexpr int LF_MAXCH = 1024;  // chunk slots in CTA 0 (two-candidate zones; more chunks take the generic path)
struct ZoneSlot {               // result of one two-candidate chunk, written by exactly one warp of the cluster
    long long Wb, Sb, Wt, St;   // (count, fixed-point sum) of candidate b and of both candidates
    unsigned int ends;          // first / last member of a and of b as chunk offsets + 1 (0: none), one byte each
    unsigned int pad;
};
struct FusedZone {             // a two-candidate zone evaluated by the warp that searched its two boundaries
    long long Wb, Sb;           // (count, fixed-point sum) of the members of candidate b
    unsigned short fa, la, fb, lb;  // first / last member of a and of b as offsets from the zone's first entry, + 1 (0: none)
    unsigned int tag;           // number of the E-step that wrote it (anything else: the zone is left to the chunk pass)
    unsigned int pad;
};
struct FastZone {
    long long rp[LF_R];  // copy of CTA 0's region positions
    int s_warp[32];
    unsigned char hd[LF_R];  // the region is a zone that the search pass has evaluated
};
struct FastConst {  // read once from the LloydDevice header
    int k, rank, world, max_iter;
    long long n, n_nz, n0, n_ent;
    const unsigned int *cnt;
    unsigned long long *cand;
    float mean, xabs_max;
    double scale, inv_scale;  // 2^(30-E) and its (exact) reciprocal
};

struct FastSmem {
    FastConst K;            // kernel-lifetime constants (same in every CTA)
    SearchConst C;
    long long hint[32][10]; // per warp and boundary slot: the tile found by the previous search (-1: none)
    int merge_cur[64];      // relocation: read cursors of the ranks' candidate lists

    RegionTableT<LF_KMAX> tab;  // built redundantly by every CTA
    float c[LF_KMAX];           // centred centroids by cluster id
    int perm[LF_KMAX];          // sorted order of the previous table build
    float top[LL_TOP];          // coarse level of the tile-sample index
    int done, iter, strict, n_reloc, n_iter, comm_error;
    union {
        TableScratchT<LF_KMAX> tb;
        FastUpdate up;
        FastZone zn;
    } u;
    // searched region boundaries: only CTA 0's copy is used; the other CTAs write it through DSMEM
    long long rpos[LF_R], rcnt[LF_R], rsum[LF_R];
    // zone partials per distinct index of THIS CTA (CTA 0 pulls all of them in the update step)
    unsigned long long zW[LF_KMAX];
    long long zS[LF_KMAX], zmin[LF_KMAX], zmax[LF_KMAX];
    NpWarpScratch np;         // leaf list of NumPy's pairwise sum over k values (built once) + its scratch
    int np_leaves;
    int cpre[LF_R];           // chunks of the zones before region r (SAFE regions: none); same numbers in every CTA
    alignas(16) ZoneSlot slot[LF_MAXCH];  // CTA 0: per-chunk results of the two-candidate zones (remote plain stores, one writer each)
    FusedZone fz[LF_KMAX];    // CTA 0: zones evaluated in the search pass, by the distinct index of candidate a (one writer each)
    unsigned int zflag;       // CTA 0: (E-step number << 1) | "a non-empty zone is left to the chunk pass"
    long long Wprev[LF_KMAX], Sprev[LF_KMAX];
    float c_emit[LF_KMAX];
    long long xbuf[2 * LF_KMAX];  // staging of the peer exchange
};

// (FastConst and SearchConst are placed in shared memory by the kernel: as kernel-lifetime registers they were spilled, and
// with 214 KB of the SM's 256 KB configured as shared memory the spills miss the remaining L1 and go to L2)
static_assert(sizeof(FastSmem) <= 227 * 1024, "FastSmem must fit the 227 KB of shared memory a CTA can opt into");


// NumPy's pairwise float32 sum (numpy/_core/src/umath/loops_utils.h.src) of a[0, n), n <= 1024, by ONE WARP: the leaves
// (<= 128 elements: 8 strided accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n % 8 tail) are
// summed by groups of 8 lanes, four leaves at a time; lane 0 folds them up the recursion tree.  Bit-identical with
// np_pairwise_small.  scratch: 16 shorts + 16 shorts + 16 floats.
static __device__ int np_leaf_list(int o, int n, NpWarpScratch &W, int cnt) {
    if (n <= 128) {
        W.off[cnt] = (short)o;
        W.len[cnt] = (short)n;
        return cnt + 1;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    cnt = np_leaf_list(o, n2, W, cnt);
    return np_leaf_list(o + n2, n - n2, W, cnt);
}
static __device__ float np_fold_leaves(int n, const NpWarpScratch &W, int &idx) {
    if (n <= 128) return W.val[idx++];
    int n2 = n / 2;
    n2 -= n2 % 8;
    const float l = np_fold_leaves(n2, W, idx);
    const float r = np_fold_leaves(n - n2, W, idx);
    return fadd(l, r);
}
// the fold over the leaves as a straight-line program (the recursion costs a stack frame in local memory -- L2 latency in
// this kernel -- on every call; the program is built once per kernel).  Returns the index of the value holding the sum.
static __device__ int np_fold_prog(int n, NpWarpScratch &W, int &leaf, int &n_ops) {
    if (n <= 128) return leaf++;
    int n2 = n / 2;
    n2 -= n2 % 8;
    const int l = np_fold_prog(n2, W, leaf, n_ops);
    const int r = np_fold_prog(n - n2, W, leaf, n_ops);
    W.fl[n_ops] = (short)l;
    W.fr[n_ops] = (short)r;
    return 16 + n_ops++;
}
// n_leaves < 0: build the leaf list first (it only depends on n: callers with a fixed n build it once and pass the count)
__device__ float np_pairwise_warp(const float *a, int n, NpWarpScratch &W, int n_leaves = -1) {  // one whole warp; result in lane 0
    const int lane = lane_id(), grp = lane >> 3, j = lane & 7;
    const bool n_leaves_given = n_leaves >= 0;
    if (n_leaves < 0) {
        if (lane == 0) n_leaves = np_leaf_list(0, n, W, 0);
        n_leaves = __shfl_sync(0xffffffffu, n_leaves, 0);
        __syncwarp();
    }
    for (int l0 = 0; l0 < n_leaves; l0 += 4) {
        const int l = l0 + grp;
        float r = 0.f;
        const bool have = l < n_leaves;
        const int o = have ? W.off[l] : 0, sz = have ? W.len[l] : 0;
        if (sz >= 8) {
            r = a[o + j];
            const int full = sz - (sz % 8);
#pragma unroll 8
            for (int i = 8; i < full; i += 8) r = fadd(r, a[o + i + j]);  // (unrolled: the loads run ahead of the add chain)
        }
        r = fadd(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = fadd(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = fadd(r, __shfl_xor_sync(0xffffffffu, r, 4));
        if (have && j == 0) {
            if (sz >= 8) {
                for (int i = sz - (sz % 8); i < sz; ++i) r = fadd(r, a[o + i]);
            } else {  // n < 8: plain sequential sum starting from 0
                r = 0.f;
                for (int i = 0; i < sz; ++i) r = fadd(r, a[o + i]);
            }
            W.val[l] = r;
        }
    }
    __syncwarp();
    float tot = 0.f;
    if (lane == 0) {
        if (n_leaves_given) {  // the fold program was built with the leaf list
            for (int t = 0; t < W.n_ops; ++t) W.val[16 + t] = fadd(W.val[W.fl[t]], W.val[W.fr[t]]);
            tot = W.val[W.root];
        } else {
            int idx = 0;
            tot = np_fold_leaves(n, W, idx);
        }
    }
    return tot;
}

// label (distinct index) of the sorted survivor at position p whose value x is already in a register
__device__ __forceinline__ int fast_label_of(const FastSmem &S, float mean, long long p, float x) {
    const RegionTableT<LF_KMAX> &T = S.tab;
    int lo = 0, hi = T.R;  // largest r with rpos[r] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (S.rpos[mid] <= p)
            lo = mid;
        else
            hi = mid;
    }
    if (T.rJ1[lo] == T.rJ2[lo]) return T.rJ1[lo];
    return zone_argmin(fsub(x, mean), T.dv, T.dcn, T.down, T.rJ2[lo], T.rJ1[lo]);
}

// label (distinct index) of the sorted survivor at position p
__device__ __forceinline__ int fast_label_at(const FastSmem &S, const float *ks, float mean, long long p) {
    const RegionTableT<LF_KMAX> &T = S.tab;
    int lo = 0, hi = T.R;  // largest r with rpos[r] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (S.rpos[mid] <= p)
            lo = mid;
        else
            hi = mid;
    }
    if (T.rJ1[lo] == T.rJ2[lo]) return T.rJ1[lo];
    return zone_argmin(fsub(ks[p], mean), T.dv, T.dcn, T.down, T.rJ2[lo], T.rJ1[lo]);
}

// ---- zone step: the float32 label rule for the few entries inside the zones.
// The zones are cut into chunks of LF_CH entries and the chunks are dealt round robin to ALL warps of the cluster (zones
// between adjacent centroids hold 40 .. 900 entries: one warp per zone left most warps idle behind the largest ones).
//   two-candidate zones (the normal case): (count, sum) of the two candidates in registers, first / last members from
//       ballots; the chunk's result goes to ITS slot in CTA 0's shared memory with plain remote stores -- one writer per
//       slot, no atomics (generic-address 64-bit atomics on another CTA's shared memory lost updates under contention on
//       sm_100a and are not used anywhere);
//   zones with more candidates (near-duplicate centroids after a relocation), and chunks beyond the slot array: the
//       generic leader loop into THIS CTA's partials (local atomics), which CTA 0 pulls over DSMEM in the update step.
__device__ __forceinline__ bool zone_is_generic(const RegionTableT<LF_KMAX> &T, int r) { return T.rJ1[r] > T.rJ2[r] + 1; }

// one warp: entries [lo, hi) of a region, any number of candidates; accumulates into the CTA's partials
__device__ __noinline__ void fast_zone_generic(FastSmem &S, const FastConst &K, const SearchConst &C, const float *__restrict__ ks,
                                               int J2, int J1, long long lo, long long hi) {
    const RegionTableT<LF_KMAX> &T = S.tab;
    const int lane = lane_id();
    const unsigned int *__restrict__ ecnt = K.cnt;
    for (long long b = lo; b < hi; b += 32) {
        const long long p = b + lane;
        const bool valid = p < hi;
        int di = -1;
        long long q = 0, c = 0;
        if (valid) {
            const float xc = fsub(ks[p], K.mean);
            di = zone_argmin(xc, T.dv, T.dcn, T.down, J2, J1);
            c = ecnt ? (long long)ecnt[p] : 1ll;
            q = fixed_qf(xc, C.scale_f, K.scale) * c;
        }
        unsigned active = __ballot_sync(0xffffffffu, valid);
        while (active) {
            const int leader = __ffs(active) - 1;
            const int L = __shfl_sync(0xffffffffu, di, leader);
            const bool mine = valid && di == L;
            const unsigned grp = __ballot_sync(0xffffffffu, mine);
            const long long sq = warp_sum_ll(mine ? q : 0);
            const long long sc = ecnt ? warp_sum_ll(mine ? c : 0) : (long long)__popc(grp);
            if (lane == leader) {
                atomicAdd(&S.zW[L], (unsigned long long)sc);
                atomicAdd((unsigned long long *)&S.zS[L], (unsigned long long)sq);
                atomicMin(&S.zmin[L], b + (__ffs(grp) - 1));
                atomicMax(&S.zmax[L], b + (31 - __clz(grp)));
            }
            active &= ~grp;
        }
    }
}

// ---- search pass: one warp takes TWO consecutive region boundaries t1 <= t2 -- in a regular table the two ends of the zone
// between two adjacent centroids.  A zone holds ~100 sorted entries, so both boundaries lie in the same 512-entry tile (or
// in two adjacent ones): the tiles are located for both thresholds at once, loaded once, both prefix (position, count,
// sum) triples come out of the same registers, and the float32 label rule of the zone's entries is evaluated on the spot.
// That is two memory round trips (tile samples, tiles) for what were two boundary searches of three round trips each, a
// cluster barrier and a second pass over the zone's entries.  Returns whether the zone was evaluated; if not (zone wider
// than two tiles, ragged last tile, unaligned arrays) it is left to the chunk pass.
// the plain search as a CALL: it is the cold path here, and inlined it would share (and blow) the register budget of the
// pair search
__device__ __noinline__ void boundary_search_call(const SearchConst &C, float t, long long *out3, long long *hint) {
    long long pos, cn, sum;
    warp_boundary_search(C, t, pos, cn, sum, hint);
    out3[0] = pos;
    out3[1] = cn;
    out3[2] = sum;
}
struct PairOut {
    long long pos1, cnt1, sum1, pos2, cnt2, sum2;
    long long Wb, Sb;
    unsigned int fa, la, fb, lb;  // offsets from pos1, + 1 (0: none)
};
// first tiles whose first key FAILS fl(key - mean) < t1 / < t2 (t1 <= t2), exactly: the shared-memory top level gives a
// bracket of <= top_step tiles, and ALL tile samples of the bracket are read in one round trip (a boundary moves by tens of
// tiles per Lloyd iteration on a 2^30-element tensor: a tile remembered from the previous iteration misses most of the time)
constexpr int LP_PER = 12;  // samples per lane in the final round: brackets of up to 384 tiles
__device__ __forceinline__ void warp_locate_pair(const SearchConst &C, float t1, float t2, long long &g1, long long &g2) {
    const int lane = lane_id();
    const float mean = C.mean;
    int a_top = -1;  // first failing top-level sample of the previous threshold (-1: none yet)
    auto bracket = [&](float t, long long &lo, long long &hi) {
        lo = 0;
        hi = C.n_tiles;  // the answer lies in [lo, hi]
        if (C.top) {
            int a = 0, b = C.top_n;  // first top index failing lies in [a, b]
            // thresholds come in ascending order: the second one usually falls between the same two top-level samples
            if (a_top >= 0 && (a_top == C.top_n || !(fsub(C.top[a_top], mean) < t))) a = b = a_top;
            else if (a_top > 0) a = a_top;  // (samples below a_top satisfy the smaller threshold, hence this one)
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
            a_top = a;
            if (a == 0) {
                hi = 0;
            } else {
                lo = (long long)(a - 1) * C.top_step + 1;
                hi = a < C.top_n ? (long long)a * C.top_step : C.n_tiles;
            }
        }
        while (hi - lo > 32 * LP_PER) {  // huge arrays only: 32-ary rounds over the tile samples
            const long long span = hi - lo, step = (span + 31) >> 5;
            const long long cs = lo + (long long)lane * step;
            const long long last = llmin2(cs + step, hi) - 1;
            const bool p = cs < hi ? (fsub(C.samp[last], mean) < t) : false;
            const int c = __popc(__ballot_sync(0xffffffffu, p));
            const long long nlo = llmin2(lo + (long long)c * step, hi);
            if (nlo >= hi) {
                lo = hi;
                break;
            }
            hi = llmin2(nlo + step, hi) - 1;
            lo = nlo;
        }
    };
    long long lo1, hi1, lo2, hi2;
    bracket(t1, lo1, hi1);
    bracket(t2, lo2, hi2);
    float sv[LP_PER];
    auto load = [&](long long lo, long long hi) {
#pragma unroll
        for (int i = 0; i < LP_PER; ++i) {
            const long long idx = lo + lane + 32 * i;
            sv[i] = idx < hi ? ld_vol_f1(C.samp + idx) : INFINITY;
        }
    };
    auto count = [&](float t) {
        int c = 0;
#pragma unroll
        for (int i = 0; i < LP_PER; ++i) c += fsub(sv[i], mean) < t;  // (the padding +inf never counts)
        return __reduce_add_sync(0xffffffffu, c);
    };
    load(lo1, hi1);
    g1 = lo1 + count(t1);
    if (lo2 != lo1 || hi2 != hi1) load(lo2, hi2);  // the thresholds straddle a top-level sample (rare)
    g2 = lo2 + count(t2);
}

__device__ __forceinline__ bool warp_pair_search(const SearchConst &C, float t1, float t2, bool want_zone, float va, float na,
                                                 float vb, float nb, bool b_wins_ties, PairOut &o, long long *pp = nullptr) {
    const int lane = lane_id();
    const float mean = C.mean;
    long long g1, g2;
    if (pp) pp[0] = clock64();
    warp_locate_pair(C, t1, t2, g1, g2);
    if (pp) pp[1] = clock64();
    // the fast path needs 16-byte aligned arrays and a tile under the first boundary (g1 = 0: it lies before all data)
    if (!(C.vec_ok && g1 >= 1)) {
        if (pp) pp[5] = g1 < 1 ? 1 : 2;
        long long r3[3];
        boundary_search_call(C, t1, r3, nullptr);
        o.pos1 = r3[0], o.cnt1 = r3[1], o.sum1 = r3[2];
        boundary_search_call(C, t2, r3, nullptr);
        o.pos2 = r3[0], o.cnt2 = r3[1], o.sum2 = r3[2];
        return false;
    }
    float4 a0, a1, a2, a3, b0, b1, b2, b3;
    uint4 c0 = make_uint4(1u, 1u, 1u, 1u), c1 = c0, c2 = c0, c3 = c0, d0 = c0, d1 = c0, d2 = c0, d3 = c0;
    long long psA = 0, pcA = 0, psB = 0, pcB = 0;
    const bool two = g2 != g1;
    constexpr int LP_MID = 2;  // tiles between A and B that are still walked here, one after the other; wider zones (near-duplicate
                               // centroids after a relocation) go to the chunk pass, where all warps share them
    if (g2 - g1 - 1 > LP_MID) want_zone = false;  // an enormous zone: left to the chunk pass
    if (pp) pp[5] = (want_zone ? 0 : 3) | ((g2 - g1) << 8);
    const long long baseA = (g1 - 1) * LL_TS, baseB = (g2 - 1) * LL_TS;
    // the last tile of the array may be ragged: what lies beyond n_ent reads as (+inf, count 0) -- above every threshold
    auto ldk = [&](long long i) {
        if (i + 3 < C.n_ent) return ld_vol_f4(C.ks + i);
        return make_float4(i < C.n_ent ? C.ks[i] : INFINITY, i + 1 < C.n_ent ? C.ks[i + 1] : INFINITY,
                           i + 2 < C.n_ent ? C.ks[i + 2] : INFINITY, INFINITY);
    };
    auto ldc = [&](long long i) {
        if (i + 3 < C.n_ent) return ld_vol_u4(C.cnt + i);
        return make_uint4(i < C.n_ent ? C.cnt[i] : 0u, i + 1 < C.n_ent ? C.cnt[i + 1] : 0u, i + 2 < C.n_ent ? C.cnt[i + 2] : 0u, 0u);
    };
    const long long ia = baseA + 4 * lane, ib = baseB + 4 * lane;
    {
        a0 = ldk(ia);
        a1 = ldk(ia + 128);
        a2 = ldk(ia + 256);
        a3 = ldk(ia + 384);
        if (two) {
            b0 = ldk(ib);
            b1 = ldk(ib + 128);
            b2 = ldk(ib + 256);
            b3 = ldk(ib + 384);
        }
        if (C.cnt) {
            c0 = ldc(ia);
            c1 = ldc(ia + 128);
            c2 = ldc(ia + 256);
            c3 = ldc(ia + 384);
            pcA = ld_vol_s64(C.ctile + (g1 - 1));
            if (two) pcB = ld_vol_s64(C.ctile + (g2 - 1));
        }
        psA = ld_vol_s64(C.ptile + (g1 - 1));
        if (two) psB = ld_vol_s64(C.ptile + (g2 - 1));
    }
    if (pp) pp[2] = clock64() + (__float_as_int(a0.x) & __float_as_int(a3.w) & (int)psA & 0);  // (after the tile has arrived)
    int n1 = 0, n2 = 0;
    long long q1 = 0, q2 = 0, w1 = 0, w2 = 0, Wb = 0, Sb = 0;
    unsigned int ma = 0, mb = 0;  // per lane: which of its 32 slots (16 of tile A, 16 of tile B) hold a member of a / of b
    const bool single = !two;
    auto one = [&](float x, unsigned int c, int slot, bool inB) {
        const float xc = fsub(x, mean);
        const long long q = fixed_qf(xc, C.scale_f, C.scale) * (long long)c;
        const bool lt1 = !inB && xc < t1, lt2 = (inB || single) ? xc < t2 : true;
        if (lt1) {
            q1 += q;
            w1 += c;
            n1 += 1;
        }
        if ((inB || single) && lt2) {
            q2 += q;
            w2 += c;
            n2 += 1;
        }
        if (want_zone && !lt1 && lt2) {  // inside the zone: the float32 rule between its two candidates
            const float m2x = fmul(-2.0f, xc);
            const float da = skl_dist(m2x, va, na), db = skl_dist(m2x, vb, nb);
            const bool isb = db < da || (db == da && b_wins_ties);
            if (isb) {
                Wb += c;
                Sb += q;
                mb |= 1u << slot;
            } else {
                ma |= 1u << slot;
            }
        }
    };
    one(a0.x, c0.x, 0, false); one(a0.y, c0.y, 1, false); one(a0.z, c0.z, 2, false); one(a0.w, c0.w, 3, false);
    one(a1.x, c1.x, 4, false); one(a1.y, c1.y, 5, false); one(a1.z, c1.z, 6, false); one(a1.w, c1.w, 7, false);
    if (two && C.cnt) {  // the second tile's multiplicities are requested only now: sixteen registers less at the peak
        d0 = ldc(ib);
        d1 = ldc(ib + 128);
        d2 = ldc(ib + 256);
        d3 = ldc(ib + 384);
    }
    one(a2.x, c2.x, 8, false); one(a2.y, c2.y, 9, false); one(a2.z, c2.z, 10, false); one(a2.w, c2.w, 11, false);
    one(a3.x, c3.x, 12, false); one(a3.y, c3.y, 13, false); one(a3.z, c3.z, 14, false); one(a3.w, c3.w, 15, false);
    if (two) {
        one(b0.x, d0.x, 16, true); one(b0.y, d0.y, 17, true); one(b0.z, d0.z, 18, true); one(b0.w, d0.w, 19, true);
        one(b1.x, d1.x, 20, true); one(b1.y, d1.y, 21, true); one(b1.z, d1.z, 22, true); one(b1.w, d1.w, 23, true);
        one(b2.x, d2.x, 24, true); one(b2.y, d2.y, 25, true); one(b2.z, d2.z, 26, true); one(b2.w, d2.w, 27, true);
        one(b3.x, d3.x, 28, true); one(b3.y, d3.y, 29, true); one(b3.z, d3.z, 30, true); one(b3.w, d3.w, 31, true);
    }
    if (pp) pp[3] = clock64() + (n1 & 0) + (int)(q1 & 0) + (int)(Sb & 0);
    n1 = __reduce_add_sync(0xffffffffu, n1);
    n2 = __reduce_add_sync(0xffffffffu, n2);
    o.pos1 = baseA + n1;
    o.sum1 = psA + warp_sum_ll(q1);
    o.cnt1 = C.cnt ? pcA + warp_sum_ll(w1) : o.pos1;
    o.pos2 = baseB + n2;
    o.sum2 = (two ? psB : psA) + warp_sum_ll(q2);
    o.cnt2 = C.cnt ? (two ? pcB : pcA) + warp_sum_ll(w2) : o.pos2;
    if (want_zone) {
        // slot s of this lane is the entry baseA + tile offset + ((s % 16) / 4) * 128 + 4 * lane + (s % 4), tile offset 0
        // for s < 16 (tile A) and boff for tile B; a lane's slots are in ascending position order, so its first / last
        // member is the lowest / highest set bit
        const unsigned boff = (unsigned)(baseB - baseA);
        auto at = [&](int sl) { return (unsigned)((sl >> 4) ? boff : 0u) + (unsigned)(((sl >> 2) & 3) * 128 + 4 * lane + (sl & 3)); };
        unsigned fa = ma ? at(__ffs(ma) - 1) : 0xffffffffu, fb = mb ? at(__ffs(mb) - 1) : 0xffffffffu;
        unsigned la = ma ? at(31 - __clz(ma)) + 1u : 0u, lb = mb ? at(31 - __clz(mb)) + 1u : 0u;
        // a zone wider than two tiles (centroids close together in a dense part of the data): the tiles between A and B
        // lie inside the zone entirely; one more round trip each
        for (long long tm = g1; tm < g2 - 1; ++tm) {
            const long long bm = tm * LL_TS;
            const float *pk = C.ks + bm + 4 * lane;
            const float4 m0 = ld_vol_f4(pk), m1 = ld_vol_f4(pk + 128), m2 = ld_vol_f4(pk + 256), m3 = ld_vol_f4(pk + 384);
            uint4 e0 = make_uint4(1u, 1u, 1u, 1u), e1 = e0, e2 = e0, e3 = e0;
            if (C.cnt) {
                const unsigned int *pc = C.cnt + bm + 4 * lane;
                e0 = ld_vol_u4(pc), e1 = ld_vol_u4(pc + 128), e2 = ld_vol_u4(pc + 256), e3 = ld_vol_u4(pc + 384);
            }
            unsigned int xa = 0, xb = 0;
            auto mid = [&](float x, unsigned int c, int slot) {
                const float xc = fsub(x, mean);
                const float m2x = fmul(-2.0f, xc);
                const float da = skl_dist(m2x, va, na), db = skl_dist(m2x, vb, nb);
                if (db < da || (db == da && b_wins_ties)) {
                    Wb += c;
                    Sb += fixed_qf(xc, C.scale_f, C.scale) * (long long)c;
                    xb |= 1u << slot;
                } else {
                    xa |= 1u << slot;
                }
            };
            mid(m0.x, e0.x, 0); mid(m0.y, e0.y, 1); mid(m0.z, e0.z, 2); mid(m0.w, e0.w, 3);
            mid(m1.x, e1.x, 4); mid(m1.y, e1.y, 5); mid(m1.z, e1.z, 6); mid(m1.w, e1.w, 7);
            mid(m2.x, e2.x, 8); mid(m2.y, e2.y, 9); mid(m2.z, e2.z, 10); mid(m2.w, e2.w, 11);
            mid(m3.x, e3.x, 12); mid(m3.y, e3.y, 13); mid(m3.z, e3.z, 14); mid(m3.w, e3.w, 15);
            const unsigned moff = (unsigned)(bm - baseA);
            auto atm = [&](int sl) { return moff + (unsigned)((sl >> 2) * 128 + 4 * lane + (sl & 3)); };
            if (xa) {
                fa = min(fa, atm(__ffs(xa) - 1));
                la = max(la, atm(31 - __clz(xa)) + 1u);
            }
            if (xb) {
                fb = min(fb, atm(__ffs(xb) - 1));
                lb = max(lb, atm(31 - __clz(xb)) + 1u);
            }
        }
        o.Wb = warp_sum_ll(Wb);
        o.Sb = warp_sum_ll(Sb);
        const unsigned z0 = (unsigned)n1;  // offset of the zone's first entry in tile A
        const unsigned rfa = __reduce_min_sync(0xffffffffu, fa), rfb = __reduce_min_sync(0xffffffffu, fb);
        const unsigned rla = __reduce_max_sync(0xffffffffu, la), rlb = __reduce_max_sync(0xffffffffu, lb);
        o.fa = rfa == 0xffffffffu ? 0u : rfa - z0 + 1u;
        o.fb = rfb == 0xffffffffu ? 0u : rfb - z0 + 1u;
        o.la = rla ? rla - z0 : 0u;  // (position + 1) - z0 = offset + 1
        o.lb = rlb ? rlb - z0 : 0u;
    }
    if (pp) pp[4] = clock64() + (int)(o.sum2 & 0);
    return want_zone;
}

// returns (uniform over the cluster):
// 0: the search pass evaluated every non-empty zone, nothing done (and no barrier needed); 1: chunks into CTA 0's slots;
// 2: some chunks took the generic path, CTA 0 has partials to pull
__device__ int fast_zone_step(FastSmem &S, FastSmem *S0, const FastConst &K, const SearchConst &C, const float *__restrict__ ks,
                              int gw, int NW, unsigned int etag, bool cta0, long long *prof = nullptr, int *dbg = nullptr) {
    const int NT = blockDim.x;
    if (prof) prof[0] = clock64();
    FastZone &Z = S.u.zn;
    const RegionTableT<LF_KMAX> &T = S.tab;
    const int R = T.R, tid = threadIdx.x, lane = lane_id();
    if (R <= 1) return 0;
    // Is any non-empty zone left?  CTA 0 looks (its own shared memory) and publishes the answer; the other CTAs poll ONE
    // word of it.  (With every CTA reading the positions and tags of all regions out of CTA 0, 23 K remote loads converged
    // on one SM in every iteration: 6 us per CTA, and CTA 0's own update step ran against that traffic.)
    if (cta0) {
        int left = 0;
        for (int r = tid; r < R; r += NT) {
            const int J2 = T.rJ2[r], J1 = T.rJ1[r];
            const bool open = J1 > J2 && S.rpos[r + 1] > S.rpos[r] && !(J1 == J2 + 1 && S.fz[J2].tag == etag);
            left |= open;
            if (dbg && open) atomicAdd(&dbg[J1 == J2 + 1 ? 0 : 1], 1);
        }
        left = __syncthreads_or(left);
        if (tid == 0) *(volatile unsigned int *)&S.zflag = (etag << 1) | (unsigned)(left != 0);
        if (!left) return 0;
    } else {
        if (tid == 0) {
            unsigned int v;
            long long spins = 0;
            do {
                v = *(volatile unsigned int *)&S0->zflag;
            } while ((v >> 1) != (etag & 0x7fffffffu) && ++spins < (1ll << 26));
            Z.s_warp[0] = (int)(v & 1u);
        }
        __syncthreads();
        const int left = Z.s_warp[0];
        __syncthreads();
        if (!left) return 0;
    }
    for (int r = tid; r < R; r += NT) {
        const int J2 = T.rJ2[r], J1 = T.rJ1[r];
        Z.rp[r] = S0->rpos[r];
        Z.hd[r] = J1 == J2 + 1 && S0->fz[J2].tag == etag;
    }
    if (tid == 0) Z.rp[R] = S0->rpos[R];
    __syncthreads();
    // chunks per zone -> exclusive prefix over the regions (every CTA computes the same numbers)
    int any_generic = 0;
    {
        const int per = (R + NT - 1) / NT;
        const int lo = min(R, tid * per), hi = min(R, lo + per);
        int sum = 0;
        for (int r = lo; r < hi; ++r) {
            const bool zone = T.rJ1[r] > T.rJ2[r] && !Z.hd[r];
            const long long sz = zone ? Z.rp[r + 1] - Z.rp[r] : 0ll;
            const long long nch = (sz + LF_CH - 1) / LF_CH;
            sum += (int)llmin2(nch, 1ll << 24);
            any_generic |= zone && sz > 0 && zone_is_generic(T, r);
        }
        const int incl = block_scan_incl<int>(sum, [](int a, int b) { return a + b; }, Z.s_warp);
        int run = incl - sum;
        for (int r = lo; r < hi; ++r) {
            S.cpre[r] = run;
            const bool zone = T.rJ1[r] > T.rJ2[r] && !Z.hd[r];
            const long long sz = zone ? Z.rp[r + 1] - Z.rp[r] : 0ll;
            run += (int)llmin2((sz + LF_CH - 1) / LF_CH, 1ll << 24);
        }
        if (tid == NT - 1) S.cpre[R] = incl;
    }
    any_generic = __syncthreads_or(any_generic);
    const int NC = S.cpre[R];
    if (NC > LF_MAXCH) any_generic = 1;
    if (prof) prof[1] = clock64();
    const unsigned int *__restrict__ ecnt = K.cnt;
#pragma unroll 1
    for (int ch = gw; ch < NC; ch += NW) {
        int lo_r = 0, hi_r = R;  // largest r with cpre[r] <= ch: the zone that holds the chunk
        while (hi_r - lo_r > 1) {
            const int mid = (lo_r + hi_r) >> 1;
            if (S.cpre[mid] <= ch)
                lo_r = mid;
            else
                hi_r = mid;
        }
        const int r = lo_r;
        const int J2 = T.rJ2[r], J1 = T.rJ1[r];
        const long long base = Z.rp[r] + (long long)(ch - S.cpre[r]) * LF_CH;
        const long long hi = llmin2(Z.rp[r + 1], base + LF_CH);
        if (J1 != J2 + 1 || ch >= LF_MAXCH) {
            fast_zone_generic(S, K, C, ks, J2, J1, base, hi);
            continue;
        }
        // two candidates a = J2 < b = J1: label b iff d_b < d_a, or d_b == d_a with the lower owner id
        const float va = T.dv[J2], vb = T.dv[J1], na = T.dcn[J2], nb = T.dcn[J1];
        const bool b_wins_ties = T.down[J1] < T.down[J2];
        long long Wb = 0, Sb = 0, Wt = 0, St = 0;
        float xv[4];
        unsigned int cv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // all loads of the chunk in flight
            const long long p = base + 32 * j + lane;
            const bool valid = p < hi;
            xv[j] = valid ? ld_vol_f1(ks + p) : 0.f;
            cv[j] = 1u;
            if (ecnt && valid) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(cv[j]) : "l"(ecnt + p));
        }
        int fa = 0, la = 0, fb = 0, lb = 0;  // chunk offsets + 1
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool valid = base + 32 * j + lane < hi;
            bool isb = false;
            if (valid) {
                const float xc = fsub(xv[j], K.mean);
                const float m2x = fmul(-2.0f, xc);
                const float da = skl_dist(m2x, va, na), db = skl_dist(m2x, vb, nb);
                isb = db < da || (db == da && b_wins_ties);
                const long long c = (long long)cv[j];
                const long long q = fixed_qf(xc, C.scale_f, K.scale) * c;
                Wt += c;
                St += q;
                if (isb) {
                    Wb += c;
                    Sb += q;
                }
            }
            const unsigned mb = __ballot_sync(0xffffffffu, valid && isb);
            const unsigned ma = __ballot_sync(0xffffffffu, valid && !isb);
            if (ma) {
                if (!fa) fa = 32 * j + __ffs(ma);
                la = 32 * j + 32 - __clz(ma);
            }
            if (mb) {
                if (!fb) fb = 32 * j + __ffs(mb);
                lb = 32 * j + 32 - __clz(mb);
            }
        }
        Wt = warp_sum_ll(Wt);
        St = warp_sum_ll(St);
        Wb = warp_sum_ll(Wb);
        Sb = warp_sum_ll(Sb);
        if (lane == 0) {
            ZoneSlot *sl = &S0->slot[ch];
            sl->Wb = Wb;
            sl->Sb = Sb;
            sl->Wt = Wt;
            sl->St = St;
            sl->ends = (unsigned)fa | ((unsigned)la << 8) | ((unsigned)fb << 16) | ((unsigned)lb << 24);
        }
    }
    if (prof) prof[2] = clock64();
    if (prof) prof[3] = clock64();
    return any_generic ? 2 : 1;
}

// CTA 0: per distinct index (count, sum, first / last member) of the ZONE entries -> U.Wd / Sd / first / last, from the chunk
// slots of the two-candidate zones and, when `pull`, from the generic partials of every CTA (DSMEM loads).
__device__ void fast_gather_zones(cg::cluster_group &cluster, FastSmem &S, int zs, unsigned int etag) {
    const bool pull = zs == 2, chunks = zs != 0;
    FastUpdate &U = S.u.up;
    const RegionTableT<LF_KMAX> &T = S.tab;
    const int NT = blockDim.x, n_cta = (int)cluster.num_blocks(), tid = threadIdx.x, m = T.m, R = T.R;
    for (int i = tid; i < m; i += NT) {
        U.aW[i] = U.aS[i] = U.bW[i] = U.bS[i] = 0;
        U.af[i] = U.bf[i] = 0x7fffffffffffffffll;
        U.al[i] = U.bl[i] = -1;
    }
    __syncthreads();
    for (int r = tid; r < R; r += NT) {
        const int J2 = T.rJ2[r], J1 = T.rJ1[r];
        if (J1 != J2 + 1) continue;
        const long long lo = S.rpos[r];
        if (S.fz[J2].tag == etag) {  // evaluated by the search pass
            const FusedZone &z = S.fz[J2];
            U.aW[J2] = (S.rcnt[r + 1] - S.rcnt[r]) - z.Wb;
            U.aS[J2] = (S.rsum[r + 1] - S.rsum[r]) - z.Sb;
            U.bW[J1] = z.Wb;
            U.bS[J1] = z.Sb;
            if (z.fa) {
                U.af[J2] = lo + z.fa - 1;
                U.al[J2] = lo + z.la - 1;
            }
            if (z.fb) {
                U.bf[J1] = lo + z.fb - 1;
                U.bl[J1] = lo + z.lb - 1;
            }
            continue;
        }
        if (!chunks) continue;
        const int c0 = S.cpre[r], c1 = min(S.cpre[r + 1], LF_MAXCH);
        if (c0 >= c1) continue;
        long long Wb = 0, Sb = 0, Wt = 0, St = 0, fa = 0x7fffffffffffffffll, la = -1, fb = 0x7fffffffffffffffll, lb = -1;
        for (int c = c0; c < c1; ++c) {
            const ZoneSlot &sl = S.slot[c];
            Wb += sl.Wb;
            Sb += sl.Sb;
            Wt += sl.Wt;
            St += sl.St;
            const long long cb = lo + (long long)(c - c0) * LF_CH - 1;  // chunk base - 1: the ends are offsets + 1
            const unsigned e = sl.ends;
            if (e & 0xffu) {
                fa = llmin2(fa, cb + (e & 0xffu));
                la = cb + ((e >> 8) & 0xffu);
            }
            if ((e >> 16) & 0xffu) {
                fb = llmin2(fb, cb + ((e >> 16) & 0xffu));
                lb = cb + (e >> 24);
            }
        }
        U.aW[J2] = Wt - Wb;
        U.aS[J2] = St - Sb;
        U.af[J2] = fa;
        U.al[J2] = la;
        U.bW[J1] = Wb;
        U.bS[J1] = Sb;
        U.bf[J1] = fb;
        U.bl[J1] = lb;
    }
    __syncthreads();
    if (pull) {
        // generic chunks: every CTA's local partials.  16 consecutive lanes take one distinct index, one CTA each (four
        // remote loads in flight per lane), and fold over the CTAs with shuffles.
        static_assert(LF_CL_MAX == 16, "the fold below is over 16 lanes");
        for (int base = 0; base < m * LF_CL_MAX; base += NT) {  // (NT is a multiple of 32: whole warps stay together)
            const int idx = base + tid, cta = idx & (LF_CL_MAX - 1), i = idx >> 4;
            long long w = 0, sm = 0, mn = 0x7fffffffffffffffll, mx = -1;
            if (i < m && cta < n_cta) {
                const volatile FastSmem *Sr = (const volatile FastSmem *)cluster.map_shared_rank(&S, cta);
                w = (long long)Sr->zW[i];
                sm = Sr->zS[i];
                mn = Sr->zmin[i];
                mx = Sr->zmax[i];
            }
#pragma unroll
            for (int o = 1; o < LF_CL_MAX; o <<= 1) {
                w += __shfl_xor_sync(0xffffffffu, w, o);
                sm += __shfl_xor_sync(0xffffffffu, sm, o);
                mn = llmin2(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = llmax2(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if (i < m && cta == 0) {  // (a distinct index is folded by exactly one group of lanes)
                U.aW[i] += w;
                U.aS[i] += sm;
                U.af[i] = llmin2(U.af[i], mn);
                U.al[i] = llmax2(U.al[i], mx);
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < m; i += NT) {
        const long long w = U.aW[i] + U.bW[i], sm = U.aS[i] + U.bS[i];
        const long long mn = llmin2(U.af[i], U.bf[i]), mx = llmax2(U.al[i], U.bl[i]);
        U.Wd[i] = w;
        U.Sd[i] = sm;
        U.first[i] = mn;
        U.last[i] = mx;
    }
}

static_assert(sizeof(ZoneSlot) * LF_MAXCH >= sizeof(uint4) * LF_POP_ENT && 2 * LF_KMAX * 2 <= LF_POP_ENT, "PopState entries alias the chunk slots");
__device__ __forceinline__ uint4 *pop_ent(FastSmem &S) { return reinterpret_cast<uint4 *>(S.slot); }

// relocation: (re)loads up to D members of candidate stream sidx, starting `off` entries from the stream's end (one thread)
__device__ __noinline__ void pop_fill(FastSmem &S, const FastConst &K, const float *__restrict__ ks, int sidx, int D, int off) {
    FastUpdate &U = S.u.up;
    PopState &P = U.pop;
    const RegionTableT<LF_KMAX> &T = S.tab;
    const unsigned int *__restrict__ ecnt = K.cnt;
    const int i = sidx >> 1;
    const bool right = sidx & 1;
    const long long first = U.first[i], last = U.last[i];
    const float c = T.dv[i], mean = K.mean;
    int n = 0, next = -1;
    P.hk[sidx] = make_uint4(0u, 0u, 0u, 0u);
    if (last >= first && last >= 0) {
        for (;;) {
            float x[LF_POP_D];
            unsigned int cn[LF_POP_D];
#pragma unroll
            for (int j = 0; j < LF_POP_D; ++j) {  // all loads in flight
                const long long p = right ? last - off - j : first + off + j;
                const bool ok = j < D && p >= first && p <= last;
                x[j] = ok ? ld_vol_f1(ks + p) : 0.f;
                cn[j] = 1u;
                if (ok && ecnt) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(cn[j]) : "l"(ecnt + p));
            }
            bool ended = false;
            int r = -1;  // region of the previous position (the walk moves one entry at a time: one search, then steps)
#pragma unroll
            for (int j = 0; j < LF_POP_D; ++j) {
                if (j >= D || ended) continue;
                const long long p = right ? last - off - j : first + off + j;
                if (p < first || p > last) {
                    ended = true;
                    continue;
                }
                const float xc = fsub(x[j], mean);
                if (right ? !(xc >= c) : !(xc < c)) {  // the other stream's half of the cluster
                    ended = true;
                    continue;
                }
                if (r < 0) {
                    int lo = 0, hi = T.R;  // largest r with rpos[r] <= p
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (S.rpos[mid] <= p)
                            lo = mid;
                        else
                            hi = mid;
                    }
                    r = lo;
                } else if (right) {
                    while (S.rpos[r] > p) --r;
                } else {
                    while (r + 1 < T.R && S.rpos[r + 1] <= p) ++r;
                }
                const int lab = T.rJ1[r] == T.rJ2[r] ? T.rJ1[r] : zone_argmin(xc, T.dv, T.dcn, T.down, T.rJ2[r], T.rJ1[r]);
                if (lab != i) continue;  // a neighbour's member inside a zone
                const FarKey key = far_key(xc, c);
                const uint4 e = make_uint4(key.d2, key.gap, key.ordx, cn[j]);
                if (n == 0)
                    P.hk[sidx] = e;
                else
                    pop_ent(S)[sidx * (D - 1) + n - 1] = e;
                ++n;
            }
            if (ended) break;
            off += D;
            next = off;
            if (n > 0) break;
            next = -1;  // nothing but neighbours so far: keep walking
        }
    }
    P.head[sidx] = 0;
    P.depth[sidx] = (unsigned char)n;
    P.next_off[sidx] = next;
}

// relocation: one WARP refills the window of candidate stream sidx from `off` entries behind the stream's end -- 32 entries
// in one coalesced round trip, every lane checks its entry, the first D members go into the window
__device__ __noinline__ void pop_refill_warp(FastSmem &S, const FastConst &K, const float *__restrict__ ks, int sidx, int D, int off) {
    FastUpdate &U = S.u.up;
    PopState &P = U.pop;
    const RegionTableT<LF_KMAX> &T = S.tab;
    const unsigned int *__restrict__ ecnt = K.cnt;
    const int lane = lane_id();
    const int i = sidx >> 1;
    const bool right = sidx & 1;
    const long long first = U.first[i], last = U.last[i];
    const float c = T.dv[i], mean = K.mean;
    int n = 0, next = -1;
    for (;;) {
        const long long p = right ? last - off - lane : first + off + lane;
        const bool inside = p >= first && p <= last;
        float x = 0.f;
        unsigned int cn = 1u;
        if (inside) {
            x = ld_vol_f1(ks + p);
            if (ecnt) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(cn) : "l"(ecnt + p));
        }
        const float xc = fsub(x, mean);
        const bool stop = !inside || (right ? !(xc >= c) : !(xc < c));  // the stream ends at the first such entry
        const unsigned stops = __ballot_sync(0xffffffffu, stop);
        const int e = stops ? __ffs(stops) - 1 : 32;
        bool member = false;
        if (lane < e) {
            int lo = 0, hi = T.R;  // largest r with rpos[r] <= p
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (S.rpos[mid] <= p)
                    lo = mid;
                else
                    hi = mid;
            }
            const int lab = T.rJ1[lo] == T.rJ2[lo] ? T.rJ1[lo] : zone_argmin(xc, T.dv, T.dcn, T.down, T.rJ2[lo], T.rJ1[lo]);
            member = lab == i;
        }
        const unsigned mem = __ballot_sync(0xffffffffu, member);
        const int rank = __popc(mem & ((1u << lane) - 1u));
        if (member && rank < D) {
            const FarKey key = far_key(xc, c);
            const uint4 ent = make_uint4(key.d2, key.gap, key.ordx, cn);
            if (rank == 0)
                P.hk[sidx] = ent;
            else
                pop_ent(S)[sidx * (D - 1) + rank - 1] = ent;
        }
        const int found = __popc(mem);
        n = min(found, D);
        if (found > D) {  // continue right behind the last member that went into the window
            const unsigned kept = __fns(mem, 0, D);  // lane of the D-th member
            next = off + (int)kept + 1;
            break;
        }
        if (e < 32) {
            next = -1;
            break;
        }
        off += 32;
        next = off;
        if (n > 0) break;
    }
    if (lane == 0) {
        if (n == 0) P.hk[sidx] = make_uint4(0u, 0u, 0u, 0u);
        P.head[sidx] = 0;
        P.depth[sidx] = (unsigned char)n;
        P.next_off[sidx] = next;
    }
    __syncwarp();
}

// ---- update step (CTA 0 only; S is its own shared memory): per-cluster counts / sums, label-equality proxy, empty-cluster
// relocation, averages, centre shift, convergence.  Mirrors update_phase of lloyd.cu statement by statement.
__device__ void fast_update_step(cg::cluster_group &cluster, FastSmem &S, const FastConst &K, const float *__restrict__ ks,
                                 const PeerComm &pc, unsigned long long &xseq, float tol, int zs, unsigned int etag,
                                 long long *prof = nullptr) {
    FastUpdate &U = S.u.up;
    const int NT = blockDim.x, n_cta = (int)cluster.num_blocks();
    if (prof) prof[0] = clock64();
    const RegionTableT<LF_KMAX> &T = S.tab;
    const int tid = threadIdx.x, k = K.k, m = T.m, R = T.R;
    const float mean = K.mean;
    const double scale = K.scale;
    const float x0 = fsub(0.f, mean);
    const int iter_now = S.iter;  // (read before any barrier: thread 0 advances it at the very end)
    // ---- 1. per distinct index: zone entries + SAFE regions.
    // The regular case first -- the table alternates SAFE(i), ZONE(i, i+1), SAFE(i+1), ... and the search pass evaluated
    // every non-empty zone (zs == 0): distinct index i is candidate b of region 2i - 1, owns region 2i and is candidate a of
    // region 2i + 1, so one thread per index adds its three pieces up directly; no scatter, no zeroing, one barrier.
    bool regular = zs == 0 && R == 2 * m - 1 && m <= NT;
    if (regular) {
        bool ok = true;
        if (tid < m) {
            const int i = tid;
            long long W = 0, Ssum = 0, fi = 0x7fffffffffffffffll, la = -1;
            ok = T.rJ1[2 * i] == i && T.rJ2[2 * i] == i;
            if (i > 0) {  // region 2i - 1: zone (i - 1 | i), this index is b
                const int r = 2 * i - 1;
                ok = ok && T.rJ2[r] == i - 1 && T.rJ1[r] == i;
                const long long lo = S.rpos[r];
                const FusedZone &z = S.fz[i - 1];
                if (S.rpos[r + 1] > lo && z.tag == etag) {
                    W += z.Wb;
                    Ssum += z.Sb;
                    if (z.fb) {
                        fi = llmin2(fi, lo + z.fb - 1);
                        la = llmax2(la, lo + z.lb - 1);
                    }
                }
            }
            {  // region 2i: SAFE
                const int r = 2 * i;
                const long long lo = S.rpos[r], hi = S.rpos[r + 1];
                if (hi > lo) {
                    W += S.rcnt[r + 1] - S.rcnt[r];
                    Ssum += S.rsum[r + 1] - S.rsum[r];
                    fi = llmin2(fi, lo);
                    la = llmax2(la, hi - 1);
                }
            }
            if (i < m - 1) {  // region 2i + 1: zone (i | i + 1), this index is a
                const int r = 2 * i + 1;
                const long long lo = S.rpos[r];
                const FusedZone &z = S.fz[i];
                if (S.rpos[r + 1] > lo && z.tag == etag) {
                    W += (S.rcnt[r + 1] - S.rcnt[r]) - z.Wb;
                    Ssum += (S.rsum[r + 1] - S.rsum[r]) - z.Sb;
                    if (z.fa) {
                        fi = llmin2(fi, lo + z.fa - 1);
                        la = llmax2(la, lo + z.la - 1);
                    }
                }
            }
            U.Wd[i] = W;
            U.Sd[i] = Ssum;
            U.first[i] = fi;
            U.last[i] = la;
        }
        if (tid < k) {
            U.W[tid] = 0;
            U.S[tid] = 0;
        }
        if (tid == 0) {
            U.same = 1;
            U.n_empty = 0;
        }
        regular = __syncthreads_and(ok) != 0;  // (an irregular table: the general path below redoes everything)
    }
    if (!regular) {
        fast_gather_zones(cluster, S, zs, etag);
        if (tid < k) {
            U.W[tid] = 0;
            U.S[tid] = 0;
        }
        if (tid == 0) {
            U.same = 1;
            U.n_empty = 0;
        }
        __syncthreads();
        for (int r = tid; r < R; r += NT) {  // SAFE regions: positions [rpos[r], rpos[r+1]) carry one label
            if (T.rJ1[r] != T.rJ2[r]) continue;
            const int di = T.rJ1[r];
            const long long lo = S.rpos[r], hi = S.rpos[r + 1];
            if (hi > lo) {
                U.Wd[di] += S.rcnt[r + 1] - S.rcnt[r];  // (J, J) occurs in at most one region: no conflicts
                U.Sd[di] += S.rsum[r + 1] - S.rsum[r];
                U.first[di] = llmin2(U.first[di], lo);
                U.last[di] = llmax2(U.last[di], hi - 1);
            }
        }
    }
    // ---- 2. zero run: the label of x'_0 = fl(0 - mean) is a region-table lookup -- the table is exact for every point of
    // the line, not only for samples (one thread; the block argmin over all distinct centroids cost three barriers)
    if (tid == NT - 1) {
        int zdi = -1;
        if (K.n0 > 0) {
            int lo = 0, hi = R;  // largest r with rstart[r] <= x0 (rstart[0] = -inf)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (T.rstart[mid] <= x0)
                    lo = mid;
                else
                    hi = mid;
            }
            zdi = T.rJ1[lo] == T.rJ2[lo] ? T.rJ1[lo] : zone_argmin(x0, T.dv, T.dcn, T.down, T.rJ2[lo], T.rJ1[lo]);
        }
        U.zdi = zdi;
    }
    __syncthreads();
    if (prof) prof[1] = clock64();
    if (K.n0 > 0 && tid == U.zdi) {
        const long long q0 = fixed_q(x0, scale);
        U.Wd[tid] += K.n0;
        U.Sd[tid] += K.n0 * q0;
    }
    __syncthreads();
    if (pc.enabled) {  // fused all-reduce of the exact per-distinct-index (count, sum) over the ranks
        if (tid < m) {
            S.xbuf[tid] = U.Wd[tid];
            S.xbuf[m + tid] = U.Sd[tid];
        }
        __syncthreads();
        if (!peer_allreduce_sum_tagged(pc, S.xbuf, 2 * m, ++xseq) && tid == 0) S.comm_error = 1;
        if (tid < m) {
            U.Wd[tid] = S.xbuf[tid];
            U.Sd[tid] = S.xbuf[m + tid];
        }
        __syncthreads();
    }
    if (prof) prof[2] = clock64();
    // ---- 3. per cluster id
    if (tid < m) {
        U.W[T.down[tid]] = U.Wd[tid];
        U.S[T.down[tid]] = U.Sd[tid];
    }
    __syncthreads();
    // ---- 4. label-equality proxy: identical exact (count, sum) per cluster as in the previous iteration
    int differs = 0;
    if (tid < k) {
        differs = (U.W[tid] != S.Wprev[tid]) || (U.S[tid] != S.Sprev[tid]);
        S.Wprev[tid] = U.W[tid];
        S.Sprev[tid] = U.S[tid];
    }
    const int any_differs = __syncthreads_or(differs);
    // ---- 5. empty clusters (ascending id) and relocation
    const int e = (tid < k) && (U.W[tid] == 0);
    const int n_empty = __syncthreads_count(e);  // (almost always 0: the scan below is skipped)
    if (tid == 0) U.n_empty = n_empty;
    if (n_empty > 0) {
        const int incl = block_scan_incl<int>(e, [](int a, int b) { return a + b; }, U.red_i);
        if (e) U.empt[incl - 1] = tid;
        __syncthreads();
    }
    if (prof) prof[3] = clock64();
    if (n_empty > 0) {
        // Candidate streams (PopState).  The farthest remaining sample overall is always at the head of one of the
        // streams (or is the zero run); pop until n_empty samples are taken.  The pops are a serial chain, so a round is
        // kept short: every thread owns the heads of its streams (one each for m <= NT / 2), a round is a warp arg-max
        // (four redux operations), one block barrier, and the same arg-max over the warps' winners; the owner of the
        // winning stream advances it out of the preloaded window, or its WARP refills the window in one coalesced round
        // trip.  Equal samples (an entry's multiplicity, the zero run) are taken in one round.
        PopState &P = U.pop;
        const int D = 2 * m * (LF_POP_D - 1) <= LF_POP_ENT ? LF_POP_D : 3;
        for (int sidx = tid; sidx < 2 * m; sidx += NT) pop_fill(S, K, ks, sidx, D, 0);
        __syncthreads();
        if (prof) prof[8] = clock64();
        int n_rounds = 0, n_refill = 0;
        long long *cand = S.xbuf;  // this rank's list, descending: (d2 << 32 | gap, ordx << 32 | old cluster id + 1)
        int pop = 0;
        {
            const int lane = lane_id(), wid = warp_id(), nw = NT >> 5;
            const int zdi = U.zdi, ZS = 2 * m;  // the zero run is one more stream
            const FarKey kz = far_key(x0, zdi >= 0 ? T.dv[zdi] : 0.f);
            long long zero_left = zdi >= 0 ? K.n0 : 0;  // kept by every thread (uniform)
            // [2][32] winners of the warps and their multiplicities, double buffered (one barrier per round); they live in
            // the part of the gather scratch that PopState does not cover
            static_assert(sizeof(PopState) <= 6 * LF_KMAX * sizeof(long long), "bf / bl of the gather scratch must stay free");
            uint4 *arg = reinterpret_cast<uint4 *>(U.bl);
            unsigned int *argw = reinterpret_cast<unsigned int *>(U.bf);
            for (;;) {
                // best head among this thread's streams (the empty head is the key (0, 0, 0), below every real key)
                uint32_t bd = 0, bg = 0, bo = 0, bs = 0xffffffffu, bw = 0;
                for (int i = tid; i < 2 * m; i += NT) {
                    const uint4 e = P.hk[i];
                    if (e.w && (bs == 0xffffffffu || e.x > bd || (e.x == bd && (e.y > bg || (e.y == bg && e.z > bo)))))
                        bd = e.x, bg = e.y, bo = e.z, bs = (uint32_t)i, bw = e.w;
                }
                const uint32_t mine = bs;
                auto warp_best = [&]() {  // arg-max over the lanes; equal keys: the lower stream
                    bool in = bs != 0xffffffffu;
                    const uint32_t M1 = __reduce_max_sync(0xffffffffu, in ? bd : 0u);
                    in = in && bd == M1;
                    const uint32_t M2 = __reduce_max_sync(0xffffffffu, in ? bg : 0u);
                    in = in && bg == M2;
                    const uint32_t M3 = __reduce_max_sync(0xffffffffu, in ? bo : 0u);
                    in = in && bo == M3;
                    bs = __reduce_min_sync(0xffffffffu, in ? bs : 0xffffffffu);
                    bd = M1, bg = M2, bo = M3;
                };
                warp_best();
                uint4 *slot = arg + ((n_rounds & 1) << 5);
                unsigned int *slotw = argw + ((n_rounds & 1) << 5);
                if (lane == 0) slot[wid] = make_uint4(bd, bg, bo, bs);
                if (mine != 0xffffffffu && mine == bs) slotw[wid] = bw;  // (the one lane that owns the warp's winner)
                __syncthreads();
                {
                    const uint4 e = lane < nw ? slot[lane] : make_uint4(0u, 0u, 0u, 0xffffffffu);
                    bd = e.x, bg = e.y, bo = e.z, bs = e.w;
                    bw = lane < nw ? slotw[lane] : 0u;
                }
                const uint32_t wmine = bs;
                warp_best();
                {
                    const unsigned who = __ballot_sync(0xffffffffu, wmine != 0xffffffffu && wmine == bs);
                    bw = __shfl_sync(0xffffffffu, bw, who ? __ffs(who) - 1 : 0);
                }
                if (zero_left > 0 && (bs == 0xffffffffu || kz.d2 > bd || (kz.d2 == bd && (kz.gap > bg || (kz.gap == bg && kz.ordx > bo)))))
                    bd = kz.d2, bg = kz.gap, bo = kz.ordx, bs = (uint32_t)ZS;
                if (bs == 0xffffffffu) break;  // this rank has no sample left
                const int ws = (int)bs;
                // samples taken from this entry: all of them, or what is still needed
                const long long avail = ws == ZS ? zero_left : (long long)bw;
                const int take = (int)llmin2(avail, (long long)(n_empty - pop));
                if (wid == 0) {
                    const int di = ws == ZS ? zdi : (ws >> 1);
                    const unsigned long long ca = ((unsigned long long)bd << 32) | bg;
                    const unsigned long long cb = ((unsigned long long)bo << 32) | (unsigned)(T.down[di] + 1);
                    for (int q = lane; q < take; q += 32) {
                        cand[2 * (pop + q)] = (long long)ca;
                        cand[2 * (pop + q) + 1] = (long long)cb;
                    }
                }
                pop += take;
                ++n_rounds;
                if (ws == ZS) {
                    zero_left -= take;
                } else if (wid == ((ws % NT) >> 5)) {  // the warp of the stream's owner
                    int refill = 0;
                    if (lane == (ws & 31)) {
                        const unsigned left = P.hk[ws].w - (unsigned)take;
                        if (left) {
                            P.hk[ws].w = left;
                        } else {
                            const int h = P.head[ws] + 1;
                            if (h < P.depth[ws]) {
                                P.hk[ws] = pop_ent(S)[ws * (D - 1) + h - 1];
                                P.head[ws] = (unsigned char)h;
                            } else if (P.next_off[ws] >= 0) {
                                refill = 1;
                            } else {
                                P.hk[ws].w = 0;
                            }
                        }
                    }
                    if (__any_sync(0xffffffffu, refill)) {  // drained deeper than the window: one coalesced round trip
                        pop_refill_warp(S, K, ks, ws, D, P.next_off[ws]);
                        ++n_refill;
                    }
                }
                if (pop >= n_empty) break;
            }
            if (prof) {
                prof[9] = clock64();
                prof[10] = n_rounds | ((long long)n_refill << 16);
            }
        }
        __syncthreads();
        if (tid == 0) U.winner = pop;
        __syncthreads();
        const int n_done = U.winner;
        if (!pc.enabled) {
            // one rank: its list is the global list.  np.max(distances) == 0 -> relocation is skipped altogether
            // (_k_means_common.pyx:192-195)
            const bool skip = n_done > 0 && ((unsigned long long)cand[0] >> 32) == 0ull;
            if (!skip) {
                for (int i = tid; i < n_done; i += NT) {
                    const unsigned long long bb = (unsigned long long)cand[2 * i + 1];
                    U.far_x[i] = ord2f((uint32_t)(bb >> 32));
                    U.far_old[i] = (int)(bb & 0xffffffffull) - 1;
                }
            }
            if (tid == 0) U.winner = skip ? 0 : n_done;  // samples to move
            __syncthreads();
        } else {
            // unused slots of this rank hold zeros; fused all-gather of the candidate lists (n_empty is the same on every rank)
            for (int i = 2 * n_done + tid; i < 2 * n_empty; i += NT) cand[i] = 0;
            __syncthreads();
            if (!peer_allgather_tagged(pc, reinterpret_cast<const unsigned long long *>(cand), 2 * n_empty, K.cand, (size_t)k * 2, ++xseq) &&
                tid == 0)
                S.comm_error = 1;
            __threadfence_block();
            __syncthreads();
            // ---- 5b. the n_empty farthest samples over all ranks (every rank's list is in descending order: a W-way merge by
            // one thread)
            if (tid == 0) {
                const int world = K.world;
                int *cur = S.merge_cur;
                for (int r = 0; r < world; ++r) cur[r] = 0;
                int n_moved = 0;
                for (int i = 0; i < n_empty; ++i) {
                    int br = -1;
                    unsigned long long ba = 0, bb = 0;
                    for (int r = 0; r < world; ++r) {
                        if (cur[r] >= n_empty) continue;
                        const unsigned long long a = K.cand[2 * ((size_t)r * k + cur[r])], b = K.cand[2 * ((size_t)r * k + cur[r]) + 1];
                        if ((b & 0xffffffffull) == 0) continue;  // list exhausted
                        if (br < 0 || a > ba || (a == ba && (b >> 32) > (bb >> 32))) {
                            br = r;
                            ba = a;
                            bb = b;
                        }
                    }
                    if (br < 0) break;
                    // np.max(distances) == 0 -> relocation is skipped altogether (_k_means_common.pyx:192-195)
                    if (i == 0 && (ba >> 32) == 0ull) break;
                    cur[br]++;
                    U.far_x[i] = ord2f((uint32_t)(bb >> 32));
                    U.far_old[i] = (int)(bb & 0xffffffffull) - 1;
                    n_moved = i + 1;
                }
                U.winner = n_moved;
            }
            __syncthreads();
        }
        // the far samples move to the empty clusters in ascending id order
        if (tid == 0) {
            const int n_moved = U.winner;
            for (int i = 0; i < n_moved; ++i) {
                const int nwid = U.empt[i], od = U.far_old[i];
                const long long q = fixed_q(U.far_x[i], scale);
                U.S[od] -= q;
                U.S[nwid] = q;
                U.W[nwid] = 1;
                U.W[od] -= 1;
            }
            S.n_reloc += n_moved;
        }
        __syncthreads();
    }
    if (prof) prof[4] = clock64();
    // ---- 6. averages (_average_centers), shift (_center_shift)
    // S / 2^s == S * 2^-s exactly (a power-of-two scaling of the double image of S), without a float64 division
    if (tid < k) U.raw[tid] = (float)__dmul_rn((double)U.S[tid], K.inv_scale);
    // argmax of the counts, lowest id on ties (np.argmax): key = count << 10 | (1023 - id), count < 2^53
    int amax_all = 0;
    {
        unsigned long long key = tid < k ? (((unsigned long long)U.W[tid] << 10) | (unsigned long long)(1023 - tid)) : 0ull;
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if (lane_id() == 0) U.red_w[warp_id()] = key;
        __syncthreads();
        const int nwk = (k + 31) >> 5;  // every thread folds the warps' winners itself: no second barrier
        key = U.red_w[0];
        for (int w = 1; w < nwk; ++w) key = U.red_w[w] > key ? U.red_w[w] : key;
        amax_all = 1023 - (int)(key & 1023ull);
    }
    if (tid < k) {
        const int amax = amax_all;
        // (float)(1.0 / (double)W): for W < 2^24 (exact in float32) the float32 division gives the same value -- double
        // rounding is innocuous for a quotient of float32 operands when the wide format has >= 2 * 24 + 2 bits
        auto avg = [&](int j) {
            const long long Wj = U.W[j];
            const float rcp = Wj < (1ll << 24) ? __fdiv_rn(1.0f, (float)Wj) : (float)__ddiv_rn(1.0, (double)Wj);
            return fmul(U.raw[j], rcp);
        };
        float cn;
        if (U.W[tid] > 0)
            cn = avg(tid);
        else  // in-place ascending loop: rows below amax see its raw sum, rows above see its average
            cn = amax < tid ? (U.W[amax] > 0 ? avg(amax) : U.raw[amax]) : U.raw[amax];
        U.cnew[tid] = cn;
        const float t = fsub(cn, S.c[tid]);
        const float r = fmul(t, t);
        // (float)sqrt((double)r) == the correctly rounded float32 sqrt: double rounding is innocuous for sqrt when the wide
        // format has >= 2 * 24 + 2 bits (Figueroa)
        const float sh = __fsqrt_rn(r);
        U.sq[tid] = fmul(sh, sh);
    }
    __syncthreads();
    if (prof) prof[5] = clock64();
    // ---- 7. convergence (_kmeans.py:721-738)
    if (warp_id() == 0) {
        const float tot = np_pairwise_warp(U.sq, k, S.np, S.np_leaves);
        int stop = 0, strict = 0;
        if (!any_differs) {
            stop = 1;
            strict = 1;
        } else if (tot <= tol) {
            stop = 1;
        }
        if (S.comm_error) stop = 1;  // an exchange timed out: the call fails; do not wait through the remaining iterations
        if (tid == 0) {
            U.strict = strict;
            U.stop = stop;
        }
    }
    __syncthreads();
    if (prof) prof[6] = clock64();
    {
        const int strict = U.strict, stop = U.stop;
        const int last_iter = iter_now + 1 >= K.max_iter;
        if (tid < k) {
            // labels of a strict stop belong to the centroids the E-step used; otherwise a final E-step with the new
            // centroids follows (emit.cu).  (A thread touches only its own centroid: no barrier between the two stores.)
            if (stop || last_iter) S.c_emit[tid] = strict ? S.c[tid] : U.cnew[tid];
            S.c[tid] = U.cnew[tid];
        }
        if (tid == 0) {
            S.iter = iter_now + 1;
            if (stop || last_iter) {
                S.done = 1;
                S.strict = strict;
                S.n_iter = iter_now + 1;
            }
        }
    }
    // (no barrier here: the caller's cluster barrier follows)
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
    ll_fast_kernel(LloydDevice *st, const float *__restrict__ ks, const float *__restrict__ samp, const long long *__restrict__ ptile,
                   const float *__restrict__ init, int want_hist, int want_log, PeerComm pc) {
    extern __shared__ __align__(16) unsigned char fast_smem_raw[];
    FastSmem &S = *reinterpret_cast<FastSmem *>(fast_smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned cta = cluster.block_rank();
    const int n_cta = (int)cluster.num_blocks();
    FastSmem *S0 = cluster.map_shared_rank(&S, 0);
    constexpr int NT = THREADS;
    const int tid = threadIdx.x;
    const int NW = n_cta * (NT / 32);                  // warps of the cluster
    const int gw = (int)cta * (NT / 32) + warp_id();  // warp index inside the cluster

    // kernel-lifetime constants: built by thread 0 in shared memory, read from there (see FastSmem)
    if (tid == 0) {
        FastConst Kc;
        Kc.k = st->k;
        Kc.rank = st->rank;
        Kc.world = st->world;
        Kc.max_iter = st->max_iter;
        Kc.n = st->n;
        Kc.n_nz = st->n_nz;
        Kc.n0 = st->n0;
        Kc.n_ent = st->n_ent;
        Kc.cnt = st->cnt;
        Kc.cand = st->cand;
        Kc.mean = st->mean;
        Kc.xabs_max = st->xabs_max;
        Kc.scale = st->scale;
        Kc.inv_scale = __ddiv_rn(1.0, Kc.scale);
        S.K = Kc;
        SearchConst Cc = search_const(st, ks, samp, ptile);
        if (Cc.n_tiles > 64) {  // top level of the tile-sample index in shared memory (the samples never change)
            Cc.top_step = (Cc.n_tiles + LL_TOP - 1) / LL_TOP;
            Cc.top_n = (int)((Cc.n_tiles + Cc.top_step - 1) / Cc.top_step);
            Cc.top = S.top;
        }
        S.C = Cc;
    }
    for (int i = tid; i < 32 * 10; i += NT) (&S.hint[0][0])[i] = -1;
    for (int i = tid; i < LF_KMAX; i += NT) S.fz[i].tag = 0u;  // E-steps are numbered from 1
    if (tid == 0) S.zflag = 0u;
    if (want_log && cta == 0 && tid < 24) st->logG[LL_LOG - 24 + tid] = 0;
    __syncthreads();
    const FastConst &K = S.K;
    const SearchConst &C = S.C;
    const int k = K.k;
    if (C.top)
        for (int i = tid; i < C.top_n; i += NT) S.top[i] = samp[(long long)i * C.top_step];
    const long long total_q = C.n_tiles > 0 ? ptile[C.n_tiles] : 0ll;
    unsigned long long xseq = (pc.enabled && cta == 0) ? *peer_counter(pc) : 0ull;
    // ---- init: centre the initial centroids; tolerance from the exact integer moments of all n samples
    if (tid < k) {
        S.c[tid] = fsub(init[tid], K.mean);  // init -= X_mean  (_kmeans.py:1493)
        S.perm[tid] = 0;
    }
    if (tid == 0) {
        S.np_leaves = np_leaf_list(0, k, S.np, 0);
        {
            int leaf = 0, n_ops = 0;
            S.np.root = np_fold_prog(k, S.np, leaf, n_ops);
            S.np.n_ops = n_ops;
        }
        S.done = 0;
        S.iter = 0;
        S.strict = 0;
        S.n_reloc = 0;
        S.n_iter = 0;
        S.comm_error = 0;
    }
    float tol = 0.f;
    if (cta == 0) {
        if (tid < k) {
            S.Wprev[tid] = -1;
            S.Sprev[tid] = 0;
        }
        if (tid == 0) {  // this rank's moments: the sorted survivors (ll_tilesum_kernel) + its zero run
            const long long q0 = fixed_q(fsub(0.f, K.mean), K.scale);
            const unsigned long long qq = (unsigned long long)(q0 * q0);
            S.xbuf[0] = st->s1 + K.n0 * q0;
            S.xbuf[1] = (long long)(st->s2_lo + (unsigned long long)K.n0 * (qq & 0x7fffffffull));
            S.xbuf[2] = (long long)(st->s2_hi + (unsigned long long)K.n0 * (qq >> 31));
        }
        __syncthreads();
        if (pc.enabled) {
            if (!peer_allreduce_sum_tagged(pc, S.xbuf, 3, ++xseq) && tid == 0) S.comm_error = 1;
        }
        if (tid == 0) {
            const __int128 s1 = (__int128)S.xbuf[0];
            const unsigned __int128 s2 = ((unsigned __int128)(unsigned long long)S.xbuf[2] << 31) + (unsigned long long)S.xbuf[1];
            const unsigned __int128 num = (unsigned __int128)K.n * s2 - (unsigned __int128)(s1 * s1);
            const double numd = __dadd_rn(__dmul_rn((double)(unsigned long long)(num >> 64), 18446744073709551616.0),
                                          (double)(unsigned long long)num);
            const double nd = (double)K.n;
            const double var_d = __ddiv_rn(__ddiv_rn(numd, __dmul_rn(nd, nd)), __dmul_rn(K.scale, K.scale));
            float t = fmul((float)var_d, (float)st->tol_rel);
            if (st->tol_rel == 0.0) t = 0.f;
            st->tol = t;
            S.u.up.red_d[0] = t;
        }
        __syncthreads();
        tol = S.u.up.red_d[0];
        __syncthreads();
    }
    auto now = []() {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    };
    const bool logger = want_log && cta == 0 && tid == 0;

    long long *hint = S.hint[warp_id()];  // tile found last time for each boundary slot of this warp (<= 1022 boundaries over >= 128 warps)

    cluster.sync();  // every CTA's shared state is initialised before anybody writes into CTA 0's
    // One loop body for the Lloyd iterations AND the closing histogram round (one more E-step against c_emit, counts only):
    // a second copy of the E-step would double the kernel's code, and this kernel is bound by instruction fetch as much as
    // by memory latency (every phase is straight-line code executed once per iteration).
    bool hist_round = false;
    unsigned int estep = 0;
    for (int it = 0;; ++it) {
        unsigned long long tl[5];
        const bool lg = logger && !hist_round;
        tl[4] = lg ? now() : 0ull;
        // ---- E-step against the centroids in S.c: table, boundary search, zones
        build_region_table(S.c, k, K.xabs_max, &S.tab, S.u.tb, S.perm);
        const int R = S.tab.R;
        if (tid < S.tab.m) {
            S.zW[tid] = 0;
            S.zS[tid] = 0;
            S.zmin[tid] = 0x7fffffffffffffffll;
            S.zmax[tid] = -1;
        }
        if (cta == 0 && tid == 0) {
            S.rpos[0] = 0;
            S.rcnt[0] = 0;
            S.rsum[0] = 0;
            S.rpos[R] = K.n_ent;
            S.rcnt[R] = K.n_nz;
            S.rsum[R] = total_q;
        }
        if (lg) tl[0] = now();
        const bool stamp = want_log > 1 && tid == 0 && it == 6 && !hist_round;
        unsigned long long *sl = reinterpret_cast<unsigned long long *>(st->logZ) + 8 * cta;  // [cta][8] ns stamps
        if (stamp) sl[0] = now();
        const unsigned int etag = ++estep;  // this E-step's number (per-thread copy, the same everywhere)
        {
            int trip = 0;
            const RegionTableT<LF_KMAX> &T = S.tab;
#pragma unroll 1
            // one warp per PAIR of region boundaries; consecutive pairs go to different CTAs (the wide zones -- centroids
            // close together where the data are dense -- are neighbours, and a CTA full of them finished last)
            const int gwi = warp_id() * n_cta + (int)cta;
            for (int b1 = 1 + 2 * gwi; b1 < R; b1 += 2 * NW, ++trip) {
                if (b1 + 1 >= R) {  // a last single boundary
                    long long r3[3];
                    boundary_search_call(C, T.rstart[b1], r3, &hint[trip < 10 ? trip : 9]);
                    if (lane_id() == 0) {
                        S0->rpos[b1] = r3[0];
                        S0->rcnt[b1] = r3[1];
                        S0->rsum[b1] = r3[2];
                    }
                    continue;
                }
                const int J2 = T.rJ2[b1], J1 = T.rJ1[b1];  // region b1 lies between the two boundaries
                const bool zone2 = J1 == J2 + 1;
                PairOut o;
                long long pp[6];
                const bool prof_pair = lg && tid == 0 && it == 6;
                const bool done = warp_pair_search(C, T.rstart[b1], T.rstart[b1 + 1], zone2, T.dv[J2], T.dcn[J2], T.dv[J1],
                                                   T.dcn[J1], T.down[J1] < T.down[J2], o, want_log ? pp : nullptr);
                if (prof_pair)
                    for (int i = 0; i < 5; ++i) st->logZ[LL_LOG - 80 + i] = pp[i] - pp[0];
                if (want_log && lane_id() == 0 && zone2 && !done) atomicAdd(&st->logG[LL_LOG - 12 + (int)(pp[5] & 0xff)], 1);
                if (want_log && lane_id() == 0 && zone2 && it == 6) atomicAdd(&st->logG[LL_LOG - 24 + (int)min(7ll, pp[5] >> 8)], 1);
                if (want_log && lane_id() == 0) atomicAdd(&st->logG[LL_LOG - 1 - (done ? 0 : (zone2 ? 1 : 2))], 1);
                if (lane_id() == 0) {
                    S0->rpos[b1] = o.pos1;
                    S0->rcnt[b1] = o.cnt1;
                    S0->rsum[b1] = o.sum1;
                    S0->rpos[b1 + 1] = o.pos2;
                    S0->rcnt[b1 + 1] = o.cnt2;
                    S0->rsum[b1 + 1] = o.sum2;
                    if (zone2) {
                        FusedZone *z = &S0->fz[J2];
                        if (done) {
                            z->Wb = o.Wb;
                            z->Sb = o.Sb;
                            z->fa = (unsigned short)o.fa;
                            z->la = (unsigned short)o.la;
                            z->fb = (unsigned short)o.fb;
                            z->lb = (unsigned short)o.lb;
                        }
                        z->tag = done ? etag : 0u;
                    }
                }
            }
        }
        if (stamp) sl[1] = now();
        cluster.sync();
        if (stamp) sl[2] = now();
        if (lg) tl[1] = now();
        int zs = 0;
        {
            long long zp[4] = {0, 0, 0, 0};
            zs = fast_zone_step(S, S0, K, C, ks, gw, NW, etag, cta == 0, lg ? zp : nullptr, want_log ? &st->logG[LL_LOG - 16] : nullptr);
            if (lg && S.iter == 6)
                for (int i = 0; i < 4; ++i) st->logZ[LL_LOG - 16 + i] = zp[i] - zp[0];
            if (stamp) sl[3] = now();
            if (want_log && cta == 0 && tid == 0) atomicAdd(&st->logG[LL_LOG - 4 - zs], 1);
            if (zs) cluster.sync();  // (uniform) nothing was left to the chunk pass: CTA 0 already has everything
        }
        if (stamp) sl[4] = now();
        if (lg) tl[2] = now();
        // ---- M-step (CTA 0), or the counts of the closing round
        if (cta == 0) {
            if (!hist_round) {
                long long up[12];
                up[8] = up[9] = up[10] = 0;
                const int it_now = S.iter;
                fast_update_step(cluster, S, K, ks, pc, xseq, tol, zs, etag, lg ? up : nullptr);
                if (lg && (it_now == 6 || it_now < 2)) {  // update profile of iterations 6, 0 and 1
                    up[7] = clock64();
                    const int at = it_now == 6 ? 32 : it_now == 0 ? 48 : 64;
                    for (int i = 0; i < 10; ++i) st->logZ[LL_LOG - at + i] = up[i] ? up[i] - up[0] : 0;
                    st->logZ[LL_LOG - at + 10] = up[10];
                }
                if (lg && it_now < LL_LOG - 64 && want_log == 1) {  // (the per-CTA stamps of NNC_LLOYD_LOG=2 share logZ)
                    long long z = 0;
                    for (int r = 0; r < R; ++r)
                        if (S.tab.rJ1[r] != S.tab.rJ2[r]) z += S.rpos[r + 1] - S.rpos[r];
                    st->logZ[it_now] = z;
                    st->logG[it_now] = R;
                    st->logM[it_now] = S.tab.m;
                    st->logE[it_now] = S.u.up.n_empty;
                }
            } else {
                long long *Wd = S.u.up.Wd;
                const RegionTableT<LF_KMAX> &T = S.tab;
                const int m = T.m;
                if (tid < k) st->hist[tid] = 0;
                fast_gather_zones(cluster, S, zs, etag);
                __syncthreads();
                for (int r = tid; r < R; r += NT) {
                    if (T.rJ1[r] != T.rJ2[r]) continue;
                    const long long c = S.rcnt[r + 1] - S.rcnt[r];
                    if (c > 0) Wd[T.rJ1[r]] += c;  // (J, J) occurs in at most one region
                }
                __syncthreads();
                if (tid == 0 && K.n0 > 0) Wd[zone_argmin(fsub(0.f, K.mean), T.dv, T.dcn, T.down, 0, m - 1)] += K.n0;
                __syncthreads();
                if (pc.enabled) {  // code histogram of all ranks: exact counts summed over the mailboxes
                    if (tid < k) S.xbuf[tid] = 0;
                    __syncthreads();
                    if (tid < m) S.xbuf[T.down[tid]] = Wd[tid];
                    __syncthreads();
                    if (!peer_allreduce_sum_tagged(pc, S.xbuf, k, ++xseq) && tid == 0) st->pad2 = 1;
                    if (tid < k) st->hist[tid] = S.xbuf[tid];
                } else if (tid < m) {
                    st->hist[T.down[tid]] = Wd[tid];
                }
            }
        }
        if (stamp) sl[5] = now();
        cluster.sync();
        if (stamp) sl[6] = now();
        if (hist_round) break;
        const int stopped = S0->done;
        if (lg && it < LL_LOG) {
            st->logT[it][3] = (unsigned int)(tl[0] - tl[4]);  // table
            st->logT[it][0] = (unsigned int)(tl[1] - tl[0]);  // search (+ barrier)
            st->logT[it][1] = (unsigned int)(tl[2] - tl[1]);  // zone (+ barrier)
            st->logT[it][2] = (unsigned int)(now() - tl[2]);  // update (+ barrier)
        }
        if (!stopped) {  // pull the new centroids from CTA 0
            if (cta != 0 && tid < k) S.c[tid] = S0->c[tid];
            __syncthreads();
            continue;
        }
        // ---- the loop has stopped: results out (CTA 0 holds them)
        if (cta == 0) {
            if (tid < k) {
                st->c[tid] = S.c[tid];
                st->c_emit[tid] = S.c_emit[tid];
            }
            if (tid == 0) {
                st->iter = S.iter;
                st->done = S.done;
                st->strict = S.strict;
                st->n_reloc = S.n_reloc;
                st->n_iter = S.n_iter;
                st->comm_error = S.comm_error;
                st->pad2 = S.comm_error;  // travels with the loop's outcome (one read-back)
            }
        }
        if (!want_hist || S0->comm_error) break;  // (after a time-out the histogram round would only wait again)
        hist_round = true;  // code histogram of the final labelling: the E-step once more, against c_emit
        if (tid < k) S.c[tid] = S0->c_emit[tid];
        __syncthreads();
    }
    if (pc.enabled && cta == 0 && tid == 0) *peer_counter(pc) = xseq;
    cluster.sync();  // no CTA exits while another may still access its shared memory
}

bool lloyd_fast_applicable(int k) { return k >= 1 && k <= LF_KMAX && !getenv("NNC_LLOYD_NO_CLUSTER"); }

// Launches the cluster loop; `st` has its header, moments (s1 / s2 of the sorted survivors) and candidate buffer set.
// Preferred shape: ONE cluster of 16 CTAs x 512 threads (a non-portable cluster size: every GPC of a B200 has >= 16 SMs;
// 128 registers per thread), else 8 x 1024 (64 registers).  NNC_LLOYD_CLUSTER=8|16 forces one.
template <int THREADS>
static cudaError_t fast_launch_shape(nnc_ctx *ctx, int n_cta, bool probe_only, LloydDevice *st, const float *d_sorted, const float *samp,
                                     const long long *ptile, const float *d_init, int want_hist, int want_log, const PeerComm &pc) {
    func_dyn_smem(ctx, (const void *)ll_fast_kernel<THREADS>, sizeof(FastSmem));
    if (n_cta > 8) {
        cudaError_t e = cudaFuncSetAttribute(ll_fast_kernel<THREADS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_cta);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = sizeof(FastSmem);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = n_cta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (probe_only) {
        int n_clusters = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n_clusters, ll_fast_kernel<THREADS>, &cfg);
        if (e != cudaSuccess) return e;
        return n_clusters >= 1 ? cudaSuccess : cudaErrorInvalidClusterSize;
    }
    return cudaLaunchKernelEx(&cfg, ll_fast_kernel<THREADS>, st, d_sorted, samp, ptile, d_init, want_hist, want_log, pc);
}

void lloyd_fast_launch(nnc_ctx *ctx, LloydDevice *st, const float *d_sorted, const float *samp, const long long *ptile,
                       const float *d_init, int want_hist, const PeerComm &pc, int k) {
    const int want_log = getenv("NNC_LLOYD_LOG") ? atoi(getenv("NNC_LLOYD_LOG")) : 0;
    if (ctx->fast_cluster == 0) {  // decide once per context (= per device)
        int want = 16;
        if (const char *e = getenv("NNC_LLOYD_CLUSTER")) want = atoi(e) == 8 ? 8 : 16;
        ctx->fast_cluster = 8;
        if (want == 16 && fast_launch_shape<512>(ctx, 16, true, st, d_sorted, samp, ptile, d_init, want_hist, want_log, pc) == cudaSuccess)
            ctx->fast_cluster = 16;
        cudaGetLastError();
    }
    // The cluster is as large as the E-step needs: one warp per pair of region boundaries, at most k - 1 pairs.  A small
    // codebook (2- to 5-bit quantisation) runs in 1 - 2 CTAs: cheaper barriers, and the clusters of a model's tensors run
    // side by side (16-CTA clusters of several streams were observed to run one after the other).
    // (four pairs per CTA -- one per warp scheduler: a pair search is a long dependent chain, co-resident searches on one
    // scheduler take turns)
    int n_cta = 1;
    while (n_cta < ctx->fast_cluster && n_cta * 4 < k - 1) n_cta <<= 1;
    if (getenv("NNC_LLOYD_FULL_CLUSTER")) n_cta = ctx->fast_cluster;
    cudaError_t e = ctx->fast_cluster == 16
                        ? fast_launch_shape<512>(ctx, n_cta, false, st, d_sorted, samp, ptile, d_init, want_hist, want_log, pc)
                        : fast_launch_shape<1024>(ctx, n_cta, false, st, d_sorted, samp, ptile, d_init, want_hist, want_log, pc);
    ctx->launches++;
    if (e != cudaSuccess) NNC_FAIL(NNC_ERR_CUDA, "ll_fast_kernel launch (cluster of %d): %s", n_cta, cudaGetErrorString(e));
}

}  // namespace nnc
