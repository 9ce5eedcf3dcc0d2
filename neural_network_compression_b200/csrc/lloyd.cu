// lloyd.cu -- Lloyd iterations of 1-D k-means on the SORTED surviving weights.
//
// Restates sklearn's _kmeans_single_lloyd / lloyd_iter_chunked_dense / _relocate_empty_clusters_dense /
// _average_centers / _center_shift (sklearn/cluster/_kmeans.py:630-758, _k_means_lloyd.pyx:23-218,
// _k_means_common.pyx:167-311), which the reference reaches through KMeans(...).fit in
// neural_network_compression/common/utility.py:237-238.
//
// Data layout (all in HBM):
//   ks[n_nz]        sorted non-zero weights (float32, raw values)
//   ptile[T+1]      int64 exclusive prefix sums of q(x') at tile granularity (tile = 1024 sorted elements),
//                   q(x') = rint(fl(x - mean) * 2^(30-E)) the fixed-point image of the centred sample
//   samp[T]         first key of every tile (32-ary search index, L2 resident)
//   the n0 pruned zeros are not stored: they are one value with multiplicity n0.
//
// One iteration = four small kernels on the context stream, no host round trip:
//   table   (1 CTA)   sort the k centroids, build the region table (table.cuh): <= 2m-1 regions, each either
//                     SAFE (one label) or a ZONE (float32 rule over a short candidate range)
//   search  (grid)    one warp per region boundary: position in ks[] and prefix sum of q up to it
//   zone    (grid)    evaluate the float32 label rule for the few samples inside the zones
//   update  (1 CTA)   per-cluster counts and sums (SAFE regions via prefix differences + zone partials +
//                     the zero run), empty-cluster relocation, averages, centre shift, convergence test
// Per-cluster sums are exact integers, so the result does not depend on reduction order or sharding.
#include <stddef.h>
#include <stdlib.h>

#include <algorithm>

#include "lloyd_shared.cuh"

namespace nnc {

// ---------------------------------------------------------------------------------------------
// prep: per-tile sums of q, samples, moments
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ll_tilesum_kernel(const float *__restrict__ ks, const unsigned int *__restrict__ cnt,
                                                         long long n_ent, float mean, double scale, float scale_f, long long *tsum,
                                                         long long *tcnt, float *samp, LloydDevice *st) {
    const long long n_tiles = (n_ent + LL_TS - 1) / LL_TS;
    const int wpb = blockDim.x >> 5, lane = lane_id();
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(ks) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(cnt) & 15u) == 0);
    long long s1 = 0;
    unsigned __int128 s2 = 0;  // sum of count * q^2 (q^2 < 2^61, count < 2^32)
    for (long long t = blockIdx.x * (long long)wpb + warp_id(); t < n_tiles; t += (long long)gridDim.x * wpb) {
        const long long base = t * LL_TS;
        long long acc = 0, cacc = 0;
        auto one = [&](float x, unsigned int cu) {
            const long long c = (long long)cu;
            const long long q = fixed_qf(fsub(x, mean), scale_f, scale);
            acc += q * c;
            cacc += c;
            s2 += (unsigned __int128)(unsigned long long)(q * q) * (unsigned long long)c;
        };
        if (vec_ok && base + LL_TS <= n_ent) {  // a full tile: four 128-bit loads of values (and counts) in flight per lane
            const float *pk = ks + base + 4 * lane;
            const float4 a0 = ld_vol_f4(pk), a1 = ld_vol_f4(pk + 128), a2 = ld_vol_f4(pk + 256), a3 = ld_vol_f4(pk + 384);
            uint4 c0 = make_uint4(1u, 1u, 1u, 1u), c1 = c0, c2 = c0, c3 = c0;
            if (cnt) {
                const unsigned int *pc = cnt + base + 4 * lane;
                c0 = ld_vol_u4(pc);
                c1 = ld_vol_u4(pc + 128);
                c2 = ld_vol_u4(pc + 256);
                c3 = ld_vol_u4(pc + 384);
            }
            if (lane == 0) samp[t] = a0.x;
            one(a0.x, c0.x); one(a0.y, c0.y); one(a0.z, c0.z); one(a0.w, c0.w);
            one(a1.x, c1.x); one(a1.y, c1.y); one(a1.z, c1.z); one(a1.w, c1.w);
            one(a2.x, c2.x); one(a2.y, c2.y); one(a2.z, c2.z); one(a2.w, c2.w);
            one(a3.x, c3.x); one(a3.y, c3.y); one(a3.z, c3.z); one(a3.w, c3.w);
        } else {
            for (int j = 0; j < LL_TS / 32; ++j) {
                const long long i = base + j * 32 + lane;
                if (i < n_ent) {
                    const float x = ks[i];
                    one(x, cnt ? cnt[i] : 1u);
                    if (j == 0 && lane == 0) samp[t] = x;
                }
            }
        }
        acc = warp_sum_ll(acc);
        if (lane == 0) tsum[t] = acc;
        if (tcnt) {
            cacc = warp_sum_ll(cacc);
            if (lane == 0) tcnt[t] = cacc;
        }
        s1 += acc;  // every lane holds the warp total; only lane 0 contributes below
    }
    unsigned long long s2lo = warp_sum_ull((unsigned long long)(s2 & 0x7fffffffull));
    unsigned long long s2hi = warp_sum_ull((unsigned long long)(s2 >> 31));
    if (lane == 0) {
        atomicAdd((unsigned long long *)&st->s1, (unsigned long long)s1);
        atomicAdd(&st->s2_lo, s2lo);
        atomicAdd(&st->s2_hi, s2hi);
    }
}

// total_q = ptile[n_tiles] (the exclusive scan of the tile sums is exclusive_scan_i64, scan.cu)
__global__ void ll_total_kernel(const long long *ptile, long long n_tiles, LloydDevice *st) { st->total_q = ptile[n_tiles]; }

// ---------------------------------------------------------------------------------------------
// init: centre the initial centroids, tolerance from exact integer moments
// ---------------------------------------------------------------------------------------------
// the zero run of this rank joins its exact moments (same 31-bit split as ll_tilesum_kernel)
__global__ void ll_moments_kernel(LloydDevice *st) {
    const long long q0 = fixed_q(fsub(0.f, st->mean), st->scale);
    const unsigned long long qq = (unsigned long long)(q0 * q0);
    st->s1 += st->n0 * q0;
    st->s2_lo += (unsigned long long)st->n0 * (qq & 0x7fffffffull);
    st->s2_hi += (unsigned long long)st->n0 * (qq >> 31);
}

__global__ void ll_init_kernel(LloydDevice *st, const float *init) {
    const int tid = threadIdx.x;
    if (tid < st->k) {
        st->c[tid] = fsub(init[tid], st->mean);  // init -= X_mean  (_kmeans.py:1493)
        st->Wprev[tid] = -1;
        st->Sprev[tid] = 0;
    }
    if (tid == 0) {
        // s1 / s2 cover all n samples of all ranks (ll_moments_kernel added the zero runs, then the all-reduce)
        __int128 s1 = (__int128)st->s1;
        unsigned __int128 s2 = ((unsigned __int128)st->s2_hi << 31) + st->s2_lo;
        unsigned __int128 num = (unsigned __int128)st->n * s2 - (unsigned __int128)(s1 * s1);
        double numd = __dadd_rn(__dmul_rn((double)(unsigned long long)(num >> 64), 18446744073709551616.0),
                                (double)(unsigned long long)num);
        double nd = (double)st->n;
        double var_d = __ddiv_rn(__ddiv_rn(numd, __dmul_rn(nd, nd)), __dmul_rn(st->scale, st->scale));
        float tol = fmul((float)var_d, (float)st->tol_rel);
        if (st->tol_rel == 0.0) tol = 0.f;
        st->tol = tol;
        st->iter = 0;
        st->done = 0;
        st->strict = 0;
        st->n_reloc = 0;
        st->n_iter = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// per-iteration kernels
// ---------------------------------------------------------------------------------------------
__device__ void table_phase(LloydDevice *st, TableScratch &S) {
    build_region_table(st->c, st->k, st->xabs_max, &st->tab, S, st->perm);
    const int tid = threadIdx.x;
    if (tid < st->k) {
        st->zW[tid] = 0;
        st->zS[tid] = 0;
        st->zmin[tid] = 0x7fffffffffffffffll;
        st->zmax[tid] = -1;
    }
    if (tid == 0) {
        const int R = st->tab.R;
        st->rpos[0] = 0;
        st->rcnt[0] = 0;
        st->rsum[0] = 0;
        st->rpos[R] = st->n_ent;
        st->rcnt[R] = st->n_nz;
        st->rsum[R] = st->total_q;
    }
}
__global__ void __launch_bounds__(TB_THREADS) ll_table_kernel(LloydDevice *st) {
    __shared__ TableScratch S;
    if (st->done) return;
    table_phase(st, S);
}

__device__ __forceinline__ void search_phase(LloydDevice *st, const SearchConst &C) {
    const int nb = st->tab.R - 1;  // boundaries 1 .. R-1
    const int wpb = blockDim.x >> 5;
    // warps of different CTAs take consecutive boundaries: the work spreads over all SMs
    for (int r = 1 + warp_id() * gridDim.x + blockIdx.x; r <= nb; r += gridDim.x * wpb) {
        long long pos, cn, sum;
        warp_boundary_search(C, st->tab.rstart[r], pos, cn, sum);
        if (lane_id() == 0) {
            st->rpos[r] = pos;
            st->rcnt[r] = cn;
            st->rsum[r] = sum;
        }
    }
}
__global__ void __launch_bounds__(256) ll_search_kernel(LloydDevice *st, const float *__restrict__ ks,
                                                        const float *__restrict__ samp,
                                                        const long long *__restrict__ ptile) {
    if (st->done) return;
    search_phase(st, search_const(st, ks, samp, ptile));
}

struct ZoneSmem {
    long long rp[2 * TB_KMAX + 2];  // copy of the region positions (the per-element lookup stays on chip)
    long long zpre[2 * TB_KMAX + 2];
    long long s_warp[32];
    unsigned long long sW[TB_KMAX];
    long long sS[TB_KMAX];
    long long sMin[TB_KMAX];
    long long sMax[TB_KMAX];
};

__device__ void zone_phase(LloydDevice *st, const float *__restrict__ ks, ZoneSmem &Z_) {
    long long *zpre = Z_.zpre, *s_warp = Z_.s_warp, *sS = Z_.sS, *sMin = Z_.sMin, *sMax = Z_.sMax;
    unsigned long long *sW = Z_.sW;
    const RegionTable &T = st->tab;
    const int R = T.R, m = T.m;
    if (R <= 1) return;
    // prefix of zone sizes over the regions (SAFE regions count 0): chunked scan by the threads of the CTA
    {
        long long *rp = Z_.rp;
        for (int r = threadIdx.x; r <= R; r += blockDim.x) rp[r] = st->rpos[r];
        __syncthreads();
        auto zsize = [&](int r) -> long long { return T.rJ1[r] > T.rJ2[r] ? rp[r + 1] - rp[r] : 0ll; };
        const int per = (R + blockDim.x - 1) / blockDim.x;
        const int lo = min(R, (int)threadIdx.x * per), hi = min(R, lo + per);
        long long sum = 0;
        for (int r = lo; r < hi; ++r) sum += zsize(r);
        long long incl = block_scan_incl<long long>(sum, [](long long a, long long b) { return a + b; }, s_warp);
        long long run = incl - sum;
        for (int r = lo; r < hi; ++r) {
            zpre[r] = run;
            run += zsize(r);
        }
        if (threadIdx.x == blockDim.x - 1) zpre[R] = incl;
    }
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        sW[i] = 0;
        sS[i] = 0;
        sMin[i] = 0x7fffffffffffffffll;
        sMax[i] = -1;
    }
    __syncthreads();
    const long long Z = zpre[R];
    if (blockIdx.x == 0 && threadIdx.x == 0 && st->iter < LL_LOG) {
        st->logZ[st->iter] = Z;
        st->logG[st->iter] = R;
        st->logM[st->iter] = m;
    }
    if (Z == 0) return;
    long long chunk = (Z + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + blockDim.x - 1) / blockDim.x * blockDim.x;
    const long long e0 = (long long)blockIdx.x * chunk, e1 = llmin2(Z, e0 + chunk);
    const float mean = st->mean;
    const double scale = st->scale;
    const int lane = lane_id();
    const unsigned int *__restrict__ ecnt = st->cnt;
    for (long long eb = e0; eb < e1; eb += blockDim.x) {  // e1 - e0 is a multiple of blockDim.x or ends at Z: uniform trips
        long long e = eb + threadIdx.x;
        bool valid = e < e1;
        int di = -1;
        long long q = 0, p = 0, c = 0;
        if (valid) {
            int lo = 0, hi = R;  // largest r with zpre[r] <= e: the non-empty zone that holds e
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (zpre[mid] <= e)
                    lo = mid;
                else
                    hi = mid;
            }
            const int r = lo;
            p = Z_.rp[r] + (e - zpre[r]);
            float xc = fsub(ks[p], mean);
            di = zone_argmin(xc, T.dv, T.dcn, T.down, T.rJ2[r], T.rJ1[r]);
            c = ecnt ? (long long)ecnt[p] : 1ll;
            q = fixed_q(xc, scale) * c;
        }
        unsigned active = __ballot_sync(0xffffffffu, valid);
        while (active) {
            int leader = __ffs(active) - 1;
            int L = __shfl_sync(0xffffffffu, di, leader);
            bool mine = valid && di == L;
            unsigned grp = __ballot_sync(0xffffffffu, mine);
            long long sq = warp_sum_ll(mine ? q : 0);
            long long sc = ecnt ? warp_sum_ll(mine ? c : 0) : (long long)__popc(grp);
            long long pf = __shfl_sync(0xffffffffu, p, __ffs(grp) - 1);
            long long pl = __shfl_sync(0xffffffffu, p, 31 - __clz(grp));
            if (lane == leader) {
                atomicAdd(&sW[L], (unsigned long long)sc);
                atomicAdd((unsigned long long *)&sS[L], (unsigned long long)sq);
                atomicMin(&sMin[L], pf);
                atomicMax(&sMax[L], pl);
            }
            active &= ~grp;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        if (sW[i]) {
            atomicAdd((unsigned long long *)&st->zW[i], sW[i]);
            atomicAdd((unsigned long long *)&st->zS[i], (unsigned long long)sS[i]);
            atomicMin(&st->zmin[i], sMin[i]);
            atomicMax(&st->zmax[i], sMax[i]);
        }
    }
}
__global__ void __launch_bounds__(256) ll_zone_kernel(LloydDevice *st, const float *__restrict__ ks) {
    if (st->done) return;
    extern __shared__ __align__(16) unsigned char zone_smem_raw[];
    zone_phase(st, ks, *reinterpret_cast<ZoneSmem *>(zone_smem_raw));
}

// ---- update ------------------------------------------------------------------------------------
// label (distinct index) of the sorted survivor at position p, from the region table + searched positions
__device__ int label_at(const LloydDevice *st, const long long *rpos, const float *ks, long long p) {
    const RegionTable &T = st->tab;
    const int R = T.R;
    int lo = 0, hi = R;  // largest r with rpos[r] <= p
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (rpos[mid] <= p)
            lo = mid;
        else
            hi = mid;
    }
    if (T.rJ1[lo] == T.rJ2[lo]) return T.rJ1[lo];
    return zone_argmin(fsub(ks[p], st->mean), T.dv, T.dcn, T.down, T.rJ2[lo], T.rJ1[lo]);
}

struct UpdateSmem {
    long long Wd[TB_KMAX], Sd[TB_KMAX];    // per distinct index, then reused per id
    long long W[TB_KMAX], S[TB_KMAX];      // per id
    long long first[TB_KMAX], last[TB_KMAX];  // member cursors per distinct index
    float raw[TB_KMAX], cnew[TB_KMAX], sq[TB_KMAX];
    int empt[TB_KMAX];
    float far_x[TB_KMAX];
    int far_old[TB_KMAX];
    float red_d[32];
    int red_i[32], red_id[32];
    uint32_t rk_d2[32], rk_gap[32], rk_ord[32];
    int rk_who[32];
    int n_empty, zdi, same, winner, stop_reloc;  // stop_reloc: scratch of the convergence step
    long long zero_left;
    long long rpos[2 * TB_KMAX + 2];  // copy of the region positions for the relocation cursors (label_at)
};

// Phases (run in one launch on a single GPU, as three launches with an all-reduce in between otherwise):
//   0  local per-distinct-index counts / sums / member cursors (+ the zero run)            -> gW, gS  [all-reduce]
//   1  per cluster id, label-equality proxy, empty clusters, LOCAL farthest candidates       -> cand    [all-gather]
//   2  merge the candidates, relocate, average, centre shift, convergence
// With a peer mailbox (pc.enabled, multi-GPU) the three phases run in ONE launch and the two exchanges happen inside
// the kernel over NVLink peer memory (peer.cuh).
__device__ void update_phase(LloydDevice *st, const float *__restrict__ ks, int phase_lo, int phase_hi, const PeerComm &pc,
                             UpdateSmem &U) {
    const RegionTable &T = st->tab;
    const int tid = threadIdx.x, k = st->k, m = T.m, R = T.R;
    const float mean = st->mean;
    const double scale = st->scale;
    const float x0 = fsub(0.f, mean);
    unsigned long long xseq = pc.enabled ? *peer_counter(pc) : 0ull;  // exchanges executed so far (same on every rank)
    if (phase_lo > 0) {  // resume: reload what the previous launch left (gW / gS now hold the global sums)
        if (tid < m) {
            U.Wd[tid] = st->gW[tid];
            U.Sd[tid] = st->gS[tid];
            U.first[tid] = st->gfirst[tid];
            U.last[tid] = st->glast[tid];
        }
        if (tid < k) {
            U.W[tid] = st->idW[tid];
            U.S[tid] = st->idS[tid];
            U.empt[tid] = st->empt_s[tid];
        }
        if (tid == 0) {
            U.zdi = st->zdi_s;
            U.n_empty = st->n_empty_s;
            U.same = st->same_s;
        }
        __syncthreads();
    }
    if (phase_lo == 0) {
    // ---- 1. per distinct index: zone partials + SAFE regions
    if (tid < m) {
        U.Wd[tid] = st->zW[tid];
        U.Sd[tid] = st->zS[tid];
        U.first[tid] = st->zmin[tid];
        U.last[tid] = st->zmax[tid];
    }
    if (tid < k) {
        U.W[tid] = 0;
        U.S[tid] = 0;
    }
    __syncthreads();
    for (int r = tid; r < R; r += TB_THREADS) {  // SAFE regions: positions [rpos[r], rpos[r+1]) carry one label
        if (T.rJ1[r] != T.rJ2[r]) continue;
        const int di = T.rJ1[r];
        long long lo = st->rpos[r], hi = st->rpos[r + 1];
        if (hi > lo) {
            U.Wd[di] += st->rcnt[r + 1] - st->rcnt[r];  // (J, J) occurs in at most one region: no conflicts
            U.Sd[di] += st->rsum[r + 1] - st->rsum[r];
            U.first[di] = llmin2(U.first[di], lo);
            U.last[di] = llmax2(U.last[di], hi - 1);
        }
    }
    __syncthreads();
    // ---- 2. zero run: label of x'_0 = fl(0 - mean) over all distinct centroids
    if (st->n0 > 0) {
        float d = INFINITY;
        int id = 0x7fffffff, di = -1;
        if (tid < m) {
            d = skl_dist(fmul(-2.0f, x0), T.dv[tid], T.dcn[tid]);
            id = T.down[tid];
            di = tid;
        }
        // block argmin by (d, id)
        for (int o = 16; o > 0; o >>= 1) {
            float d2 = __shfl_xor_sync(0xffffffffu, d, o);
            int id2 = __shfl_xor_sync(0xffffffffu, id, o);
            int di2 = __shfl_xor_sync(0xffffffffu, di, o);
            if (d2 < d || (d2 == d && id2 < id)) {
                d = d2;
                id = id2;
                di = di2;
            }
        }
        if (lane_id() == 0) {
            U.red_d[warp_id()] = d;
            U.red_id[warp_id()] = id;
            U.red_i[warp_id()] = di;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < TB_THREADS / 32; ++w) {
                if (U.red_d[w] < d || (U.red_d[w] == d && U.red_id[w] < id)) {
                    d = U.red_d[w];
                    id = U.red_id[w];
                    di = U.red_i[w];
                }
            }
            U.zdi = di;
            const long long q0 = fixed_q(x0, scale);
            U.Wd[di] += st->n0;
            U.Sd[di] += st->n0 * q0;
        }
        __syncthreads();
    } else if (tid == 0) {
        U.zdi = -1;
    }
    __syncthreads();
    if (phase_hi == 0) {
        if (tid < TB_KMAX) {
            st->gW[tid] = tid < m ? U.Wd[tid] : 0;
            st->gS[tid] = tid < m ? U.Sd[tid] : 0;
        }
        if (tid < m) {
            st->gfirst[tid] = U.first[tid];
            st->glast[tid] = U.last[tid];
        }
        if (tid == 0) st->zdi_s = U.zdi;
        return;
    }
    if (pc.enabled) {  // fused all-reduce of the exact per-distinct-index (count, sum) over the ranks
        if (tid < m) {
            st->xbuf[tid] = U.Wd[tid];
            st->xbuf[m + tid] = U.Sd[tid];
        }
        __syncthreads();
        if (!peer_allreduce_sum(pc, st->xbuf, 2 * m, ++xseq) && tid == 0) st->comm_error = 1;
        if (tid < m) {
            U.Wd[tid] = st->xbuf[tid];
            U.Sd[tid] = st->xbuf[m + tid];
        }
        __syncthreads();
    }
    }  // phase 0
    if (phase_lo <= 1) {
    // ---- 3. per cluster id
    if (tid < m) {
        U.W[T.down[tid]] = U.Wd[tid];
        U.S[T.down[tid]] = U.Sd[tid];
    }
    if (tid == 0) {
        U.same = 1;
        U.n_empty = 0;
    }
    __syncthreads();
    // ---- 4. label-equality proxy: identical exact (count, sum) per cluster as in the previous iteration
    if (tid < k) {
        if (U.W[tid] != st->Wprev[tid] || U.S[tid] != st->Sprev[tid]) U.same = 0;
        st->Wprev[tid] = U.W[tid];
        st->Sprev[tid] = U.S[tid];
    }
    __syncthreads();
    // ---- 5. empty clusters (ascending id) and relocation
    {
        int e = (tid < k) && (U.W[tid] == 0);
        int incl = block_scan_incl<int>(e, [](int a, int b) { return a + b; }, U.red_i);
        if (e) U.empt[incl - 1] = tid;
        if (tid == TB_THREADS - 1) U.n_empty = incl;
        __syncthreads();
    }
    const int n_empty = U.n_empty;
    if (tid == 0 && st->iter < LL_LOG) st->logE[st->iter] = n_empty;
    if (n_empty > 0) {
        // Streams of candidates: for every distinct index its members walked from the left end and from the
        // right end (|x' - c| is V-shaped along a cluster's sorted members), plus the zero run.  The farthest
        // remaining sample overall is always at the head of one of the streams; pop n_empty times.
        // Thread di owns both cursors of distinct index di; the owner of the zero run's cluster also owns
        // the zero run (a third head).
        long long pl = -1, pr = -2;  // empty stream when pl > pr
        // an entry stands for cnt identical samples: the cursors pop them one by one (reml / remr: samples left in
        // the entry under the left / right cursor; when the cursors meet, the left one owns what is left)
        unsigned long long reml = 0, remr = 0;
        const unsigned int *__restrict__ ecnt = st->cnt;
        auto cnt_at = [&](long long p) -> unsigned long long { return ecnt ? (unsigned long long)ecnt[p] : 1ull; };
        float cown = 0.f;
        FarKey kl{0, 0, 0}, kr{0, 0, 0};
        float xl = 0.f, xr = 0.f;
        bool has_l = false, has_r = false;
        if (tid < m) {
            cown = T.dv[tid];
            if (U.last[tid] >= U.first[tid] && U.last[tid] >= 0) {
                pl = U.first[tid];
                pr = U.last[tid];
                reml = cnt_at(pl);
                remr = cnt_at(pr);
            }
        }
        auto refresh = [&]() {
            has_l = has_r = false;
            if (tid < m && pl <= pr) {
                xl = fsub(ks[pl], mean);
                kl = far_key(xl, cown);
                has_l = true;
                if (pr > pl) {
                    xr = fsub(ks[pr], mean);
                    kr = far_key(xr, cown);
                    has_r = true;
                }
            }
        };
        for (int r = tid; r <= R; r += TB_THREADS) U.rpos[r] = st->rpos[r];
        __syncthreads();
        refresh();
        if (tid == 0) U.zero_left = st->n0;
        __syncthreads();
        const FarKey kz = far_key(x0, U.zdi >= 0 ? T.dv[U.zdi] : 0.f);
        unsigned long long *my_cand = st->cand + (size_t)st->rank * k * 2;
        int n_done = 0;
        for (int pop = 0; pop < n_empty; ++pop) {
            // best head of this thread: 0 = left, 1 = right, 2 = zero run
            FarKey best{0, 0, 0};
            int who = -1;
            if (has_l) {
                best = kl;
                who = 0;
            }
            if (has_r && (who < 0 || far_before(kr, best))) {
                best = kr;
                who = 1;
            }
            if (tid == U.zdi && U.zero_left > 0 && (who < 0 || far_before(kz, best))) {
                best = kz;
                who = 2;
            }
            int owner = who >= 0 ? tid : -1;
            // block argmax
            for (int o = 16; o > 0; o >>= 1) {
                FarKey ob;
                ob.d2 = __shfl_xor_sync(0xffffffffu, best.d2, o);
                ob.gap = __shfl_xor_sync(0xffffffffu, best.gap, o);
                ob.ordx = __shfl_xor_sync(0xffffffffu, best.ordx, o);
                int oo = __shfl_xor_sync(0xffffffffu, owner, o);
                if (oo >= 0 && (owner < 0 || far_before(ob, best) ||
                                (!far_before(best, ob) && oo < owner))) {
                    best = ob;
                    owner = oo;
                }
            }
            if (lane_id() == 0) {
                U.rk_d2[warp_id()] = best.d2;
                U.rk_gap[warp_id()] = best.gap;
                U.rk_ord[warp_id()] = best.ordx;
                U.rk_who[warp_id()] = owner;
            }
            __syncthreads();
            if (tid == 0) {  // fold of the warp winners: only the first ceil(m / 32) warps own streams
                const int nw = (m + 31) >> 5;
                for (int w = 1; w < nw; ++w) {
                    FarKey ob{U.rk_d2[w], U.rk_gap[w], U.rk_ord[w]};
                    int oo = U.rk_who[w];
                    if (oo >= 0 && (owner < 0 || far_before(ob, best) || (!far_before(best, ob) && oo < owner))) {
                        best = ob;
                        owner = oo;
                    }
                }
                U.winner = owner;
            }
            __syncthreads();
            if (U.winner < 0) break;  // this rank has no sample left
            if (tid == U.winner) {
                // candidate = (dist^2, ulp gap | x', old cluster id + 1): the local list comes out in descending order
                const FarKey kk = who == 0 ? kl : (who == 1 ? kr : kz);
                const int old_id = who == 2 ? T.down[U.zdi] : T.down[tid];
                my_cand[2 * pop] = ((unsigned long long)kk.d2 << 32) | kk.gap;
                my_cand[2 * pop + 1] = ((unsigned long long)kk.ordx << 32) | (unsigned)(old_id + 1);
                if (who == 2) {
                    U.zero_left -= 1;
                } else {
                    // take one sample of the entry; when it is used up, advance the cursor to the next member of this
                    // distinct index
                    if (who == 0) {
                        if (reml > 1) {
                            reml -= 1;
                        } else {
                            do {
                                ++pl;
                            } while (pl <= pr && label_at(st, U.rpos, ks, pl) != tid);
                            if (pl <= pr) reml = pl == pr ? remr : cnt_at(pl);
                            refresh();
                        }
                    } else {
                        if (remr > 1) {
                            remr -= 1;
                        } else {
                            do {
                                --pr;
                            } while (pr >= pl && label_at(st, U.rpos, ks, pr) != tid);
                            if (pr > pl) remr = cnt_at(pr);
                            refresh();
                        }
                    }
                }
            }
            n_done = pop + 1;
            __syncthreads();
        }
        __syncthreads();
        // unused slots of this rank, and (before the all-gather) every slot of the other ranks, hold zeros
        for (int i = tid; i < st->world * k; i += TB_THREADS) {
            const int r = i / k, j = i - r * k;
            if (r != st->rank || j >= n_done) {
                st->cand[2 * (size_t)i] = 0;
                st->cand[2 * (size_t)i + 1] = 0;
            }
        }
        __syncthreads();
    }
    if (phase_hi == 1) {
        if (tid < k) {
            st->idW[tid] = U.W[tid];
            st->idS[tid] = U.S[tid];
            st->empt_s[tid] = U.empt[tid];
        }
        if (tid == 0) {
            st->n_empty_s = U.n_empty;
            st->same_s = U.same;
        }
        return;
    }
    if (pc.enabled && U.n_empty > 0) {  // fused all-gather of the candidate lists (n_empty is the same on every rank)
        const int cnt = 2 * U.n_empty;
        unsigned long long *mine = st->cand + (size_t)st->rank * k * 2;
        for (int i = tid; i < cnt; i += TB_THREADS) st->xbuf[i] = (long long)mine[i];
        __syncthreads();
        if (!peer_allgather(pc, reinterpret_cast<const unsigned long long *>(st->xbuf), cnt, st->cand, (size_t)k * 2, ++xseq) &&
            tid == 0)
            st->comm_error = 1;
    }
    }  // phase 1
    if (pc.enabled && tid == 0) *peer_counter(pc) = xseq;
    // ---- 5b. relocation: the n_empty farthest samples over all ranks (every rank's list is already in descending
    // order: a W-way merge by one thread), moved to the empty clusters in ascending id order
    if (U.n_empty > 0) {
        __threadfence_block();
        if (tid == 0) {
            const int n_empty = U.n_empty, world = st->world;
            int cur[64];
            for (int r = 0; r < world; ++r) cur[r] = 0;
            int n_done = 0;
            bool skip = false;
            for (int i = 0; i < n_empty; ++i) {
                int br = -1;
                unsigned long long ba = 0, bb = 0;
                for (int r = 0; r < world; ++r) {
                    if (cur[r] >= k) continue;
                    const unsigned long long a = st->cand[2 * ((size_t)r * k + cur[r])], b = st->cand[2 * ((size_t)r * k + cur[r]) + 1];
                    if ((b & 0xffffffffull) == 0) continue;  // list exhausted
                    if (br < 0 || a > ba || (a == ba && (b >> 32) > (bb >> 32))) {
                        br = r;
                        ba = a;
                        bb = b;
                    }
                }
                if (br < 0) break;
                // np.max(distances) == 0 -> relocation is skipped altogether (_k_means_common.pyx:192-195)
                if (i == 0 && (ba >> 32) == 0ull) {
                    skip = true;
                    break;
                }
                cur[br]++;
                U.far_x[i] = ord2f((uint32_t)(bb >> 32));
                U.far_old[i] = (int)(bb & 0xffffffffull) - 1;
                n_done = i + 1;
            }
            if (!skip) {
                for (int i = 0; i < n_done; ++i) {
                    const int nw = U.empt[i], od = U.far_old[i];
                    const long long q = fixed_q(U.far_x[i], scale);
                    U.S[od] -= q;
                    U.S[nw] = q;
                    U.W[nw] = 1;
                    U.W[od] -= 1;
                }
                st->n_reloc += n_done;
            }
        }
        __syncthreads();
    }
    // ---- 6. averages (_average_centers), shift (_center_shift)
    if (tid < k) U.raw[tid] = (float)__ddiv_rn((double)U.S[tid], scale);
    __syncthreads();
    // argmax of the counts, lowest id on ties (np.argmax).  One thread, on purpose: a block reduction of (count, id)
    // pairs with long long shuffles here made the per-phase kernels (ll_update_kernel) fault with "misaligned address"
    // on sm_100a while the identical code inside ll_loop_kernel ran -- not understood, so the plain loop stays.
    if (tid == 0) {
        int amax = 0;
        for (int j = 1; j < k; ++j)
            if (U.W[j] > U.W[amax]) amax = j;
        U.winner = amax;
    }
    __syncthreads();
    if (tid < k) {
        const int amax = U.winner;
        auto avg = [&](int j) { return fmul(U.raw[j], (float)__ddiv_rn(1.0, (double)U.W[j])); };
        float cn;
        if (U.W[tid] > 0)
            cn = avg(tid);
        else  // in-place ascending loop: rows below amax see its raw sum, rows above see its average
            cn = amax < tid ? (U.W[amax] > 0 ? avg(amax) : U.raw[amax]) : U.raw[amax];
        U.cnew[tid] = cn;
        float t = fsub(cn, st->c[tid]);
        float r = fmul(t, t);
        float sh = (float)__dsqrt_rn((double)r);
        U.sq[tid] = fmul(sh, sh);
    }
    __syncthreads();
    // ---- 7. convergence (_kmeans.py:721-738)
    if (tid == 0) {
        float tot = np_pairwise_small(U.sq, k);
        int stop = 0, strict = 0;
        if (U.same) {
            stop = 1;
            strict = 1;
        } else if (tot <= st->tol) {
            stop = 1;
        }
        U.same = strict;
        U.stop_reloc = stop;
    }
    __syncthreads();
    {
        const int strict = U.same, stop = U.stop_reloc;
        const int last_iter = st->iter + 1 >= st->max_iter;
        if (tid < k) {
            // labels of a strict stop belong to the centroids the E-step used; otherwise a final E-step
            // with the new centroids follows (emit.cu)
            if (stop || last_iter) st->c_emit[tid] = strict ? st->c[tid] : U.cnew[tid];
            st->c[tid] = U.cnew[tid];
        }
        if (tid == 0) {
            st->iter += 1;
            if (stop || last_iter) {
                st->done = 1;
                st->strict = strict;
                st->n_iter = st->iter;
            }
        }
    }
}
template <int PHASE_LO, int PHASE_HI>
__global__ void __launch_bounds__(TB_THREADS) ll_update_kernel(LloydDevice *st, const float *__restrict__ ks, PeerComm pc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (st->done) return;
    update_phase(st, ks, PHASE_LO, PHASE_HI, pc, *reinterpret_cast<UpdateSmem *>(smem_raw));
}

// ---- final labelling histogram ---------------------------------------------------------------------
// After the loop has stopped, the labels the emission pass will produce are those of the E-step against c_emit.
// Their per-cluster counts follow from one more table / search / zone round on the sorted survivors -- far
// cheaper than histogramming n labels in the streaming emission kernel.
__device__ void final_begin_phase(LloydDevice *st, int clear_done) {
    const int tid = threadIdx.x;
    if (tid < st->k) {
        st->c_save[tid] = st->c[tid];
        st->c[tid] = st->c_emit[tid];
        st->hist[tid] = 0;
    }
    __syncthreads();
    if (tid == 0 && clear_done) st->done = 0;  // the per-phase kernels skip their work while `done` is set
}
__global__ void __launch_bounds__(TB_KMAX) ll_final_begin_kernel(LloydDevice *st) { final_begin_phase(st, 1); }

__device__ void count_phase(LloydDevice *st, long long *Wd /*[TB_KMAX], shared*/) {
    const RegionTable &T = st->tab;
    const int tid = threadIdx.x, k = st->k, m = T.m, R = T.R;
    if (tid < m) Wd[tid] = st->zW[tid];
    __syncthreads();
    for (int r = tid; r < R; r += TB_THREADS) {
        if (T.rJ1[r] != T.rJ2[r]) continue;
        const long long c = st->rcnt[r + 1] - st->rcnt[r];
        if (c > 0) Wd[T.rJ1[r]] += c;  // (J, J) occurs in at most one region
    }
    __syncthreads();
    if (tid == 0 && st->n0 > 0) Wd[zone_argmin(fsub(0.f, st->mean), T.dv, T.dcn, T.down, 0, m - 1)] += st->n0;
    __syncthreads();
    if (tid < m) st->hist[T.down[tid]] = Wd[tid];
    if (tid < k) st->c[tid] = st->c_save[tid];
    __syncthreads();
    if (tid == 0) st->done = 1;
}
__global__ void __launch_bounds__(TB_THREADS) ll_count_kernel(LloydDevice *st) {
    __shared__ long long Wd[TB_KMAX];
    count_phase(st, Wd);
}

// ---- the whole loop in ONE launch -------------------------------------------------------------------------------
// Launched cooperatively with one CTA per SM (all resident).  Per iteration: CTA 0 builds the region table; all CTAs
// search the region boundaries; all CTAs evaluate the zones; CTA 0 updates the centroids (with the in-kernel peer
// exchange on several GPUs) and, unless the loop has stopped, builds the next table right away -- three grid barriers
// per iteration, no launch boundary, no host round trip until the loop has converged.  The code histogram of the final
// labelling (one more table / search / zone round) runs in the same launch.
//
// Grid barrier: every CTA's thread 0 publishes its arrival with a release-ordered atomic and spins with acquire loads
// until all CTAs of the current epoch have arrived; the surrounding __syncthreads extend the ordering to the whole
// CTA (the acquire invalidates the SM's L1, so the plain loads of the next phase see the other CTAs' stores).
// A CTA that waits for minutes gives up and raises *bail (checked by the host): the kernel then runs
// through without waiting instead of hanging the device.
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &epoch, volatile int *bail) {
    __syncthreads();
    epoch += 1;
    if (threadIdx.x == 0 && !*bail) {
        const unsigned int target = epoch * gridDim.x;
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned int v;
        long long spins = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            // longer than the longest wait of CTA 0 inside a peer exchange (peer.cuh: 2^26 polls): a CTA must not give up
            // here while CTA 0 is still, legitimately, waiting for a slow rank
            if (++spins > (1ll << 27)) {
                *bail = 1;
                break;
            }
            __nanosleep(64);  // 148 pollers on one line: leave the L2 slice room for the arrivals
        }
    }
    __syncthreads();
}

union LoopSmem {
    TableScratch tb;
    ZoneSmem zn;
    UpdateSmem up;
    long long wd[TB_KMAX];
};

__global__ void __launch_bounds__(TB_THREADS, 1) ll_loop_kernel(LloydDevice *st, const float *__restrict__ ks,
                                                               const float *__restrict__ samp,
                                                               const long long *__restrict__ ptile, int max_iter, int want_hist,
                                                               PeerComm pc) {
    extern __shared__ __align__(16) unsigned char loop_smem_raw[];
    LoopSmem &S = *reinterpret_cast<LoopSmem *>(loop_smem_raw);
    __shared__ float s_top[LL_TOP];
    SearchConst C = search_const(st, ks, samp, ptile);
    if (C.n_tiles > 64) {  // top level of the tile-sample index in shared memory (the samples never change)
        C.top_step = (C.n_tiles + LL_TOP - 1) / LL_TOP;
        C.top_n = (int)((C.n_tiles + C.top_step - 1) / C.top_step);
        for (int i = threadIdx.x; i < C.top_n; i += blockDim.x) s_top[i] = samp[(long long)i * C.top_step];
        C.top = s_top;
    }
    unsigned int epoch = 0;
    unsigned int *bar = &st->gbar;
    volatile int *bail = &st->gbail;
    const volatile int *done = &st->done;
    if (blockIdx.x == 0) table_phase(st, S.tb);
    grid_barrier(bar, epoch, bail);
    int stopped = 0;
    auto now = []() {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    };
    const bool logger = blockIdx.x == 0 && threadIdx.x == 0;
    for (int it = 0; it < max_iter && !stopped; ++it) {
        const unsigned long long t0 = logger ? now() : 0ull;
        search_phase(st, C);
        grid_barrier(bar, epoch, bail);
        const unsigned long long t1 = logger ? now() : 0ull;
        zone_phase(st, ks, S.zn);
        grid_barrier(bar, epoch, bail);
        const unsigned long long t2 = logger ? now() : 0ull;
        unsigned long long t3 = 0;
        if (blockIdx.x == 0) {
            update_phase(st, ks, 0, 2, pc, S.up);
            __syncthreads();
            t3 = logger ? now() : 0ull;
            if (!*done) table_phase(st, S.tb);
        }
        grid_barrier(bar, epoch, bail);
        if (logger && it < LL_LOG) {
            st->logT[it][0] = (unsigned int)(t1 - t0);
            st->logT[it][1] = (unsigned int)(t2 - t1);
            st->logT[it][2] = (unsigned int)(t3 - t2);
            st->logT[it][3] = (unsigned int)(now() - t3);
        }
        stopped = *done || *bail;  // `done` is only written by the update phase: stable until every CTA has read it
    }
    if (logger) {  // cost of the bare grid barrier (diagnostics)
        const unsigned long long tb = now();
        st->logT[LL_LOG - 1][0] = (unsigned int)tb;
    }
    grid_barrier(bar, epoch, bail);
    grid_barrier(bar, epoch, bail);
    grid_barrier(bar, epoch, bail);
    grid_barrier(bar, epoch, bail);
    if (logger) st->logT[LL_LOG - 1][1] = (unsigned int)now();
    if (!want_hist || !stopped || *bail) return;
    // code histogram of the final labelling (`done` stays set: nobody reads it from here on)
    if (blockIdx.x == 0) {
        final_begin_phase(st, 0);
        __syncthreads();
        table_phase(st, S.tb);
    }
    grid_barrier(bar, epoch, bail);
    search_phase(st, C);
    grid_barrier(bar, epoch, bail);
    zone_phase(st, ks, S.zn);
    grid_barrier(bar, epoch, bail);
    if (blockIdx.x == 0) count_phase(st, S.wd);
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
LloydResult lloyd_run(nnc_ctx *ctx, LloydHandle &h, const float *h_init, int max_iter, double tol_rel,
                      float *h_centred_final, float *h_centred_emit, int64_t *h_hist) {
    const int k = h.k;
    if (k < 1 || k > TB_KMAX) NNC_FAIL(NNC_ERR_UNSUPPORTED, "k-means: k = %d outside [1, %d]", k, TB_KMAX);
    func_dyn_smem(ctx, (const void *)ll_update_kernel<0, 2>, sizeof(UpdateSmem));
    func_dyn_smem(ctx, (const void *)ll_update_kernel<0, 0>, sizeof(UpdateSmem));
    func_dyn_smem(ctx, (const void *)ll_update_kernel<1, 1>, sizeof(UpdateSmem));
    func_dyn_smem(ctx, (const void *)ll_update_kernel<2, 2>, sizeof(UpdateSmem));
    func_dyn_smem(ctx, (const void *)ll_zone_kernel, sizeof(ZoneSmem));
    LloydDevice *st = arena_alloc_t<LloydDevice>(ctx, 1);
    h.d_state = st;
    const long long n_ent = h.d_cnt ? h.n_ent : h.n_nz;  // entries of the sorted array
    const long long n_tiles = (n_ent + LL_TS - 1) / LL_TS;
    h.n_tiles = n_tiles;
    long long *tsum = arena_alloc_t<long long>(ctx, n_tiles + 1);
    long long *ptile = arena_alloc_t<long long>(ctx, n_tiles + 2);
    long long *tcnt = h.d_cnt ? arena_alloc_t<long long>(ctx, n_tiles + 1) : nullptr;
    long long *ctile = h.d_cnt ? arena_alloc_t<long long>(ctx, n_tiles + 2) : nullptr;
    float *samp = arena_alloc_t<float>(ctx, n_tiles + 1);
    float *d_init = arena_alloc_t<float>(ctx, k);
    h.d_ptile = ptile;
    h.d_samp = samp;

    // fixed-point exponent from the data range (zeros included through min/max over all elements)
    const DevScalars &sc = *ctx->h_scal;
    const float mean = h.mean;
    volatile float xmin = ord2f(sc.min_ord) - mean, xmax = ord2f(sc.max_ord) - mean;
    float xabs = fmaxf(fabsf(xmin), fabsf(xmax));
    int E = xabs > 0.f ? ilogbf(xabs) + 1 : 0;
    double scale = ldexp(1.0, 30 - E);

    LloydHeader hs;
    memset(&hs, 0, sizeof(hs));
    hs.k = k;
    hs.max_iter = max_iter;
    hs.fixed_exp = E;
    hs.n = h.n;
    hs.n_nz = h.n_nz;
    hs.n_ent = n_ent;
    hs.cnt = h.d_cnt;
    hs.ctile = ctile;
    hs.n0 = h.n0;
    hs.n_tiles = n_tiles;
    hs.mean = mean;
    hs.xabs_max = xabs;
    hs.scale = scale;
    hs.tol_rel = tol_rel;
    const int world = ctx->world;
    if (world > 64) NNC_FAIL(NNC_ERR_UNSUPPORTED, "k-means: at most 64 ranks");
    hs.rank = ctx->rank;
    hs.world = world;
    hs.cand = arena_alloc_t<unsigned long long>(ctx, (size_t)world * k * 2);
    NNC_CUDA(cudaMemsetAsync(st, 0, sizeof(LloydDevice), ctx->stream));
    NNC_CUDA(cudaMemcpyAsync(static_cast<LloydHeader *>(st), &hs, sizeof(hs), cudaMemcpyHostToDevice, ctx->stream));
    NNC_CUDA(cudaMemcpyAsync(d_init, h_init, sizeof(float) * k, cudaMemcpyHostToDevice, ctx->stream));
    if (n_tiles > 0) {
        int grid = (int)std::min<long long>((long long)ctx->sm_count * 8, (n_tiles + 7) / 8);
        const int sh = 30 - E;  // 2^(30-E) as a float when it is a normal one (fixed_qf), else the float64 scaling
        float scale_f = 0.f;
        if (sh >= -126 && sh <= 127) scale_f = ldexpf(1.0f, sh);
        NNC_LAUNCH(ctx, ll_tilesum_kernel, grid, 256, 0, h.d_sorted, h.d_cnt, n_ent, mean, scale, scale_f, tsum, tcnt, samp, st);
    }
    exclusive_scan_i64(ctx, tsum, n_tiles, ptile);
    if (ctile) exclusive_scan_i64(ctx, tcnt, n_tiles, ctile);
    const bool peer = world > 1 && ctx->peer_enabled && 2 * k <= PEER_WORDS;
    // the cluster kernel (lloyd_fast.cu) folds the scalar prologue (total, moments, their all-reduce, init) into its start
    const bool fast = lloyd_fast_applicable(k) && (world == 1 || peer) && !getenv("NNC_LLOYD_MULTI_LAUNCH") && !getenv("NNC_LLOYD_SPLIT");
    if (!fast) {
        NNC_LAUNCH(ctx, ll_total_kernel, 1, 1, 0, ptile, n_tiles, st);
        NNC_LAUNCH(ctx, ll_moments_kernel, 1, 1, 0, st);
        comm_allreduce(ctx, reinterpret_cast<int64_t *>(&st->s1), 3, 0);  // s1, s2_lo, s2_hi are consecutive
        NNC_LAUNCH(ctx, ll_init_kernel, 1, TB_KMAX, 0, st, d_init);
    }
    prof_mark(ctx, "lloyd_prep");

    struct Ctl {  // mirrors LloydDevice from `iter` on: one read-back brings the loop's outcome and the barrier bail-out flag
        int iter, done, strict, n_reloc, n_iter, pad;
        unsigned int gbar;
        int gbail;
    } ctl;
    static_assert(offsetof(LloydDevice, gbail) - offsetof(LloydDevice, iter) == offsetof(Ctl, gbail), "Ctl mirrors LloydDevice");
    const int search_grid = std::max(1, std::min(ctx->sm_count * 2, (2 * k + 7) / 8));
    const int zone_grid = ctx->sm_count * 2;
    // NNC_LLOYD_MULTI_LAUNCH: one launch per phase (the path taken when no peer mailbox is available);
    // NNC_LLOYD_SPLIT: additionally the update in its three launches with the all-reduces in between (no-ops on one rank)
    const bool split_update = getenv("NNC_LLOYD_SPLIT") != nullptr;
    const bool one_launch = (world == 1 || peer) && !getenv("NNC_LLOYD_MULTI_LAUNCH") && !split_update;
    PeerComm pc;
    memset(&pc, 0, sizeof(pc));
    if (peer) {
        pc.enabled = 1;
        pc.rank = ctx->rank;
        pc.world = world;
        for (int r = 0; r < world; ++r) pc.mail[r] = static_cast<unsigned long long *>(ctx->peer_mail[r]);
    }
    bool hist_done = false;
    if (fast) {
        const bool kt = ctx->ktime && (ctx->kfilter.empty() || strstr("ll_fast_kernel", ctx->kfilter.c_str()));
        if (kt) klaunch_begin(ctx, "ll_fast_kernel");
        lloyd_fast_launch(ctx, st, h.d_sorted, samp, ptile, d_init, h_hist ? 1 : 0, pc, k);
        if (kt) klaunch_end(ctx);
        NNC_CUDA(cudaMemcpyAsync(&ctl, &st->iter, sizeof(Ctl), cudaMemcpyDeviceToHost, ctx->stream));
        if (world > 1) NNC_CUDA(cudaStreamSynchronize(ctx->stream));  // one rank: the outcome is read with the results below (one sync)
        hist_done = h_hist != nullptr;
    } else if (one_launch) {
        func_dyn_smem(ctx, (const void *)ll_loop_kernel, sizeof(LoopSmem));
        const float *ks_arg = h.d_sorted;
        const float *samp_arg = samp;
        const long long *ptile_arg = ptile;
        int max_iter_arg = max_iter, want_hist = h_hist ? 1 : 0;
        void *args[] = {&st, &ks_arg, &samp_arg, &ptile_arg, &max_iter_arg, &want_hist, &pc};
        const bool kt = ctx->ktime && (ctx->kfilter.empty() || strstr("ll_loop_kernel", ctx->kfilter.c_str()));
        if (kt) klaunch_begin(ctx, "ll_loop_kernel");
        NNC_CUDA(cudaLaunchCooperativeKernel((const void *)ll_loop_kernel, dim3(ctx->sm_count), dim3(TB_THREADS), args, sizeof(LoopSmem),
                                             ctx->stream));
        ctx->launches++;
        if (kt) klaunch_end(ctx);
        NNC_CUDA(cudaMemcpyAsync(&ctl, &st->iter, sizeof(Ctl), cudaMemcpyDeviceToHost, ctx->stream));
        NNC_CUDA(cudaStreamSynchronize(ctx->stream));
        hist_done = h_hist != nullptr;
    } else {
        const int batch = 8;
        int launched = 0;
        for (;;) {
            for (int b = 0; b < batch && launched < max_iter; ++b, ++launched) {
                NNC_LAUNCH(ctx, ll_table_kernel, 1, TB_THREADS, 0, st);
                NNC_LAUNCH(ctx, ll_search_kernel, search_grid, 256, 0, st, h.d_sorted, samp, ptile);
                NNC_LAUNCH(ctx, ll_zone_kernel, zone_grid, 256, sizeof(ZoneSmem), st, h.d_sorted);
                if ((world == 1 || peer) && !split_update) {
                    NNC_LAUNCH(ctx, (ll_update_kernel<0, 2>), 1, TB_THREADS, sizeof(UpdateSmem), st, h.d_sorted, pc);
                } else {
                    // exact integer partials: the sums are identical on every rank and for every rank count
                    NNC_LAUNCH(ctx, (ll_update_kernel<0, 0>), 1, TB_THREADS, sizeof(UpdateSmem), st, h.d_sorted, pc);
                    comm_allreduce(ctx, reinterpret_cast<int64_t *>(st->gW), 2 * TB_KMAX, 0);  // gW, gS
                    NNC_LAUNCH(ctx, (ll_update_kernel<1, 1>), 1, TB_THREADS, sizeof(UpdateSmem), st, h.d_sorted, pc);
                    comm_allreduce(ctx, reinterpret_cast<int64_t *>(hs.cand), world * k * 2, 0);  // all-gather by sum
                    NNC_LAUNCH(ctx, (ll_update_kernel<2, 2>), 1, TB_THREADS, sizeof(UpdateSmem), st, h.d_sorted, pc);
                }
            }
            NNC_CUDA(cudaMemcpyAsync(&ctl, &st->iter, sizeof(Ctl), cudaMemcpyDeviceToHost, ctx->stream));
            NNC_CUDA(cudaStreamSynchronize(ctx->stream));
            if (ctl.done || launched >= max_iter) break;
        }
    }
    const bool deferred_ctl = fast && world == 1;  // ctl arrives with the final synchronize
    if (!deferred_ctl) {
        if (!fast && one_launch && ctl.gbail) NNC_FAIL(NNC_ERR_INTERNAL, "k-means: the grid barrier of the loop kernel timed out (iter %d)", ctl.iter);
        if (!ctl.done) NNC_FAIL(NNC_ERR_INTERNAL, "k-means: loop ended without a stop decision (iter %d)", ctl.iter);
    }
    if (world > 1 && ctx->peer_enabled) {
        int comm_error = fast ? ctl.pad : 0;  // the cluster kernel reports it with the loop's outcome
        if (!fast) NNC_CUDA(cudaMemcpy(&comm_error, &st->comm_error, sizeof(int), cudaMemcpyDeviceToHost));
        if (comm_error) {
            ctx->peer_enabled = false;  // the ranks' exchange counters may differ now: no further in-kernel exchanges on this context
            NNC_FAIL(NNC_ERR_COMM, "k-means: a rank did not arrive at an in-kernel peer exchange (time-out); peer exchange disabled on this context");
        }
    }
    NNC_CUDA(cudaMemcpyAsync(h_centred_final, st->c, sizeof(float) * k, cudaMemcpyDeviceToHost, ctx->stream));
    NNC_CUDA(cudaMemcpyAsync(h_centred_emit, st->c_emit, sizeof(float) * k, cudaMemcpyDeviceToHost, ctx->stream));
    float tol_h = 0.f;
    NNC_CUDA(cudaMemcpyAsync(&tol_h, &st->tol, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (h_hist) {
        if (!hist_done) {
            NNC_LAUNCH(ctx, ll_final_begin_kernel, 1, TB_KMAX, 0, st);
            NNC_LAUNCH(ctx, ll_table_kernel, 1, TB_THREADS, 0, st);
            NNC_LAUNCH(ctx, ll_search_kernel, search_grid, 256, 0, st, h.d_sorted, samp, ptile);
            NNC_LAUNCH(ctx, ll_zone_kernel, zone_grid, 256, sizeof(ZoneSmem), st, h.d_sorted);
            NNC_LAUNCH(ctx, ll_count_kernel, 1, TB_THREADS, 0, st);
        }
        if (!(fast && peer)) comm_allreduce(ctx, reinterpret_cast<int64_t *>(st->hist), k, 0);  // (the cluster kernel summed it over the mailboxes)
        static_assert(sizeof(long long) == sizeof(int64_t), "histogram element size");
        NNC_CUDA(cudaMemcpyAsync(h_hist, st->hist, sizeof(int64_t) * k, cudaMemcpyDeviceToHost, ctx->stream));
    }
    NNC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (deferred_ctl && !ctl.done) NNC_FAIL(NNC_ERR_INTERNAL, "k-means: loop ended without a stop decision (iter %d)", ctl.iter);
    prof_mark(ctx, "lloyd_iters");
    if (getenv("NNC_LLOYD_LOG")) {
        const int cnt = std::min(ctl.n_iter, LL_LOG);
        std::vector<long long> z(cnt);
        std::vector<int> g(cnt), mm(cnt), ee(cnt);
        NNC_CUDA(cudaMemcpy(z.data(), st->logZ, sizeof(long long) * cnt, cudaMemcpyDeviceToHost));
        NNC_CUDA(cudaMemcpy(g.data(), st->logG, sizeof(int) * cnt, cudaMemcpyDeviceToHost));
        NNC_CUDA(cudaMemcpy(mm.data(), st->logM, sizeof(int) * cnt, cudaMemcpyDeviceToHost));
        NNC_CUDA(cudaMemcpy(ee.data(), st->logE, sizeof(int) * cnt, cudaMemcpyDeviceToHost));
        std::vector<unsigned int> tt(4 * (size_t)cnt);
        NNC_CUDA(cudaMemcpy(tt.data(), st->logT, sizeof(unsigned int) * 4 * cnt, cudaMemcpyDeviceToHost));
        unsigned int tb[2];
        NNC_CUDA(cudaMemcpy(tb, st->logT[LL_LOG - 1], sizeof(tb), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[nnc lloyd] four bare grid barriers: %.2f us\n", (tb[1] - tb[0]) * 1e-3);
        long long zp[4], up[12];
        if (atoi(getenv("NNC_LLOYD_LOG")) > 1) {  // per-CTA time stamps of iteration 6 (cluster kernel)
            unsigned long long sl[16 * 8];
            NNC_CUDA(cudaMemcpy(sl, st->logZ, sizeof(sl), cudaMemcpyDeviceToHost));
            for (int c = 0; c < 16; ++c)
                if (sl[8 * c])
                    fprintf(stderr, "[nnc lloyd] cta %2d ns: search %5lld | syncA %5lld | zone %5lld | syncB %5lld | update %5lld | syncC %5lld\n", c,
                            (long long)(sl[8 * c + 1] - sl[8 * c]), (long long)(sl[8 * c + 2] - sl[8 * c + 1]),
                            (long long)(sl[8 * c + 3] - sl[8 * c + 2]), (long long)(sl[8 * c + 4] - sl[8 * c + 3]),
                            (long long)(sl[8 * c + 5] - sl[8 * c + 4]), (long long)(sl[8 * c + 6] - sl[8 * c + 5]));
        }
        NNC_CUDA(cudaMemcpy(zp, st->logZ + LL_LOG - 16, sizeof(zp), cudaMemcpyDeviceToHost));
        NNC_CUDA(cudaMemcpy(up, st->logZ + LL_LOG - 32, sizeof(long long) * 8, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[nnc lloyd] zone profile (cycles): prefix %lld elements %lld flush %lld\n", zp[1], zp[2], zp[3]);
        fprintf(stderr, "[nnc lloyd] update profile (cycles): safe %lld zero-run %lld per-id+empties %lld reloc %lld average %lld conv %lld end %lld\n", up[1], up[2], up[3], up[4], up[5], up[6], up[7]);
        {
            int d3[8];
            NNC_CUDA(cudaMemcpy(d3, st->logG + LL_LOG - 24, sizeof(d3), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[nnc lloyd] iteration 6: zones by tiles between their two boundaries (0, 1, ..., >= 7): %d %d %d %d %d %d %d %d\n", d3[0], d3[1],
                    d3[2], d3[3], d3[4], d3[5], d3[6], d3[7]);
            int dbg[8], d2[8];
            NNC_CUDA(cudaMemcpy(dbg, st->logG + LL_LOG - 8, sizeof(dbg), cudaMemcpyDeviceToHost));
            NNC_CUDA(cudaMemcpy(d2, st->logG + LL_LOG - 16, sizeof(d2), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[nnc lloyd] zones not evaluated: boundary before the data %d, ragged / unaligned %d, too wide %d | chunk pass saw: two-candidate zones %d, generic zones %d\n",
                    d2[5], d2[6], d2[7], d2[0], d2[1]);
            fprintf(stderr, "[nnc lloyd] pair searches: %d zones evaluated, %d zones left to the chunk pass, %d other pairs | E-steps with zs = 0 / 1 / 2: %d %d %d\n",
                    dbg[7], dbg[6], dbg[5], dbg[4], dbg[3], dbg[2]);
        }
        {
            long long pp[5];
            NNC_CUDA(cudaMemcpy(pp, st->logZ + LL_LOG - 80, sizeof(pp), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[nnc lloyd] pair search of warp 0 (cycles): located %lld tiles loaded %lld scanned %lld reduced %lld\n", pp[1], pp[2], pp[3], pp[4]);
        }
        for (int w = 0; w < 2; ++w) {
            NNC_CUDA(cudaMemcpy(up, st->logZ + LL_LOG - 48 - 16 * w, sizeof(up), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[nnc lloyd] update profile of iteration %d (cycles): safe %lld zero-run %lld per-id+empties %lld reloc %lld average %lld conv %lld end %lld | relocation: streams filled %lld pops done %lld (%lld rounds, %lld refills)\n", w, up[1], up[2], up[3], up[4], up[5], up[6], up[7], up[8], up[9], up[10] & 0xffff, up[10] >> 16);
        }
        for (int i = 0; i < cnt; ++i)
            fprintf(stderr, "[nnc lloyd] iter %d zone_elems %lld (%.3f%% of survivors) groups %d distinct %d empty %d | us: search %.1f zone %.1f update %.1f table %.1f\n", i, z[i],
                    h.n_nz ? 100.0 * (double)z[i] / (double)h.n_nz : 0.0, g[i], mm[i], ee[i], tt[4 * i] * 1e-3, tt[4 * i + 1] * 1e-3,
                    tt[4 * i + 2] * 1e-3, tt[4 * i + 3] * 1e-3);
    }
    LloydResult r;
    r.n_iter = ctl.n_iter;
    r.strict = ctl.strict;
    r.n_reloc = ctl.n_reloc;
    r.fixed_exp = E;
    r.tol = tol_h;
    return r;
}

}  // namespace nnc
