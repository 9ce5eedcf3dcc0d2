// peer.cuh -- in-kernel exchange between the ranks of one box over NVLink / NVSwitch peer memory.
//
// Every rank owns a "mailbox" in its own HBM, IPC-mapped by all the other ranks:
//     data  [2 parities][world senders][words]      flags [2 parities][world senders]      counter
//     tagged data [2 parities][world senders][words]   (slots of the tagged exchange: they only ever hold tagged words)
// An exchange with sequence number seq: each rank stores its vector into slot [seq & 1][rank] of EVERY mailbox
// (plain 8-byte stores through the peer mapping), fences, then raises flag [seq & 1][rank] = seq everywhere; it
// then waits until all the flags of its OWN mailbox have reached seq and reads the world vectors locally.  Sums
// are taken in rank order, so every rank computes bit-identical results.  seq is the count of exchanges executed so
// far -- every rank executes the same sequence of exchanges, so it is the same number everywhere; it lives in the
// rank's own mailbox (peer_counter) across kernels and calls.  Strictly alternating parities make back-to-back
// exchanges safe: a rank can be at most one exchange ahead of the slowest one (it needs everybody's flag to finish
// an exchange), so it never writes the slot somebody is still reading.
//
// Used by the single-CTA Lloyd update kernel (lloyd.cu): the per-cluster (count, sum) all-reduce and the
// relocation-candidate all-gather happen inside the kernel, with no launch boundary and no NCCL call per iteration.
// Kernels of different ranks run on different GPUs; never emulate several ranks on one GPU with this.
#pragma once
#include <stdint.h>

namespace nnc {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_WORDS = 2048;  // 8-byte words per slot

struct PeerComm {
    int enabled = 0, rank = 0, world = 1, pad = 0;
    unsigned long long *mail[PEER_MAX_WORLD];  // mailbox of every rank, as mapped in THIS process (own entry: local memory)
};

__host__ __device__ inline size_t peer_mailbox_bytes(int world) {
    return sizeof(unsigned long long) * (4 * (size_t)world * PEER_WORDS + 2 * (size_t)world * 16 + 16);
}
// number of exchanges this rank has executed (its own mailbox; read at kernel start, written back at kernel end)
__device__ __forceinline__ unsigned long long *peer_counter(const PeerComm &pc) {
    return pc.mail[pc.rank] + 2 * (size_t)pc.world * PEER_WORDS + 2 * (size_t)pc.world * 16;
}
__device__ __forceinline__ unsigned long long *peer_slot(unsigned long long *mail, int world, int parity, int sender) {
    return mail + ((size_t)parity * world + sender) * PEER_WORDS;
}
__device__ __forceinline__ unsigned long long *peer_slot_tagged(unsigned long long *mail, int world, int parity, int sender) {
    return mail + 2 * (size_t)world * PEER_WORDS + 2 * (size_t)world * 16 + 16 + ((size_t)parity * world + sender) * PEER_WORDS;
}
__device__ __forceinline__ unsigned long long *peer_flag(unsigned long long *mail, int world, int parity, int sender) {
    return mail + 2 * (size_t)world * PEER_WORDS + ((size_t)parity * world + sender) * 16;  // one 128-byte line each
}

__device__ __forceinline__ void st_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Publishes data[0, count) (count <= PEER_WORDS) of this rank and waits for everybody's.  Called by all threads of
// ONE CTA.  Returns false on time-out (a rank never arrived): the caller must flag the error, not hang the GPU.
__device__ inline bool peer_publish_and_wait(const PeerComm &pc, const unsigned long long *data, int count, unsigned long long seq) {
    __shared__ int s_ok;
    const int parity = (int)(seq & 1ull), tid = threadIdx.x, nt = blockDim.x;
    for (int p = 0; p < pc.world; ++p) {
        unsigned long long *dst = peer_slot(pc.mail[p], pc.world, parity, pc.rank);
        for (int i = tid; i < count; i += nt) st_sys_u64(dst + i, data[i]);
    }
    __threadfence_system();
    if (tid == 0) s_ok = 1;
    __syncthreads();
    if (tid < pc.world) st_release_sys_u64(peer_flag(pc.mail[tid], pc.world, parity, pc.rank), seq);
    if (tid < pc.world) {
        const unsigned long long *f = peer_flag(pc.mail[pc.rank], pc.world, parity, tid);
        long long spins = 0;
        while (ld_acquire_sys_u64(f) < seq) {
            if (++spins > (1ll << 26)) {  // tens of seconds: give up instead of hanging the device (shorter than the grid barrier's limit)
                s_ok = 0;
                break;
            }
        }
    }
    __syncthreads();
    return s_ok != 0;
}

// all-reduce (sum, rank order) of data[0, count) in place
__device__ inline bool peer_allreduce_sum(const PeerComm &pc, long long *data, int count, unsigned long long seq) {
    const bool ok = peer_publish_and_wait(pc, reinterpret_cast<const unsigned long long *>(data), count, seq);
    const int parity = (int)(seq & 1ull);
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        long long s = 0;
        for (int r = 0; r < pc.world; ++r) s += (long long)ld_sys_u64(peer_slot(pc.mail[pc.rank], pc.world, parity, r) + i);
        data[i] = s;
    }
    __syncthreads();
    return ok;
}

// all-reduce (sum, rank order) of data[0, count), count <= PEER_WORDS / 2, with the sequence number travelling INSIDE the
// data: every 64-bit value goes out as two 8-byte words (tag | low half, tag | high half), tag = the low bits of seq with
// the top bit set (never the 0 of a fresh mailbox).  An 8-byte store is single-copy atomic, so a word that carries the
// right tag carries its payload: no fence between data and flag, no flag -- one NVLink one-way latency per exchange
// instead of store / system fence / flag / poll / read.  Its slots are separate from the flagged exchange's (a stale
// untagged word must never be taken for a tagged one); parity rule and sequence counter are shared with it (a rank can be
// at most one exchange ahead of the slowest one).
__device__ __forceinline__ unsigned long long peer_tag(unsigned long long seq) { return ((seq & 0x7fffffffull) | 0x80000000ull) << 32; }
// sends data[0, count) of this rank as tagged words to every mailbox (all threads of one CTA)
__device__ __forceinline__ void peer_send_tagged(const PeerComm &pc, const unsigned long long *data, int count, unsigned long long seq) {
    const int parity = (int)(seq & 1ull);
    const unsigned long long tag = peer_tag(seq);
    for (int p = 0; p < pc.world; ++p) {
        unsigned long long *dst = peer_slot_tagged(pc.mail[p], pc.world, parity, pc.rank);
        for (int i = threadIdx.x; i < count; i += blockDim.x) {
            const unsigned long long v = data[i];
            st_sys_u64(dst + 2 * i, tag | (v & 0xffffffffull));
            st_sys_u64(dst + 2 * i + 1, tag | (v >> 32));
        }
    }
}
// value i of the senders g .. g + 3 (those below world), polled out of this rank's own mailbox until all have arrived;
// the loads of the four senders are in flight together.  false: timed out.
__device__ __forceinline__ bool peer_poll4_tagged(const PeerComm &pc, unsigned long long seq, int i, int g, unsigned long long *vals) {
    const int parity = (int)(seq & 1ull);
    const unsigned long long tag = peer_tag(seq);
    unsigned long long lo[4], hi[4];
    unsigned pending = (1u << min(4, pc.world - g)) - 1u;
    long long spins = 0;
    while (pending) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (pending >> r & 1u) {
                const unsigned long long *src = peer_slot_tagged(pc.mail[pc.rank], pc.world, parity, g + r) + 2 * i;
                lo[r] = ld_sys_u64(src);
                hi[r] = ld_sys_u64(src + 1);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if ((pending >> r & 1u) && (lo[r] & 0xffffffff00000000ull) == tag && (hi[r] & 0xffffffff00000000ull) == tag) {
                pending &= ~(1u << r);
                vals[r] = (hi[r] << 32) | (lo[r] & 0xffffffffull);
            }
        if (pending && ++spins > (1ll << 24)) return false;  // ~ seconds: give up instead of hanging the device
    }
    return true;
}

__device__ inline bool peer_allreduce_sum_tagged(const PeerComm &pc, long long *data, int count, unsigned long long seq) {
    __shared__ int s_ok_t;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_ok_t = 1;
    peer_send_tagged(pc, reinterpret_cast<const unsigned long long *>(data), count, seq);
    __syncthreads();  // everybody has read data[] (it is overwritten below), s_ok_t is set
    for (int i = tid; i < count; i += nt) {
        long long sum = 0;
        for (int g = 0; g < pc.world; g += 4) {
            unsigned long long v[4] = {0ull, 0ull, 0ull, 0ull};
            if (!peer_poll4_tagged(pc, seq, i, g, v)) s_ok_t = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (g + r < pc.world) sum += (long long)v[r];  // rank order
        }
        data[i] = sum;
    }
    __syncthreads();
    return s_ok_t != 0;
}

// all-gather with tagged words: out[r * stride + i] = rank r's data[i], i < count <= PEER_WORDS / 2 (out: any memory)
__device__ inline bool peer_allgather_tagged(const PeerComm &pc, const unsigned long long *data, int count, unsigned long long *out,
                                             size_t stride, unsigned long long seq) {
    __shared__ int s_ok_g;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_ok_g = 1;
    peer_send_tagged(pc, data, count, seq);
    __syncthreads();
    for (int i = tid; i < count; i += nt) {
        for (int g = 0; g < pc.world; g += 4) {
            unsigned long long v[4] = {0ull, 0ull, 0ull, 0ull};
            if (!peer_poll4_tagged(pc, seq, i, g, v)) s_ok_g = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (g + r < pc.world) out[(size_t)(g + r) * stride + i] = v[r];
        }
    }
    __syncthreads();
    return s_ok_g != 0;
}

// all-gather: out[r * stride + i] = rank r's data[i], i < count
__device__ inline bool peer_allgather(const PeerComm &pc, const unsigned long long *data, int count, unsigned long long *out,
                                      size_t stride, unsigned long long seq) {
    const bool ok = peer_publish_and_wait(pc, data, count, seq);
    const int parity = (int)(seq & 1ull);
    for (int r = 0; r < pc.world; ++r)
        for (int i = threadIdx.x; i < count; i += blockDim.x)
            out[r * stride + i] = ld_sys_u64(peer_slot(pc.mail[pc.rank], pc.world, parity, r) + i);
    __syncthreads();
    return ok;
}

}  // namespace nnc
