// The one exchange step of the sharded path, fused into the epoch tail: exact accumulators in NVLink peer memory.
//
// The reference sums the per-block partial updates with a Dask `sum` (xpysom.py:545-558); with one process per GPU
// that is ONE sum of the per-BMU accumulators [S | counts] across the ranks per epoch.  For the small buffers of
// most maps (config 2: 0.5 MB of int64) an NCCL all-reduce is pure latency plus a stream hand-over in the middle of
// the epoch's launch chain.  Here the accumulator of every rank lives in a "mailbox" in its own HBM that all peers of
// the node have opened through CUDA IPC:
//
//     [0, 512)          flags[2][world]  u32   written by the peers: "rank r has finished accumulating exchange seq"
//     [1024, ...)       acc[2][words]    u64   this rank's exact accumulators, alternating by exchange parity
//
// and the kernel that rounds the integer sums to the fp32 S, c of the neighbourhood apply (phase 0 of
// epoch_tail_kernel, or accum_finalize_kernel on large maps) reads the accumulators of ALL ranks directly:
//   1. one block tells every peer that this rank's BMU kernel of exchange `seq` is complete (stream order; release,
//      system scope);
//   2. every block waits until its own mailbox holds all `world` flags of `seq` (acquire, system scope);
//   3. every element is the INTEGER sum over the ranks' accumulators (remote loads, all in flight at once) -- exact,
//      so every rank, and one GPU holding all the rows, produce the same bits without any ordering rule;
//   4. the rank clears its accumulator of the PREVIOUS exchange (the other parity): a peer publishes `seq` only
//      after it has finished reading `seq - 1`, so nobody can still be reading it, and the next epoch accumulates
//      into a clean buffer while slower peers are still reading this epoch's.
// No separate all-reduce kernel, no NCCL call, no extra copy.  Waits are bounded: a missing peer traps (a CUDA
// error on the host), it never hangs the GPU.
#pragma once
#include <string.h>
#include "common.cuh"

namespace somb200 {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_FLAGS_OFF = 0, PEER_DATA_OFF = 1024;

// what the finalize phase needs on the device (passed by value; world == 0: not sharded this way)
struct PeerView {
    int world = 0, rank = 0;
    unsigned seq = 0;
    int par = 0;                       // flag row and accumulator of this exchange
    size_t words = 0;                  // 64-bit words per accumulator
    uint8_t *p[PEER_MAX_WORLD] = {};
};

struct PeerComm {
    int world = 0, rank = 0;
    size_t words = 0, bytes = 0;
    uint8_t *mailbox = nullptr;
    uint8_t *table[PEER_MAX_WORLD] = {};
    bool opened[PEER_MAX_WORLD] = {};
    unsigned seq = 0;                  // exchanges completed so far
    bool connected = false;
    unsigned long long *acc(int par) const {
        return reinterpret_cast<unsigned long long *>(mailbox + PEER_DATA_OFF) + (size_t)par * words;
    }
    int next_par() const { return (int)((seq + 1u) & 1u); }
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_peer_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ const unsigned long long *peer_acc(const PeerView &V, int r) {
    return reinterpret_cast<const unsigned long long *>(V.p[r] + PEER_DATA_OFF) + (size_t)V.par * V.words;
}

// steps 1 and 2; every thread of the grid calls it before reading any accumulator
__device__ __forceinline__ void peer_exchange_begin(const PeerView &V) {
    if (blockIdx.x == 0 && (int)threadIdx.x < V.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned *>(V.p[threadIdx.x] + PEER_FLAGS_OFF) + V.par * V.world + V.rank, V.seq);
    }
    if ((int)threadIdx.x < V.world) {
        const unsigned *f = reinterpret_cast<const unsigned *>(V.p[V.rank] + PEER_FLAGS_OFF) + V.par * V.world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != V.seq) {
            if (clock64() - t0 > 20000000000LL) {            // ~10 s: a peer never arrived
                printf("som_b200: peer exchange timeout (rank %d waiting for rank %d, seq %u, saw %u)\n", V.rank,
                       (int)threadIdx.x, V.seq, ld_acquire_sys(f));
                __trap();
            }
        }
    }
    __syncthreads();
}

// step 3 for two adjacent elements (16-byte remote loads): the integer sums over the ranks, four loads in flight per
// thread and round
__device__ __forceinline__ ulonglong2 ld_peer_u64x2(const unsigned long long *p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ ulonglong2 peer_sum2(const PeerView &V, size_t e) {
    ulonglong2 s = make_ulonglong2(0ull, 0ull);
    for (int r = 0; r < V.world; r += 4) {
        ulonglong2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (r + u < V.world) ? ld_peer_u64x2(peer_acc(V, r + u) + e) : make_ulonglong2(0ull, 0ull);
        s.x += (v[0].x + v[1].x) + (v[2].x + v[3].x);
        s.y += (v[0].y + v[1].y) + (v[2].y + v[3].y);
    }
    return s;
}

// The finalize phase shared by accum_finalize_kernel and phase 0 of epoch_tail_kernel (grid-stride over `nthr`
// threads): int64 sums -> the fp32 S (K, D) and c (K); double(S_int) * 2^-q is exact up to 2^53, then rounded ONCE.
// Local accumulator (V.world == 0): its `reps` replicas are summed and cleared as they are read.  Peer accumulators: summed over the ranks, the OTHER
// parity's local accumulator is cleared (step 4).
__device__ __forceinline__ void accum_finalize_elements(unsigned long long *Si, int reps, size_t rep_words, const float *qinv,
                                                        int k, int d, int lds, float *S, float *c, int clear, const PeerView &V,
                                                        int64_t tid, int64_t nthr) {
    const int64_t tot = (int64_t)k * lds, all = tot + k;
    if (V.world > 0) {
        peer_exchange_begin(V);
        unsigned long long *old = reinterpret_cast<unsigned long long *>(V.p[V.rank] + PEER_DATA_OFF) + (size_t)(V.par ^ 1) * V.words;
        // pairs of words (the accumulators are 16-byte aligned and `words` is even; a word past the last count is padding)
        for (int64_t e = 2 * tid; e < all; e += 2 * nthr) {
            const ulonglong2 s2 = peer_sum2(V, (size_t)e);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t eh = e + h;
                if (eh >= all) break;
                const unsigned long long u = h ? s2.y : s2.x;
                if (eh < tot) {
                    const int row = (int)(eh / lds), col = (int)(eh % lds);
                    if (col < d) S[(int64_t)row * d + col] = (float)((double)(long long)u * (double)qinv[col]);
                } else {
                    c[eh - tot] = (float)u;
                }
            }
            const ulonglong2 o = *reinterpret_cast<const ulonglong2 *>(old + e);
            if (o.x | o.y) *reinterpret_cast<ulonglong2 *>(old + e) = make_ulonglong2(0ull, 0ull);
        }
        return;
    }
    for (int64_t e = tid; e < all; e += nthr) {
        unsigned long long u = 0ull;
        for (int r = 0; r < reps; ++r) {                       // the replicas of a local accumulator (common.cuh)
            const unsigned long long w = Si[(size_t)r * rep_words + e];
            u += w;
            if (clear && w) Si[(size_t)r * rep_words + e] = 0ull;
        }
        const long long v = (long long)u;
        if (e < tot) {
            const int row = (int)(e / lds), col = (int)(e % lds);
            if (col < d) S[(int64_t)row * d + col] = (float)((double)v * (double)qinv[col]);
        } else {
            c[e - tot] = (float)u;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
static PeerComm *g_peer_comms[8] = {};          // communicators of this process (one per model being trained)

inline int peer_create(size_t words, int world, int rank, PeerComm **out, void *handle64) {
    SOM_REQUIRE(words > 0 && world >= 2 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world && out && handle64,
                SOM_E_BADARG, "peer_create: bad argument (world %d, rank %d)", world, rank);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    int slot = -1;
    for (int i = 0; i < 8; ++i) if (!g_peer_comms[i]) { slot = i; break; }
    SOM_REQUIRE(slot >= 0, SOM_E_BADARG, "peer_create: too many live communicators in this process");
    PeerComm *c = new PeerComm();
    c->world = world; c->rank = rank;
    c->words = (words + 63) / 64 * 64;
    c->bytes = (size_t)PEER_DATA_OFF + (size_t)2 * c->words * 8;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&c->mailbox), c->bytes);
    if (e == cudaSuccess) e = cudaMemset(c->mailbox, 0, c->bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->mailbox);
    if (e != cudaSuccess) {
        if (c->mailbox) cudaFree(c->mailbox);
        delete c;
        return check_cuda(e, "peer_create (cudaMalloc / cudaIpcGetMemHandle)");
    }
    memcpy(handle64, &h, 64);
    c->table[rank] = c->mailbox;
    g_peer_comms[slot] = c;
    *out = c;
    return 0;
}

inline int peer_connect(PeerComm *c, const void *all_handles) {
    SOM_REQUIRE(c && all_handles && !c->connected, SOM_E_BADARG, "peer_connect: bad argument");
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t *>(all_handles) + (size_t)r * 64, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return check_cuda(e, "cudaIpcOpenMemHandle");
        c->table[r] = static_cast<uint8_t *>(p);
        c->opened[r] = true;
    }
    c->connected = true;
    return 0;
}

// the communicator whose NEXT exchange accumulates into `acc` (nullptr: a plain local accumulator)
inline PeerComm *peer_lookup(const void *acc) {
    if (!acc) return nullptr;
    for (int i = 0; i < 8; ++i) {
        PeerComm *c = g_peer_comms[i];
        if (c && c->connected && acc == (const void *)c->acc(c->next_par())) return c;
    }
    return nullptr;
}

// the view of the next exchange; the caller launches exactly one finalize phase with it
inline PeerView peer_next_exchange(PeerComm *c) {
    PeerView V;
    c->seq += 1;
    if (c->seq == 0) c->seq = 2;                                  // wrap-around keeps the parity alternating
    V.world = c->world; V.rank = c->rank; V.seq = c->seq; V.par = (int)(c->seq & 1u); V.words = c->words;
    for (int r = 0; r < c->world; ++r) V.p[r] = c->table[r];
    return V;
}

// The caller guarantees (with a barrier over the ranks) that no peer is still reading this rank's mailbox.
inline int peer_destroy(PeerComm *c) {
    if (!c) return 0;
    cudaDeviceSynchronize();
    for (int i = 0; i < 8; ++i) if (g_peer_comms[i] == c) g_peer_comms[i] = nullptr;
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->table[r]);
    if (c->mailbox) cudaFree(c->mailbox);
    delete c;
    return 0;
}

}  // namespace somb200
