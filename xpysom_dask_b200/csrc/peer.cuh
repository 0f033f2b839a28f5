// One-shot all-reduce of the per-epoch [S | c] partials over NVLink peer memory (single node).
//
// The reference reduces its per-block partial updates with a Dask `sum` (xpysom.py:574-583); the sharded
// path here needs ONE sum of K*D + K floats per epoch across the ranks.  At config 2 that is 0.26 MB: an NCCL
// all-reduce of that size is pure latency (~0.1 ms per epoch against a 0.42 ms epoch).  Every rank owns a
// "mailbox" in its own HBM, opened by all peers through CUDA IPC:
//
//     [0, 512)        flags[2][world]  u32   (written by the peers: "rank r published sequence number seq")
//     [512, 520)      block counters of the two parities (local)
//     [1024, ...)     buf[2][n_pad]    f32   (this rank's published partials, double-buffered by seq parity)
//
// One kernel per call, on the caller's stream:  copy the local partials into buf[seq & 1]; the last block to
// finish publishes `seq` into every peer's flag row (release, system scope); every block waits until all `world`
// flags of its own mailbox carry `seq` (acquire, system scope); then each block sums its slice over the ranks IN
// RANK ORDER (bitwise identical on every rank) straight from the peers' mailboxes and writes it back in place.
// Double buffering makes one barrier per call enough: a rank overwrites buf[p] for seq + 2 only after it has seen
// every peer's flag for seq + 1, which a peer publishes only after it has finished reading seq.
// Waits are bounded: a missing peer traps (-> CUDA error on the host), it never hangs the GPU.
#pragma once
#include "common.cuh"

namespace somb200 {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_FLAGS_OFF = 0, PEER_CTR_OFF = 512, PEER_DATA_OFF = 1024;
constexpr int PEER_THREADS = 256, PEER_MAX_BLOCKS = 64;       // all blocks must be co-resident (they wait on each other)

struct PeerTable { uint8_t *p[PEER_MAX_WORLD]; };

struct PeerComm {
    int world = 0, rank = 0;
    int64_t floats = 0, n_pad = 0;
    size_t bytes = 0;
    uint8_t *mailbox = nullptr;
    PeerTable table{};
    bool opened[PEER_MAX_WORLD] = {};
    unsigned seq = 0;
    bool connected = false;
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_volatile_f1(const float *p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(PEER_THREADS)
peer_allreduce_kernel(PeerTable T, int world, int rank, float *__restrict__ data, int64_t n, int64_t n_pad, unsigned seq) {
    const int par = (int)(seq & 1u);
    uint8_t *mine = T.p[rank];
    float *mybuf = reinterpret_cast<float *>(mine + PEER_DATA_OFF) + (int64_t)par * n_pad;
    const int64_t n4 = n >> 2;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;

    // 1. publish the local partials
    for (int64_t i = tid; i < n4; i += nthr)
        reinterpret_cast<float4 *>(mybuf)[i] = reinterpret_cast<const float4 *>(data)[i];
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthr) mybuf[i] = data[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *ctr = reinterpret_cast<unsigned *>(mine + PEER_CTR_OFF) + par;
        if (atomicAdd(ctr, 1u) == gridDim.x - 1) {          // last block: everything of this rank is in its mailbox
            *ctr = 0u;                                      // (next used two calls from now)
            __threadfence_system();
            for (int r = 0; r < world; ++r)
                st_release_sys(reinterpret_cast<unsigned *>(T.p[r] + PEER_FLAGS_OFF) + par * world + rank, seq);
        }
    }
    // 2. wait for every rank's flag in MY mailbox
    if (threadIdx.x < world) {
        const unsigned *f = reinterpret_cast<const unsigned *>(mine + PEER_FLAGS_OFF) + par * world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != seq) {
            if (clock64() - t0 > 6000000000LL) {            // ~3 s: a peer never arrived
                printf("som_b200: peer all-reduce timeout (rank %d waiting for rank %d, seq %u, saw %u)\n", rank,
                       (int)threadIdx.x, seq, ld_acquire_sys(f));
                __trap();
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    // 3. sum over the ranks in rank order, in place
    for (int64_t i = tid; i < n4; i += nthr) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float4 v = ld_volatile_f4(reinterpret_cast<const float4 *>(T.p[r] + PEER_DATA_OFF) + (int64_t)par * (n_pad >> 2) + i);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        reinterpret_cast<float4 *>(data)[i] = a;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthr) {
        float a = 0.f;
        for (int r = 0; r < world; ++r)
            a += ld_volatile_f1(reinterpret_cast<const float *>(T.p[r] + PEER_DATA_OFF) + (int64_t)par * n_pad + i);
        data[i] = a;
    }
}

inline int peer_create(int64_t floats, int world, int rank, PeerComm **out, void *handle64) {
    SOM_REQUIRE(floats > 0 && world >= 2 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world && out && handle64,
                SOM_E_BADARG, "peer_create: bad argument (world %d, rank %d)", world, rank);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PeerComm *c = new PeerComm();
    c->world = world; c->rank = rank; c->floats = floats;
    c->n_pad = (floats + 63) / 64 * 64;
    c->bytes = (size_t)PEER_DATA_OFF + (size_t)2 * c->n_pad * sizeof(float);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&c->mailbox), c->bytes);
    if (e == cudaSuccess) e = cudaMemset(c->mailbox, 0, c->bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->mailbox);
    if (e != cudaSuccess) {
        if (c->mailbox) cudaFree(c->mailbox);
        delete c;
        return check_cuda(e, "peer_create (cudaMalloc / cudaIpcGetMemHandle)");
    }
    memcpy(handle64, &h, 64);
    c->table.p[rank] = c->mailbox;
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
        c->table.p[r] = static_cast<uint8_t *>(p);
        c->opened[r] = true;
    }
    c->connected = true;
    return 0;
}

inline int peer_allreduce(PeerComm *c, float *data, int64_t floats, cudaStream_t st) {
    SOM_REQUIRE(c && c->connected && data && floats > 0 && floats <= c->floats, SOM_E_BADARG,
                "peer_allreduce: bad argument (floats %lld, capacity %lld)", (long long)floats, (long long)(c ? c->floats : 0));
    SOM_REQUIRE((reinterpret_cast<uintptr_t>(data) & 15) == 0, SOM_E_SHAPE, "peer_allreduce: data must be 16-byte aligned");
    c->seq += 1;
    if (c->seq == 0) c->seq = 2;                                  // wrap-around keeps the parity alternating
    int blocks = (int)ceil_div(floats, (int64_t)PEER_THREADS * 4);
    if (blocks > PEER_MAX_BLOCKS) blocks = PEER_MAX_BLOCKS;
    peer_allreduce_kernel<<<blocks, PEER_THREADS, 0, st>>>(c->table, c->world, c->rank, data, floats, c->n_pad, c->seq);
    return check_cuda(cudaGetLastError(), "peer_allreduce_kernel launch");
}

inline int peer_destroy(PeerComm *c) {
    if (!c) return 0;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->table.p[r]);
    if (c->mailbox) cudaFree(c->mailbox);
    delete c;
    return 0;
}

}  // namespace somb200
