// Shared helpers for the B200 batch-SOM kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include "../../include/som_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsom_b200 is written for sm_100a (B200) only"
#endif

namespace somb200 {

constexpr int kKPad = 256;   // neuron padding of the prepared codebook (one UMMA N tile)
constexpr int kDPad = 32;    // feature padding: one 128-byte swizzle row of fp32

__host__ __device__ inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
__host__ __device__ inline int64_t ceil_div(int64_t v, int64_t m) { return (v + m - 1) / m; }

// Workspace carved out of the caller's buffer by prepare_codebook and read by the BMU kernels.
struct WsLayout {
    int    k_pad, d_pad;
    size_t aux_off;   // k_pad floats: |w|^2 (euclidean) / 1/|w| (cosine, SIMT) / bias (tensor-core path)
    size_t bias_off;  // k_pad floats: additive bias of the tensor-core epilogue (+inf on padding)
    size_t whi_off;   // k_pad x d_pad floats: TF32 "hi" part of the scaled codebook
    size_t wlo_off;   // k_pad x d_pad floats: TF32 "lo" part
    int    d_pad64;   // feature padding of the fp16 operand copies: one 128-byte swizzle row of halves
    size_t w16hi_off; // k_pad x d_pad64 halves: fp16 "hi" part of the scaled codebook times 2^b_k
    size_t w16lo_off; // k_pad x d_pad64 halves: fp16 "lo" part
    size_t wsinv_off; // k_pad floats: 2^-b_k, the inverse of the per-neuron power-of-two scale
    size_t cnt_off;   // k_pad int32: exact per-BMU counts of the fused kernel (zero between launches)
    size_t done_off;  // one uint32: CTAs-finished ticket of the fused kernel (zero between launches); the grid
                      // barrier of epoch_tail_kernel lives at +256 (arrivals) and +320 (generation)
    size_t gstat_off; // codebook statistics of the current prepare: [0] bits of max_k amax_k, [1] ~bits of the
                      // smallest non-zero amax_k (both via atomicMax, zeroed by prepare), [2] uniform-scale flag
    size_t amax_off;  // k_pad floats: amax_k = max_c |w'_k[c]|
    size_t total;
};

__host__ inline WsLayout ws_layout(int k, int d) {
    WsLayout L;
    L.k_pad = (int)round_up(k, kKPad);
    L.d_pad = (int)round_up(d, kDPad);
    size_t off = 0;
    L.aux_off = off;  off += round_up((size_t)L.k_pad * 4, 1024);
    L.bias_off = off; off += round_up((size_t)L.k_pad * 4, 1024);
    L.whi_off = off;  off += round_up((size_t)L.k_pad * L.d_pad * 4, 1024);
    L.wlo_off = off;  off += round_up((size_t)L.k_pad * L.d_pad * 4, 1024);
    L.d_pad64 = (int)round_up(d, 64);
    L.w16hi_off = off; off += round_up((size_t)L.k_pad * L.d_pad64 * 2, 1024);
    L.w16lo_off = off; off += round_up((size_t)L.k_pad * L.d_pad64 * 2, 1024);
    L.wsinv_off = off; off += round_up((size_t)L.k_pad * 4, 1024);
    L.cnt_off = off;  off += round_up((size_t)L.k_pad * 4, 1024);
    L.done_off = off; off += 1024;
    L.gstat_off = off; off += 1024;
    L.amax_off = off; off += round_up((size_t)L.k_pad * 4, 1024);
    L.total = off;
    return L;
}

// ---- error plumbing (host) -------------------------------------------------
void set_error(const char *fmt, ...);
int  check_cuda(cudaError_t e, const char *what);

#define SOM_CUDA(call)                                                 \
    do {                                                               \
        int _rc = ::somb200::check_cuda((call), #call);                \
        if (_rc) return _rc;                                           \
    } while (0)

#define SOM_REQUIRE(cond, code, ...)                                   \
    do {                                                               \
        if (!(cond)) { ::somb200::set_error(__VA_ARGS__); return (code); } \
    } while (0)

// ---- device helpers -------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 16-byte vector reduction into global memory (sm_90+): one L2 atomic transaction for 4 floats.
__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// power-of-two factor that brings a row whose largest magnitude is amax into [2^14, 2^15): the fp16
// hi/lo split then keeps 22 significant bits of every element that matters (exact scaling, undone
// exactly in the epilogue).  Zero / non-finite rows are left alone.
__device__ __forceinline__ float pow2_scale_for(float amax) {
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int a = 14 - ilogbf(amax);
    a = a < -100 ? -100 : (a > 100 ? 100 : a);
    return ldexpf(1.f, a);
}

// lexicographic (value, index) minimum: the first minimum wins, as numpy's argmin.
__device__ __forceinline__ void argmin_merge(float &v, int &i, float ov, int oi) {
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
}


// ---- programmatic dependent launch -------------------------------------------------------------------
// The epoch is a serial chain of short kernels around the BMU kernel (stats -> split -> BMU -> tables -> apply ->
// merge); launched with the programmatic-stream-serialization attribute, each one is made resident while its
// predecessor is still running and blocks in pdl_wait() until that grid has completed and flushed, which hides
// the launch latency between them (~6 us per boundary at config 2).  Every kernel launched through launch_pdl
// MUST call pdl_wait() before its first global-memory access.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace somb200
