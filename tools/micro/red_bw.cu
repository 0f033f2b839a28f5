// Microbenchmark: throughput of the per-BMU accumulation S[bmu[r], :] += x[r, :] on B200 for the candidate
// order-independent (exact fixed-point, int64) forms against today's fp32 vector reduction.
//   A  red.global.add.v4.f32                       (fp32, arrival order: the round-1 path)
//   B  red.global.add.u64 per element              (exact: x * 2^q -> int64)
//   C  cp.reduce.async.bulk ... .add.u64 per row    (exact; one TMA bulk reduction per <= 128-column piece of a row,
//                                                    staged as int64 in shared memory)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bw red_bw.cu ;  run: ./red_bw
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long *addr, long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" :: "l"(addr), "l"(v) : "memory");
}

// one warp per row-group: LPR lanes own a row (float4 per lane and pass)
__global__ void __launch_bounds__(128) k_f32v4(const float *X, const int *bmu, int64_t n, int d, float *S) {
    const int lane = threadIdx.x & 31, d4 = d >> 2;
    const int lpr = d4 >= 32 ? 32 : d4, rpp = 32 / lpr, sub = lane / lpr, c0 = lane % lpr;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp * rpp; r0 < n; r0 += warps * rpp) {
        const int64_t r = r0 + sub;
        if (r >= n) continue;
        const int b = __ldg(bmu + r);
        for (int c4 = c0; c4 < d4; c4 += lpr)
            red_add_v4(S + (int64_t)b * d + c4 * 4, __ldg(reinterpret_cast<const float4 *>(X + r * d) + c4));
    }
}

__global__ void __launch_bounds__(128) k_u64(const float *X, const int *bmu, int64_t n, int d, const float *scale,
                                             unsigned long long *S) {
    const int lane = threadIdx.x & 31, d4 = d >> 2;
    const int lpr = d4 >= 32 ? 32 : d4, rpp = 32 / lpr, sub = lane / lpr, c0 = lane % lpr;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp * rpp; r0 < n; r0 += warps * rpp) {
        const int64_t r = r0 + sub;
        if (r >= n) continue;
        const int b = __ldg(bmu + r);
        for (int c4 = c0; c4 < d4; c4 += lpr) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(X + r * d) + c4);
            const float4 s = __ldg(reinterpret_cast<const float4 *>(scale) + c4);
            unsigned long long *dst = S + (int64_t)b * d + c4 * 4;
            red_add_u64(dst + 0, __float2ll_rn(v.x * s.x));
            red_add_u64(dst + 1, __float2ll_rn(v.y * s.y));
            red_add_u64(dst + 2, __float2ll_rn(v.z * s.z));
            red_add_u64(dst + 3, __float2ll_rn(v.w * s.w));
        }
    }
}

// bulk: each warp owns NBUF staging buffers of PIECE columns (int64); a row is sent as ceil(d / PIECE) bulk reductions
constexpr int PIECE = 128, NBUF = 4;
__global__ void __launch_bounds__(128) k_bulk(const float *X, const int *bmu, int64_t n, int d, const float *scale,
                                              unsigned long long *S, int R, size_t rep_stride) {
    S += (size_t)(blockIdx.x % R) * rep_stride;       // R replicas of the accumulator spread hot rows over R L2 lines
    __shared__ __align__(128) long long stage[4][NBUF][PIECE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, d4 = d >> 2;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t it = 0;
    // small rows: several rows per piece-pass (lanes split over rows), each row still its own bulk op
    const int lpr = d4 >= 32 ? 32 : d4, rpp = 32 / lpr, sub = lane / lpr, c0 = lane % lpr;
    for (int64_t r0 = warp * rpp; r0 < n; r0 += warps * rpp) {
        const int64_t r = r0 + sub;
        const bool live = r < n;
        const int b = live ? __ldg(bmu + r) : 0;
        for (int p0 = 0; p0 < d; p0 += PIECE, ++it) {
            const int buf = it % NBUF;
            // the buffer was last used NBUF pieces ago: at most NBUF - 1 groups may still be reading
            asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(NBUF - 1) : "memory");
            __syncwarp();
            const int cols = min(PIECE, d - p0);             // columns of this piece
            long long *st = stage[wib][buf];
            // rows of this pass are packed in the buffer: row `sub` at offset sub * cols (only when rpp > 1, d <= 128)
            long long *mine = st + (rpp > 1 ? sub * d : 0);
            if (live)
                for (int c4 = c0; c4 * 4 < cols; c4 += lpr) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(X + r * d + p0) + c4);
                    const float4 s = __ldg(reinterpret_cast<const float4 *>(scale + p0) + c4);
                    longlong2 a, bq;
                    a.x = __float2ll_rn(v.x * s.x); a.y = __float2ll_rn(v.y * s.y);
                    bq.x = __float2ll_rn(v.z * s.z); bq.y = __float2ll_rn(v.w * s.w);
                    reinterpret_cast<longlong2 *>(mine + c4 * 4)[0] = a;
                    reinterpret_cast<longlong2 *>(mine + c4 * 4)[1] = bq;
                }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (live && c0 == 0) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(mine);
                unsigned long long *dst = S + (int64_t)b * d + p0;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;"
                             :: "l"(dst), "r"(src), "r"(cols * 8) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    struct Cfg { int64_t n; int d, k; const char *name; };
    const Cfg cfgs[] = {{1000000, 64, 1024, "c2"}, {2000000, 16, 1600, "c3"}, {200000, 784, 10000, "c4/5"}, {500000, 128, 2500, "c5"}};
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (const Cfg &c : cfgs) {
        for (int hot = 0; hot < 2; ++hot) {
            std::vector<float> hx((size_t)c.n * c.d); std::vector<int> hb(c.n); std::vector<float> hs(c.d);
            uint32_t s = 12345u;
            auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
            for (auto &v : hx) v = (rnd() & 0xffffff) / 16777216.0f;
            for (auto &v : hb) v = hot ? ((rnd() % 10 < 9) ? (int)(rnd() % 16) * (c.k / 16) : (int)(rnd() % c.k)) : (int)(rnd() % c.k);
            for (auto &v : hs) v = 1099511627776.0f;   // 2^40
            float *X, *S32, *scale; int *bmu; unsigned long long *S64;
            CK(cudaMalloc(&X, hx.size() * 4)); CK(cudaMalloc(&bmu, hb.size() * 4)); CK(cudaMalloc(&scale, c.d * 4));
            CK(cudaMalloc(&S32, (size_t)c.k * c.d * 4)); CK(cudaMalloc(&S64, (size_t)c.k * c.d * 8));
            CK(cudaMemcpy(X, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(bmu, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(scale, hs.data(), c.d * 4, cudaMemcpyHostToDevice));
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            const int grid = sms * 4;
            const int RMAX = 16;
            cudaFree(S64); CK(cudaMalloc(&S64, (size_t)c.k * c.d * 8 * RMAX));
            for (int which = 0; which < 6; ++which) {
                const int R = which < 3 ? 1 : which == 3 ? 4 : which == 4 ? 8 : 16;
                float best = 1e30f;
                for (int rep = 0; rep < 4; ++rep) {
                    CK(cudaMemset(S32, 0, (size_t)c.k * c.d * 4)); CK(cudaMemset(S64, 0, (size_t)c.k * c.d * 8 * RMAX));
                    CK(cudaEventRecord(e0));
                    if (which == 0) k_f32v4<<<grid, 128>>>(X, bmu, c.n, c.d, S32);
                    else if (which == 1) k_u64<<<grid, 128>>>(X, bmu, c.n, c.d, scale, S64);
                    else k_bulk<<<grid, 128>>>(X, bmu, c.n, c.d, scale, S64, R, (size_t)c.k * c.d);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep > 0 && ms < best) best = ms;
                }
                // checksum: total of S must equal the total of x (fixed point: exactly)
                double tot = 0;
                if (which == 0) { std::vector<float> h((size_t)c.k * c.d); CK(cudaMemcpy(h.data(), S32, h.size() * 4, cudaMemcpyDeviceToHost)); for (float v : h) tot += v; }
                else { std::vector<long long> h((size_t)c.k * c.d * RMAX); CK(cudaMemcpy(h.data(), S64, h.size() * 8, cudaMemcpyDeviceToHost)); for (long long v : h) tot += (double)v / 1099511627776.0; }
                double ref = 0; for (float v : hx) ref += v;
                printf("%-5s %s  %-22s %8.3f ms  %6.2f elem/clk/SM(@1.965GHz)  %7.1f GB/s of x   sum rel err %.2e\n", c.name,
                       hot ? "hot " : "unif", which == 0 ? "red.v4.f32" : which == 1 ? "red.u64 scalar" : which == 2 ? "bulk reduce u64" : which == 3 ? "bulk u64, 4 replicas" : which == 4 ? "bulk u64, 8 replicas" : "bulk u64, 16 replicas",
                       best, (double)c.n * c.d / (best * 1e-3) / 1.965e9 / sms, (double)c.n * c.d * 4 / (best * 1e-3) / 1e9,
                       (tot - ref) / ref);
            }
            cudaFree(X); cudaFree(bmu); cudaFree(scale); cudaFree(S32); cudaFree(S64);
        }
    }
    return 0;
}
