// Microbenchmark: TMEM read throughput (tcgen05.ld.32x32b.x32) and the cost of the argmin epilogue body
// per 128x256 accumulator tile, with 4 or 8 reader warps.   nvcc -arch=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_dep(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

// mode 0: loads only (xor-reduce so they are not dead); 1: + bias(smem) + 8-way running argmin;
// 2: argmin without bias; 3: value-only min (no index)
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int warps, int tiles, long long *cyc, float *sink) {
    __shared__ uint32_t slot;
    __shared__ __align__(16) float bias[256];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 256) bias[threadIdx.x] = threadIdx.x * 0.001f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = slot;
    float bv[8]; int bi[8]; int li[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < 8; ++a) { bv[a] = 1e30f; bi[a] = 0; }
    uint32_t x = 0;
    long long t0 = 0, t1 = 0;
    if (warp < warps) {
        const int q = warp & 3, h = warp >> 2;
        const int ncols = warps == 8 ? 128 : 256;
        const uint32_t taddr = base + ((uint32_t)(q * 32) << 16) + (warps == 8 ? h * 128 : 0);
        t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            if (MODE == 6) {
                // immediate local index (j + 32 c) written by a predicated move; tile base fixed up once per tile
                float old[8];
#pragma unroll
                for (int a = 0; a < 8; ++a) old[a] = bv[a];
                const float rsg = 1.5f + threadIdx.x * 1e-3f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c * 32 >= ncols) break;
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 32, v);
                    wait_dep(v);
                    const float4 *b4 = reinterpret_cast<const float4 *>(bias + (h * 128 + c * 32) % 256);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b = b4[j4];
                        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = j4 * 4 + e;
                            const float sc = fmaf(bb[e], rsg, __uint_as_float(v[j]));
                            if (sc < bv[j & 7]) { bv[j & 7] = sc; li[j & 7] = c * 32 + j; }
                        }
                    }
                }
#pragma unroll
                for (int a = 0; a < 8; ++a) if (bv[a] != old[a]) bi[a] = t;
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < ncols / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
                wait_dep(v);
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x ^= v[j];
                } else if (MODE == 3) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) bv[j & 7] = fminf(bv[j & 7], __uint_as_float(v[j]));
                } else if (MODE >= 7 && MODE <= 10) {
                    // pipe-balanced updates: compare on the ALU pipe, index (a float) and optionally the value
                    // written by predicated FMA-pipe instructions
                    float *fi = reinterpret_cast<float *>(bi);
                    const float basef = (float)(t * 256 + c * 32);
                    const float4 *b4 = reinterpret_cast<const float4 *>(bias + (h * 128 + c * 32) % 256);
                    float one; asm volatile("mov.f32 %0, 0f3f800000;" : "=f"(one));
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float sc = __uint_as_float(v[j]);
                        if (MODE == 10) sc += reinterpret_cast<const float *>(b4)[j];
                        const float jf = (float)j;
                        if (MODE == 7)
                            asm("{\n .reg .pred p;\n setp.lt.f32 p, %2, %0;\n @p add.f32 %1, %3, %4;\n selp.f32 %0, %2, %0, p;\n}"
                                : "+f"(bv[j & 3]), "+f"(fi[j & 3]) : "f"(sc), "f"(basef), "f"(jf));
                        else if (MODE == 8 || MODE == 10)
                            asm("{\n .reg .pred p;\n setp.lt.f32 p, %2, %0;\n @p add.f32 %1, %3, %4;\n @p fma.rn.f32 %0, %2, %5, 0f80000000;\n}"
                                : "+f"(bv[j & 3]), "+f"(fi[j & 3]) : "f"(sc), "f"(basef), "f"(jf), "f"(one));
                        else
                            asm("{\n .reg .pred p;\n setp.lt.f32 p, %2, %0;\n @p add.f32 %1, %3, %4;\n @p add.f32 %0, %2, 0f80000000;\n}"
                                : "+f"(bv[j & 3]), "+f"(fi[j & 3]) : "f"(sc), "f"(basef), "f"(jf));
                    }
                } else if (MODE == 4 || MODE == 5) {
                    // scaled score (2 FMA-pipe ops) made non-negative, compared as unsigned bits with the
                    // DPX min-with-predicate: 1 ALU op for the value + 1 for the index
                    const float4 *b4 = reinterpret_cast<const float4 *>(bias + (h * 128 + c * 32) % 256);
                    const float rs = 1.5f + threadIdx.x * 1e-3f, crow = 100.f;
                    uint32_t *kv = reinterpret_cast<uint32_t *>(bv);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b = b4[j4], s4 = b4[(j4 + 3) & 7];
                        const float bb[4] = {b.x, b.y, b.z, b.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
                        float sc[4];
                        if (MODE == 4) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) sc[e] = fmaf(__uint_as_float(v[j4 * 4 + e]), ss[e], fmaf(bb[e], rs, crow));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; e += 2) {
                                unsigned long long a2, b2, s2, r2, c2, t2, o2;
                                asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "r"(v[j4 * 4 + e]), "r"(v[j4 * 4 + e + 1]));
                                asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "f"(bb[e]), "f"(bb[e + 1]));
                                asm("mov.b64 %0, {%1, %2};" : "=l"(s2) : "f"(ss[e]), "f"(ss[e + 1]));
                                asm("mov.b64 %0, {%1, %1};" : "=l"(r2) : "f"(rs));
                                asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(crow));
                                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t2) : "l"(b2), "l"(r2), "l"(c2));
                                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(o2) : "l"(a2), "l"(s2), "l"(t2));
                                asm("mov.b64 {%0, %1}, %2;" : "=f"(sc[e]), "=f"(sc[e + 1]) : "l"(o2));
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = j4 * 4 + e;
                            bool keep;
                            kv[j & 7] = __vibmin_u32(kv[j & 7], __float_as_uint(sc[e]), &keep);
                            if (!keep) bi[j & 7] = t * 256 + c * 32 + j;
                        }
                    }
                } else {
                    const float4 *b4 = reinterpret_cast<const float4 *>(bias + (h * 128 + c * 32) % 256);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float4 b = MODE == 1 ? b4[j4] : make_float4(0, 0, 0, 0);
                        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = j4 * 4 + e;
                            const float sc = __uint_as_float(v[j]) + bb[e];
                            if (sc < bv[j & 7]) { bv[j & 7] = sc; bi[j & 7] = t * 256 + c * 32 + j; }
                        }
                    }
                }
            }
        }
        t1 = clock64();
    }
    float s = 0; int si = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) { s += bv[a]; si += bi[a] * 256 + li[a]; }
    if (sink && (x == 0x12345678 || s == 1.2345f)) sink[threadIdx.x] = s + si;
    if (threadIdx.x % 32 == 0 && warp < warps) cyc[blockIdx.x * 8 + warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(512));
}

template <int MODE>
void run(const char *name, int warps) {
    long long *cyc; float *sink;
    cudaMalloc(&cyc, 148 * 8 * 8); cudaMalloc(&sink, 4096);
    const int tiles = 200;
    k<MODE><<<148, 256>>>(warps, tiles, cyc, sink);
    k<MODE><<<148, 256>>>(warps, tiles, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148 * 8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int b = 0; b < 148; ++b) for (int w = 0; w < warps; ++w) mx = h[b * 8 + w] > mx ? h[b * 8 + w] : mx;
    printf("%-28s warps=%d: %7.1f cycles per 128x256 tile  (%.1f B/clk/SM, %.2f elem/clk/SM)  [%s]\n", name, warps,
           (double)mx / tiles, 131072.0 * tiles / mx, 32768.0 * tiles / mx, cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    for (int w = 4; w <= 8; w += 4) {
        run<0>("LDTM only", w);
        run<3>("LDTM + value-only min", w);
        run<2>("LDTM + argmin (no bias)", w);
        run<1>("LDTM + smem bias + argmin", w);
        run<4>("scaled + vibmin argmin", w);
        run<5>("scaled(f32x2) + vibmin", w);
        run<6>("uniform fma + imm-index", w);
        run<7>("setp+selp+fadd.idx", w);
        run<8>("setp+@ffma.val+@fadd.idx", w);
        run<9>("setp+@fadd.val+@fadd.idx", w);
        run<10>("bias + setp+@ffma+@fadd", w);
    }
    return 0;
}
