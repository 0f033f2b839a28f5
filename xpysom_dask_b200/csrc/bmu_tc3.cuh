// K1 (v3): fp16-split tensor-core distance contraction + fused BMU argmin + fused accumulate.
//
// Same contract and CTA-pair structure as bmu_tc2.cuh, different arithmetic: kind::f16 MMAs run
// at twice the kind::tf32 rate and fp16 carries the same 11-bit significand as TF32, so splitting
// every fp32 operand into hi = rn_f16(v), lo = rn_f16(v - hi) and accumulating
// lo*hi + hi*lo + hi*hi in fp32 (TMEM) keeps the 3xTF32 accuracy (|v - hi - lo| <= 2^-22 |v|) at
// half the tensor-pipe time.  fp16's narrow exponent is handled by EXACT power-of-two scaling:
//   x^ = x * 2^a_r   (per sample row, a_r from row_scale_kernel; max |x^| in [2^14, 2^15))
//   w^ = w' * 2^b_k  (per neuron, from prepare_codebook_kernel)
// and the epilogue minimises  acc * 2^-b_k + bias_k * 2^a_r  = 2^a_r (x . w'_k + bias_k), a positive
// row-constant multiple of the score, so the argmin is unchanged.  Elements far below their row's
// maximum fall into fp16's subnormal range and lose RELATIVE precision, but their absolute error
// stays below 2^-39 of the row maximum, far inside the stated BMU epsilon.
//
// Operand staging (per CTA, 32 KB slots, SWIZZLE_128B, 64 features = 128 B of halves per row):
//   A ring (3 slots): the raw fp32 X chunk [128 rows x 64 floats] lands by TMA (two 16 KB boxes);
//       the converter warpgroup reads it into registers, synchronises, and overwrites the slot IN
//       PLACE with the fp16 hi tile (first 16 KB) and lo tile (second 16 KB).
//       When D <= 128 the A tiles of a 256-row tile stay RESIDENT across all neuron tiles
//       (loaded and converted once per row tile); for larger D they stream once per neuron tile.
//   B ring (3 slots): this CTA's half (128 neurons) of the W'hi and W'lo fp16 tiles.
// 12 MMAs (M=256 N=256 K=16 over the CTA pair) per 64-feature block.
//
// Three instantiations of the pipeline (MODE):
//   0  resident   D <= 128: the converted X tiles of a row tile stay in shared memory across its neuron tiles
//   1  streaming  larger D: X chunks stream once per neuron tile
//   2  packed     D <= 16 (resident): ONE 64-column tile carries the three products of the split as three 16-column
//                 K steps -- x^ = [hi | lo | hi | 0] against w^ = [hi | hi | lo | 0] (prepared by codebook_split_kernel):
//                 3 MMAs per neuron tile instead of 7 on the TF32 kernel, half the W' traffic.
// FOLD (resident, one feature block: D <= 64): the bias joins the contraction through ONE extra kind::tf32 MMA per
// neuron tile, acc += u_r * B_k with u_r = 2^a_r (the row scale, exact in TF32) and B_k = bias_k 2^b_k split into three
// TF32 pieces (33 bits; TF32's 8-bit exponent has no range problem where an fp16 fold column would overflow).  The
// epilogue then only compares raw accumulators: 3 instructions per score instead of 4.25 (the epilogue is bound by
// instruction issue: 128 x 256 scores / 32 lanes / 4 schedulers).  The fold operands are 32 bytes per row in the
// no-swizzle K-major core-matrix layout: A_f is written by the converter (one 16-byte store per row and row tile),
// B_f arrives as a 4 KB image per 128 neurons (TMA, prepared by codebook_split_kernel).  With one feature block per
// row tile the A ring needs two slots, the third one's shared memory hosts the fold operands.
//
// Warp roles per CTA (640 threads): 0 B producer | 1 MMA issuer (leader) + TMEM alloc | 2 A producer |
// 3 spare | 4-7 converter | 8-15 epilogue (two warps per TMEM lane quarter) | 16-19 scatter.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "bmu_tc.cuh"
#include "bmu_tc2.cuh"

namespace somb200 {
namespace tc3 {

using tc::FusedAcc;
using tc::smem_u32;
using namespace tc2;   // cluster / 2-SM wrappers

constexpr int BM = 128;                  // sample rows per CTA (TMEM lanes)
constexpr int TBN = 256;                 // neurons per accumulator tile (UMMA N over the pair); two stages in TMEM
constexpr int TBNH = TBN / 2;            // this CTA's share of a neuron tile (B rows)
constexpr int NACC = 2;
constexpr int BK = 64;                   // features per block: 128 bytes of fp16
constexpr int UMMA_K = 16;
constexpr int NA = 3, NB = 3;
constexpr int SLOT_BYTES = 32 * 1024;    // A: raw fp32 [128 x 64] -> hi | lo ;  B: W'hi half | W'lo half
constexpr int HALF_SLOT = 16 * 1024;
constexpr int FOLD_TILE_BYTES = 128 * 32;          // fold operand tile: 128 rows x 8 tf32
constexpr int NAF = 2;                             // A ring slots when the bias is folded (one feature block per row tile):
                                                   // the third slot's shared memory hosts the fold operands
constexpr int NB_PACKED = 6;                       // W' ring of the packed mode: 16 KB slots
constexpr int NUM_THREADS = 640;
constexpr int APROD_WARP = 2, CONV_WARP0 = 4, EPI_WARP0 = 8, SCAT_WARP0 = 16;
constexpr int EPI_THREADS = 256;
constexpr int RESIDENT_MAX_KB = 2;       // D <= 128: X tiles resident across neuron tiles
constexpr int MAXB = NB_PACKED;                    // barrier slots of the W' ring
constexpr int NUM_BARS = 3 * NA + 2 * MAXB + 2 * NACC + 4;
constexpr int EPI_STAGE_BYTES = 8 * 2 * 128 * 4;   // per epilogue warp: 128 bias + 128 inverse-scale floats
constexpr int SMEM_BYTES = (NA + NB) * SLOT_BYTES + EPI_STAGE_BYTES + tc::SCAT_STAGE_BYTES + 2 * BM * 8 + 2 * BM * 4 + NUM_BARS * 8 + 64 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory of the fp16 kernel");
static_assert(NAF * FOLD_TILE_BYTES + NB_PACKED * FOLD_TILE_BYTES <= SLOT_BYTES, "fold operands live in the third A slot");

// no-swizzle ("interleave") K-major shared-memory matrix descriptor: 8-row x 16-byte core matrices; the two K chunks of
// a 32-byte row are LBO bytes apart, consecutive 8-row groups SBO bytes apart (cute: ((8,n),2):((1,SBO),LBO) in 16-byte units)
__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D = f32 (bit 4), A = B = f16 (format 0), K-major, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 over the pair)
__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24); }

// packed fp32x2 arithmetic (sm_100): one instruction for two lanes of a register pair
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t lo, uint32_t hi) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}

// running argmin on the scaled scores: sc = acc * 2^-b_k + bias_k * 2^a_r
struct RunMinScaled : tc::RunMin {
    // uniform codebook scale: sc = acc + bias_k * rsg  (one FFMA per score, one shared-memory operand)
    __device__ __forceinline__ void chunk_uniform(const uint32_t (&acc)[32], const float *bias32, float rsg, int colbase) {
        const float4 *b4 = reinterpret_cast<const float4 *>(bias32);
        const float basef = (float)colbase;
        float4 bq[8];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) bq[j4] = b4[j4];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float bb[4] = {bq[j4].x, bq[j4].y, bq[j4].z, bq[j4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j4 * 4 + e;
                upd(j % tc::EPI_ACC, fmaf(bb[e], rsg, __uint_as_float(acc[j])), basef, (float)j);
            }
        }
    }
    // folded bias, per-neuron codebook scales: sc = acc * 2^-b_k
    __device__ __forceinline__ void chunk_scaled_nobias(const uint32_t (&acc)[32], const float *winv32, int colbase) {
        const float4 *s4 = reinterpret_cast<const float4 *>(winv32);
        const float basef = (float)colbase;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 s = s4[j4];
            const float ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j4 * 4 + e;
                upd(j % tc::EPI_ACC, __uint_as_float(acc[j]) * ss[e], basef, (float)j);
            }
        }
    }
    __device__ __forceinline__ void chunk(const uint32_t (&acc)[32], const float *bias32, const float *winv32,
                                          float rs, int colbase) {
        const float4 *b4 = reinterpret_cast<const float4 *>(bias32);     // shared memory, broadcast reads
        const float4 *s4 = reinterpret_cast<const float4 *>(winv32);
        const uint64_t rs2 = pack2(rs, rs);
        const float basef = (float)colbase;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b = b4[j4], s = s4[j4];
            float sc[4];
            unpack2(fma2(pack2u(acc[j4 * 4 + 0], acc[j4 * 4 + 1]), pack2(s.x, s.y), mul2(pack2(b.x, b.y), rs2)), sc[0], sc[1]);
            unpack2(fma2(pack2u(acc[j4 * 4 + 2], acc[j4 * 4 + 3]), pack2(s.z, s.w), mul2(pack2(b.z, b.w), rs2)), sc[2], sc[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j4 * 4 + e;
                upd(j % tc::EPI_ACC, sc[e], basef, (float)j);
            }
        }
    }
};

// One A item: raw fp32 chunk (two SWIZZLE_128B boxes of [128 rows x 32 floats]) -> scaled fp16 hi | lo tiles,
// in place, by NCONV cooperating threads (bar.sync 2 separates the read of the whole slot from its overwrite).
template <int NCONV>
__device__ __forceinline__ void convert_item(uint8_t *slot, int t, int64_t row0, int64_t n, const float *__restrict__ xscale) {
    constexpr int PER = 2048 / NCONV;          // float4 per thread
    const float4 *raw = reinterpret_cast<const float4 *>(slot);
    float4 v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = raw[t + NCONV * i];               // linear: conflict-free
    asm volatile("bar.sync 2, %0;" :: "n"(NCONV) : "memory");                // everyone has read the slot
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int e = t + NCONV * i;
        const int box = e >> 10, r = (e & 1023) >> 3, cpos = e & 7;
        const int j = cpos ^ (r & 7);                 // logical 16-byte chunk within the 32 floats
        const int k0 = box * 32 + j * 4;              // first of 4 consecutive features
        const int64_t grow = row0 + r;
        const float rs = grow < n ? __ldg(xscale + grow) : 1.f;
        const float x0 = v[i].x * rs, x1 = v[i].y * rs, x2 = v[i].z * rs, x3 = v[i].w * rs;
        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y);
        const __half2 l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
        // fp16 tile [128 rows x 64 halves], SWIZZLE_128B: 16-byte chunk (k0 / 8) ^ (r % 8)
        const int off = r * 128 + (((k0 >> 3) ^ (r & 7)) << 4) + ((k0 & 7) << 1);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t *>(&h01); hv.y = *reinterpret_cast<const uint32_t *>(&h23);
        lv.x = *reinterpret_cast<const uint32_t *>(&l01); lv.y = *reinterpret_cast<const uint32_t *>(&l23);
        *reinterpret_cast<uint2 *>(slot + off) = hv;
        *reinterpret_cast<uint2 *>(slot + HALF_SLOT + off) = lv;
    }
}

// Packed A item (D <= 16): thread t owns row t of the raw box [128 rows x 32 floats] (only the first 16 floats are
// features, the rest is the TMA zero fill) and rewrites the row, in place, as 64 halves [hi | lo | hi | 0] x 16.
__device__ __forceinline__ void convert_item_packed(uint8_t *slot, int t, float rs) {
    const int sw = t & 7;
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4 *>(slot + t * 128 + ((j ^ sw) << 4));
    asm volatile("bar.sync 2, 128;" ::: "memory");                          // everyone has read the box
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float x0 = v[j].x * rs, x1 = v[j].y * rs, x2 = v[j].z * rs, x3 = v[j].w * rs;
        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y);
        const __half2 l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
        hi[2 * j] = *reinterpret_cast<const uint32_t *>(&h01); hi[2 * j + 1] = *reinterpret_cast<const uint32_t *>(&h23);
        lo[2 * j] = *reinterpret_cast<const uint32_t *>(&l01); lo[2 * j + 1] = *reinterpret_cast<const uint32_t *>(&l23);
    }
    // logical 16-byte chunks of the 128-byte row: 0,1 = hi | 2,3 = lo | 4,5 = hi | 6,7 = zero
    uint8_t *row = slot + t * 128;
    const uint4 h0 = make_uint4(hi[0], hi[1], hi[2], hi[3]), h1 = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    const uint4 l0 = make_uint4(lo[0], lo[1], lo[2], lo[3]), l1 = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(row + ((0 ^ sw) << 4)) = h0;
    *reinterpret_cast<uint4 *>(row + ((1 ^ sw) << 4)) = h1;
    *reinterpret_cast<uint4 *>(row + ((2 ^ sw) << 4)) = l0;
    *reinterpret_cast<uint4 *>(row + ((3 ^ sw) << 4)) = l1;
    *reinterpret_cast<uint4 *>(row + ((4 ^ sw) << 4)) = h0;
    *reinterpret_cast<uint4 *>(row + ((5 ^ sw) << 4)) = h1;
    *reinterpret_cast<uint4 *>(row + ((6 ^ sw) << 4)) = z;
    *reinterpret_cast<uint4 *>(row + ((7 ^ sw) << 4)) = z;
}

constexpr int MODE_RESIDENT = 0, MODE_STREAM = 1, MODE_PACKED = 2;

template <int MODE, bool FOLD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
bmu_tc3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
               const __grid_constant__ CUtensorMap map_wlo, const __grid_constant__ CUtensorMap map_fold,
               const float *__restrict__ bias, const float *__restrict__ wsinv, const unsigned int *__restrict__ gstat,
               const float *__restrict__ xscale, int64_t n, int num_pair_tiles, int num_n_tiles, int num_k_blocks,
               int32_t *__restrict__ bmu_out, float *__restrict__ best_out, const FusedAcc acc) {
    static_assert(!(FOLD && MODE == MODE_STREAM), "the bias is folded only with resident X tiles");
    static_assert(!(MODE == MODE_PACKED && !FOLD), "the packed mode always folds");
    pdl_wait(); pdl_trigger();        // programmatic dependent launch (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr int NAR = FOLD ? NAF : NA;                      // A ring slots in use
    constexpr int NBR = MODE == MODE_PACKED ? NB_PACKED : NB; // W' ring slots (packed: one 16 KB tile per slot)
    constexpr int BSLOT = MODE == MODE_PACKED ? HALF_SLOT : SLOT_BYTES;
    const uint32_t a_base = smem_base, b_base = smem_base + NA * SLOT_BYTES;
    // fold operands in the third A slot: A_f ring (one tile per A slot in use), then the B_f ring (one tile per W' slot)
    const uint32_t af_base = a_base + NAF * SLOT_BYTES, bf_base = af_base + NAF * FOLD_TILE_BYTES;
    float *epi_stage = reinterpret_cast<float *>(smem + (NA + NB) * SLOT_BYTES);
    uint8_t *scat_stage = smem + (NA + NB) * SLOT_BYTES + EPI_STAGE_BYTES;                 // [4 warps][SCAT_NBUF][128] int64
    uint8_t *tail = scat_stage + tc::SCAT_STAGE_BYTES;

    float    *mrg_v = reinterpret_cast<float *>(tail);                   // [2][BM]
    int      *mrg_i = reinterpret_cast<int *>(tail + 2 * BM * 4);        // [2][BM]
    int      *bmu_s = reinterpret_cast<int *>(tail + 2 * BM * 8);        // [2][BM]
    uint64_t *bars  = reinterpret_cast<uint64_t *>(tail + 2 * BM * 8 + 2 * BM * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto afull_bar  = [&](int s) { return bar0 + 8u * s; };                       // local: raw X chunk landed
    auto aready_bar = [&](int s) { return bar0 + 8u * (NA + s); };                // leader: both converters done
    auto aempty_bar = [&](int s) { return bar0 + 8u * (2 * NA + s); };            // local: A slot consumed
    auto bfull_bar  = [&](int s) { return bar0 + 8u * (3 * NA + s); };            // leader: both W' halves landed
    auto bempty_bar = [&](int s) { return bar0 + 8u * (3 * NA + MAXB + s); };     // local: B slot consumed
    auto tfull_bar  = [&](int a) { return bar0 + 8u * (3 * NA + 2 * MAXB + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * NA + 2 * MAXB + NACC + a); };
    auto bfullq_bar = [&](int b) { return bar0 + 8u * (3 * NA + 2 * MAXB + 2 * NACC + b); };
    auto bemptyq_bar = [&](int b) { return bar0 + 8u * (3 * NA + 2 * MAXB + 2 * NACC + 2 + b); };
    constexpr uint32_t kIdescF16 = idesc_f16(TBN);
    // Narrow last tile (bias in the epilogue only): when the map is not a multiple of 256 neurons the last tile's MMAs
    // cover just the real neurons rounded up to 16 (UMMA N granularity over a CTA pair) -- 16 instead of 256 columns at
    // 100 x 100 neurons (2.3 % of that kernel's tensor work).  The pair's halves of such a tile are its first
    // n_last / 2 B rows of EACH CTA, so CTA r loads neurons nt * 256 + r * n_last / 2 ...; accumulator column c is
    // still neuron nt * 256 + c.  Columns the MMA did not write hold stale TMEM data, which the +inf bias of the padding
    // neurons turns into +inf / NaN (never a minimum); with the bias folded into the contraction every tile stays full.
    const int n_last = FOLD ? TBN : (int)round_up(acc.k - (num_n_tiles - 1) * TBN, 16);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bars + NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const bool fused = acc.S != nullptr;
    constexpr bool resident = MODE != MODE_STREAM;           // A tiles live across the neuron tiles
    constexpr bool packed = MODE == MODE_PACKED;
    const bool uniform = gstat[2] != 0u;                     // one power-of-two scale for the whole codebook
    // Streaming mode (large D) converts one X chunk per MMA block, which 4 warps cannot sustain, while the
    // epilogue has a whole row of k blocks per tile to drain one accumulator: epilogue warps 12-15 join the
    // converter and warps 8-11 drain all 256 columns.
    constexpr bool conv_extra = MODE == MODE_STREAM;
    constexpr int nconv = conv_extra ? 256 : 128;            // converter threads per CTA
    constexpr int nepi = conv_extra ? 128 : 256;             // epilogue threads per CTA
    const int probe = (acc.dbg == 9 && blockIdx.x == 0) ? 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) { tc::mbar_init(afull_bar(s), 1); tc::mbar_init(aready_bar(s), 2 * nconv / 32); tc::mbar_init(aempty_bar(s), 1); }
        for (int s = 0; s < MAXB; ++s) { tc::mbar_init(bfull_bar(s), 1); tc::mbar_init(bempty_bar(s), 1); }
        for (int a = 0; a < NACC; ++a) { tc::mbar_init(tfull_bar(a), 1); tc::mbar_init(tempty_bar(a), 2 * nepi / 32); }
        for (int b = 0; b < 2; ++b) { tc::mbar_init(bfullq_bar(b), 4); tc::mbar_init(bemptyq_bar(b), 4); }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&map_x); tc::tma_prefetch_desc(&map_whi); tc::tma_prefetch_desc(&map_wlo);
        if (FOLD) tc::tma_prefetch_desc(&map_fold);
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32((const void *)tmem_slot), 512);
    tc::tc_fence_before();
    cluster_sync();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== B producer: this CTA's half of W'hi / W'lo per (row tile, neuron tile, k block) ====
        // (the whole warp runs the loop, one elected lane issues: addresses stay in uniform registers)
        {
            uint32_t it = 0;
            for (int pt = pair; pt < num_pair_tiles; pt += num_pairs)
                for (int nt = 0; nt < num_n_tiles; ++nt)
                    for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                        const int s = it % NBR; const uint32_t ph = (it / NBR) & 1;
                        tc::mbar_wait(bempty_bar(s), ph ^ 1);
                        const uint32_t st = b_base + s * BSLOT;
                        if (tc::elect_one()) {
                            // bytes of both CTAs: hi (+ lo unless packed) halves (+ the fold images)
                            if (leader) tc::mbar_expect_tx(bfull_bar(s), (packed ? 2 : 4) * TBNH * 128 + (FOLD ? 2 * FOLD_TILE_BYTES : 0));
                            const int nrow = nt * TBN + (int)rank * ((nt == num_n_tiles - 1 ? n_last : TBN) / 2);
                            tma_load_2d_2sm(st, &map_whi, kb * BK, nrow, bfull_bar(s));
                            if (!packed) tma_load_2d_2sm(st + HALF_SLOT, &map_wlo, kb * BK, nrow, bfull_bar(s));
                            if (FOLD)      // this CTA's 4 KB fold image: 16 rows of 64 floats of the (k_pad / 8, 64) view, no swizzle
                                tma_load_2d_2sm(bf_base + s * FOLD_TILE_BYTES, &map_fold, 0, (nt * 2 + (int)rank) * 16, bfull_bar(s));
                        }
                        __syncwarp();
                    }
        }
    } else if (warp == APROD_WARP) {
        // ===================== A producer: raw fp32 X chunks [128 rows x 64 features] ==========================
        {
            uint32_t ia = 0;
            for (int pt = pair; pt < num_pair_tiles; pt += num_pairs) {
                const int reps = resident ? 1 : num_n_tiles;
                for (int rep = 0; rep < reps; ++rep)
                    for (int kb = 0; kb < num_k_blocks; ++kb, ++ia) {
                        const int s = ia % NAR; const uint32_t ph = (ia / NAR) & 1;
                        tc::mbar_wait(aempty_bar(s), ph ^ 1);
                        const uint32_t st = a_base + s * SLOT_BYTES;
                        const int row0 = pt * (2 * BM) + (int)rank * BM;
                        if (tc::elect_one()) {
                            tc::mbar_expect_tx(afull_bar(s), packed ? HALF_SLOT : SLOT_BYTES);
                            tc::tma_load_2d(st, &map_x, kb * BK, row0, afull_bar(s));
                            if (!packed) tc::tma_load_2d(st + HALF_SLOT, &map_x, kb * BK + 32, row0, afull_bar(s));
                        }
                        __syncwarp();
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) ==========================================================
        if (leader) {
            uint32_t it = 0, acc_it = 0, tile_it = 0;
            for (int pt = pair; pt < num_pair_tiles; pt += num_pairs, ++tile_it)
                for (int nt = 0; nt < num_n_tiles; ++nt, ++acc_it) {
                    const int a = acc_it % NACC; const uint32_t aph = (acc_it / NACC) & 1;
                    tc::dbg_stamp(probe, 0, acc_it);                 // MMA: starts waiting for the tile's inputs
                    const uint32_t tmem_d = tmem_base + (uint32_t)(a * TBN);
                    const uint32_t idesc = (FOLD || nt != num_n_tiles - 1) ? kIdescF16 : idesc_f16(n_last);
                    for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                        const uint32_t ia = resident ? (tile_it * (uint32_t)num_k_blocks + kb) : it;
                        const int sa = ia % NAR; const uint32_t pha = (ia / NAR) & 1;
                        const int sb = it % NBR; const uint32_t phb = (it / NBR) & 1;
                        mbar_wait_cluster(aready_bar(sa), pha);       // hi/lo tiles of both CTAs written
                        mbar_wait_cluster(bfull_bar(sb), phb);        // W' halves of both CTAs landed
                        if (kb == 0) {                                // operands first, then the accumulator
                            mbar_wait_cluster(tempty_bar(a), aph ^ 1);
                            tc::dbg_stamp(probe, 1, acc_it);         // MMA: accumulator free
                        }
                        tc::tc_fence_after();
                        if (kb == 0) tc::dbg_stamp(probe, 2, acc_it);    // MMA: operands of the first k block ready
                        const uint32_t sta = a_base + sa * SLOT_BYTES, stb = b_base + sb * BSLOT;
                        const uint64_t a_hi = tc::make_smem_desc(sta), a_lo = tc::make_smem_desc(sta + HALF_SLOT);
                        const uint64_t b_hi = tc::make_smem_desc(stb), b_lo = tc::make_smem_desc(stb + HALF_SLOT);
                        const int dl = acc.d - kb * BK;          // real features in this block; the rest is zero fill
                        const int kk_n = dl >= BK ? BK / UMMA_K : (dl + UMMA_K - 1) / UMMA_K;
                        if (tc::elect_one()) {
                            if (packed) {
                                // one tile, three K steps: hi*hi, lo*hi, hi*lo (operands packed along K)
#pragma unroll
                                for (int kk = 0; kk < 3; ++kk) {
                                    const uint64_t off = (uint64_t)((kk * UMMA_K * 2) >> 4);
                                    umma_f16_2sm(tmem_d, a_hi + off, b_hi + off, idesc, kk != 0);
                                }
                            } else {
#pragma unroll
                                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                                    if (kk >= kk_n) break;
                                    const uint64_t off = (uint64_t)((kk * UMMA_K * 2) >> 4);    // +32 B along K
                                    umma_f16_2sm(tmem_d, a_lo + off, b_hi + off, idesc, (kb | kk) != 0);
                                    umma_f16_2sm(tmem_d, a_hi + off, b_lo + off, idesc, 1);
                                    umma_f16_2sm(tmem_d, a_hi + off, b_hi + off, idesc, 1);
                                }
                            }
                            if (FOLD)      // acc += 2^a_r * (bias_k 2^b_k): one kind::tf32 step, K = 8 (three pieces + zeros)
                                umma_tf32_2sm(tmem_d, make_smem_desc_nosw(af_base + sa * FOLD_TILE_BYTES, 128, 256),
                                              make_smem_desc_nosw(bf_base + sb * FOLD_TILE_BYTES, 128, 256), kIdesc2, 1);
                            umma_commit_2sm(bempty_bar(sb));
                            if (!resident || nt == num_n_tiles - 1) umma_commit_2sm(aempty_bar(sa));
                            if (kb == num_k_blocks - 1) {
                                umma_commit_2sm(tfull_bar(a));
                                tc::dbg_stamp(probe, 3, acc_it);     // MMA: all MMAs of the tile issued + committed
                            }
                        }
                        __syncwarp();
                    }
                }
        }
    } else if ((warp >= CONV_WARP0 && warp < EPI_WARP0) || (conv_extra && warp >= EPI_WARP0 + 4 && warp < SCAT_WARP0)) {
        // ===================== converter: raw fp32 -> scaled fp16 hi | lo, in place ==============================
        const int t = warp < EPI_WARP0 ? (int)threadIdx.x - CONV_WARP0 * 32
                                       : 128 + (int)threadIdx.x - (EPI_WARP0 + 4) * 32;     // 0..nconv-1
        uint32_t ia = 0;
        for (int pt = pair; pt < num_pair_tiles; pt += num_pairs) {
            const int reps = resident ? 1 : num_n_tiles;
            const int64_t row0 = (int64_t)pt * (2 * BM) + (int64_t)rank * BM;
            for (int rep = 0; rep < reps; ++rep)
                for (int kb = 0; kb < num_k_blocks; ++kb, ++ia) {
                    const int s = ia % NAR; const uint32_t ph = (ia / NAR) & 1;
                    float rs_t = 1.f;                         // scale of THIS thread's row (fold operand, packed tile)
                    if ((FOLD || packed) && row0 + t < n) rs_t = __ldg(xscale + row0 + t);
                    tc::mbar_wait(afull_bar(s), ph);
                    uint8_t *slot = smem + s * SLOT_BYTES;
                    if (packed)          convert_item_packed(slot, t, rs_t);
                    else if (conv_extra) convert_item<256>(slot, t, row0, n, xscale);
                    else                 convert_item<128>(slot, t, row0, n, xscale);
                    if (FOLD) {
                        // A_f row t: K chunk 0 = (u, u, u, 0) with u = 2^a_r, K chunk 1 = 0 (no-swizzle K-major core matrices)
                        uint8_t *af = smem + (af_base - smem_base) + s * FOLD_TILE_BYTES + (t >> 3) * 256 + (t & 7) * 16;
                        *reinterpret_cast<float4 *>(af) = make_float4(rs_t, rs_t, rs_t, 0.f);
                        *reinterpret_cast<float4 *>(af + 128) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    tc::fence_proxy_async();
                    __syncwarp();                                           // one arrival per warp (see bmu_tc2.cuh)
                    if (lane == 0) mbar_arrive_cluster(map_to_cta(aready_bar(s), 0));
                }
        }
    } else if (warp >= EPI_WARP0 && warp < (conv_extra ? EPI_WARP0 + 4 : SCAT_WARP0)) {
        // ===================== epilogue: TMEM -> registers -> running argmin ====================================
        // resident mode: 8 warps, two per TMEM lane quarter, 128 columns each (h = column half);
        // streaming mode: 4 warps (8-11), all 256 columns each
        const int q = warp & 3;
        const int h = conv_extra ? 0 : (warp - EPI_WARP0) >> 2;
        constexpr int ncols = conv_extra ? TBN : TBN / 2;
        const bool ld_lane = lane < ncols / 4 || conv_extra;       // lanes that stage this warp's bias slice
        const int row_in_tile = q * 32 + lane;
        const float winv0 = __ldg(wsinv);                    // 2^-b of the uniform codebook scale
        // folded bias + uniform codebook scale: the raw accumulator is already a row-constant multiple of the score,
        // nothing is staged; otherwise this warp's private shared-memory slice holds ncols bias values, then ncols
        // inverse scales
        const bool staged = !(FOLD && uniform);
        float *wb = epi_stage + (warp - EPI_WARP0) * (conv_extra ? 512 : 256);
        float *wsv = wb + ncols;
        uint32_t acc_it = 0, tile_it = 0;
        // row scale of the FIRST tile; later ones are fetched one tile ahead (no exposed global latency)
        float rs_next = 1.f;
        {
            const int64_t row = (int64_t)pair * (2 * BM) + (int64_t)rank * BM + row_in_tile;
            if (pair < num_pair_tiles && row < n) rs_next = __ldg(xscale + row);
        }
        float4 nb = make_float4(0.f, 0.f, 0.f, 0.f), ns = nb;
        if (staged && ld_lane) {
            nb = __ldg(reinterpret_cast<const float4 *>(bias + h * (TBN / 2)) + lane);
            ns = __ldg(reinterpret_cast<const float4 *>(wsinv + h * (TBN / 2)) + lane);
        }
        float4 nb2 = nb, ns2 = ns;                               // columns 128..255 (streaming mode only)
        if (conv_extra) {
            nb2 = __ldg(reinterpret_cast<const float4 *>(bias + 128) + lane);
            ns2 = __ldg(reinterpret_cast<const float4 *>(wsinv + 128) + lane);
        }
        for (int pt = pair; pt < num_pair_tiles; pt += num_pairs, ++tile_it) {
            const int64_t row = (int64_t)pt * (2 * BM) + (int64_t)rank * BM + row_in_tile;
            const float rs = rs_next;
            {
                const int64_t nrow = row + (int64_t)num_pairs * (2 * BM);
                rs_next = (pt + num_pairs < num_pair_tiles && nrow < n) ? __ldg(xscale + nrow) : 1.f;
            }
            // uniform codebook scale 2^b: argmin of acc * 2^-b + bias * rs == argmin of acc + bias * (rs * 2^b)
            const float rsg = rs / winv0;
            RunMinScaled rm; rm.reset();
            for (int nt = 0; nt < num_n_tiles; ++nt, ++acc_it) {
                const int a = acc_it % NACC; const uint32_t aph = (acc_it / NACC) & 1;
                const int col0 = nt * TBN + h * (TBN / 2);
                if (staged) {
                    __syncwarp();
                    if (ld_lane) {
                        reinterpret_cast<float4 *>(wb)[lane] = nb;
                        reinterpret_cast<float4 *>(wsv)[lane] = ns;
                    }
                    if (conv_extra) {
                        reinterpret_cast<float4 *>(wb + 128)[lane] = nb2;
                        reinterpret_cast<float4 *>(wsv + 128)[lane] = ns2;
                    }
                    __syncwarp();
                    {   // prefetch the next neuron tile's slice (wraps to tile 0 for the next row tile)
                        const int nn = (nt + 1 < num_n_tiles ? nt + 1 : 0) * TBN + h * (TBN / 2);
                        if (ld_lane) {
                            nb = __ldg(reinterpret_cast<const float4 *>(bias + nn) + lane);
                            ns = __ldg(reinterpret_cast<const float4 *>(wsinv + nn) + lane);
                        }
                        if (conv_extra) {
                            nb2 = __ldg(reinterpret_cast<const float4 *>(bias + nn + 128) + lane);
                            ns2 = __ldg(reinterpret_cast<const float4 *>(wsinv + nn + 128) + lane);
                        }
                    }
                }
                if (warp == EPI_WARP0 && lane == 0) tc::dbg_stamp(probe, 4, acc_it);   // EPI: starts waiting for the tile
                tc::mbar_wait(tfull_bar(a), aph);
                tc::tc_fence_after();
                if (warp == EPI_WARP0 && lane == 0) tc::dbg_stamp(probe, 5, acc_it);   // EPI: tile complete in TMEM
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TBN + h * (TBN / 2));
                auto drain = [&](const uint32_t (&v)[32], int c) {
                    if (FOLD) {
                        if (uniform) rm.chunk_nobias(v, col0 + c * 32);
                        else         rm.chunk_scaled_nobias(v, wsv + c * 32, col0 + c * 32);
                    } else {
                        if (uniform) rm.chunk_uniform(v, wb + c * 32, rsg, col0 + c * 32);
                        else         rm.chunk(v, wb + c * 32, wsv + c * 32, rs, col0 + c * 32);
                    }
                };
                // (measured in round 2: a second tcgen05.ld buffer in flight, 64-column loads and 8 instead of 4 running
                // minima all made this loop SLOWER or no faster.  The drain of a 128 x 256 tile takes ~1400 cycles here;
                // tools/micro/tmem_bw.cu in isolation: 450 - 530 for the TMEM reads alone, 1106 with the compare and the two
                // predicated writes per score on 8 warps, 951 on 16: instruction issue, with 2 warps per scheduler.)
                // Chunks that hold only padding neurons (the last tile of a map whose size is not a multiple of 256) are
                // skipped: their scores are +inf.  40 x 40 neurons: 6.25 instead of 7 tiles to drain per row tile.
                int nch = ncols / 32;
                {
                    const int left = acc.k - col0;
                    nch = left <= 0 ? 0 : (left + 31) >> 5 < nch ? (left + 31) >> 5 : nch;
                }
#pragma unroll 1
                for (int c = 0; c < nch; ++c) {
                    uint32_t v[32];
                    tc::tmem_ld32(taddr + c * 32, v);
                    tc::tmem_ld_wait_dep(v);
                    drain(v, c);
                }
                tc::tc_fence_before();
                if (warp == EPI_WARP0 && lane == 0) tc::dbg_stamp(probe, 6, acc_it);   // EPI: this warp drained its half
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty_bar(a), 0));
            }
            float best; int bidx;
            rm.result(best, bidx);
            const int mb = (tile_it & 1) * BM;
            if (!conv_extra) {       // two warps per row: merge the column halves through shared memory
                if (h == 1) { mrg_v[mb + row_in_tile] = best; mrg_i[mb + row_in_tile] = bidx; }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if (h == 0) {
                if (!conv_extra) argmin_merge(best, bidx, mrg_v[mb + row_in_tile], mrg_i[mb + row_in_tile]);
                if (row < n) {
                    if (bmu_out) bmu_out[row] = bidx;
                    // undo the scaling (exact powers of two): the score x . w' + bias
                    if (best_out) best_out[row] = FOLD ? (uniform ? best * winv0 / rs : best / rs) : (uniform ? best / rsg : best / rs);
                }
                if (fused) {
                    const int b = tile_it & 1; const uint32_t bph = (tile_it >> 1) & 1;
                    tc::mbar_wait(bemptyq_bar(b), bph ^ 1);
                    bmu_s[b * BM + row_in_tile] = (row < n) ? bidx : -1;
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(bfullq_bar(b));
                }
            }
        }
    } else if (warp >= SCAT_WARP0) {
        // ===================== scatter: exact per-BMU sums (common.cuh: scatter_rows_exact) =====================
        if (fused) {
            const int wq = warp - SCAT_WARP0;
            long long *stage = reinterpret_cast<long long *>(scat_stage) + wq * (tc::SCAT_NBUF * ACC_PIECE);
            uint32_t tile_it = 0, bulk_it = 0;
            const FusedAcc racc = acc.for_cta(blockIdx.x >> 1);          // this pair's replica of the accumulator
            for (int pt = pair; pt < num_pair_tiles; pt += num_pairs, ++tile_it) {
                const int b = tile_it & 1; const uint32_t bph = (tile_it >> 1) & 1;
                tc::mbar_wait_relaxed(bfullq_bar(b), bph, 400);
                const int64_t row0 = (int64_t)pt * (2 * BM) + (int64_t)rank * BM;
                scatter_rows_exact<tc::SCAT_NBUF>(racc, bmu_s + b * BM, row0, BM, wq, 4, lane, stage, bulk_it);
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(bemptyq_bar(b));
            }
            bulk_wait_all();          // every bulk reduction of this thread has been performed
        }
    }

    __syncwarp();
    tc::tc_fence_before();
    cluster_sync();
    if (warp == 1) { tc::tc_fence_after(); tmem_dealloc_2sm(tmem_base, 512); }
}

inline int make_map_2d_f16(CUtensorMap *m, const void *base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                           uint32_t box_inner, uint32_t box_outer) {
    tc::EncodeTiledFn enc = tc::get_encode_fn();
    SOM_REQUIRE(enc != nullptr, SOM_E_NODEVICE, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SOM_REQUIRE(r == CUDA_SUCCESS, SOM_E_SHAPE, "cuTensorMapEncodeTiled(f16) failed with CUresult %d", (int)r);
    return 0;
}

inline int launch_bmu_tc3(const float *X, int64_t n, int d, int64_t ldx, const float *xscale, int k, const WsLayout &L,
                          uint8_t *ws, int32_t *bmu, float *best, const AccTarget &T, int sm_count, cudaStream_t st) {
    SOM_REQUIRE(L.k_pad < (1 << 24), SOM_E_SHAPE,
                "tensor-core BMU kernels track the winning neuron as an exact fp32 integer: at most 2^24 neurons (k=%d)", k);
    SOM_REQUIRE(tc::shape_ok(X, n, d, ldx), SOM_E_SHAPE,
                "tensor-core BMU kernel needs ldx %% 4 == 0 and a 16-byte aligned X for TMA (d=%d ldx=%lld)", d, (long long)ldx);
    SOM_REQUIRE(xscale != nullptr, SOM_E_BADARG, "the fp16-split kernel needs the per-row scales (som_b200_prepare_samples)");
    const int num_k_blocks = L.d_pad64 / BK;
    const bool packed = d <= 16;
    const bool resident = num_k_blocks <= RESIDENT_MAX_KB;
    static const int no_fold = tc::env_int("SOM_B200_NO_FOLD");        // experiments builds: bias in the epilogue
    const bool fold = packed || (resident && num_k_blocks == 1 && !no_fold);
    CUtensorMap mx, mhi, mlo, mfold;
    int rc;
    if ((rc = tc::make_map_2d(&mx, X, (uint64_t)d, (uint64_t)n, (uint64_t)ldx * 4, 32, BM))) return rc;
    // fold operand image: 4 KB per 128 neurons, viewed as rows of 64 floats; one box = 16 rows = one image, no swizzle
    if ((rc = tc::make_map_2d(&mfold, ws + L.wfold_off, 64, (uint64_t)L.k_pad / 8, 256, 64, 16, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    if ((rc = make_map_2d_f16(&mhi, ws + L.w16hi_off, (uint64_t)L.d_pad64, (uint64_t)L.k_pad, (uint64_t)L.d_pad64 * 2, BK, TBNH))) return rc;
    if ((rc = make_map_2d_f16(&mlo, ws + L.w16lo_off, (uint64_t)L.d_pad64, (uint64_t)L.k_pad, (uint64_t)L.d_pad64 * 2, BK, TBNH))) return rc;
    static bool attr_set[64] = {};
    if (tc::first_launch_on_device(attr_set)) {
        SOM_CUDA(cudaFuncSetAttribute(bmu_tc3_kernel<MODE_RESIDENT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SOM_CUDA(cudaFuncSetAttribute(bmu_tc3_kernel<MODE_RESIDENT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SOM_CUDA(cudaFuncSetAttribute(bmu_tc3_kernel<MODE_STREAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SOM_CUDA(cudaFuncSetAttribute(bmu_tc3_kernel<MODE_PACKED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    const int num_pair_tiles = (int)ceil_div(n, 2 * BM);
    const int num_n_tiles = L.k_pad / TBN;
    int pairs = sm_count / 2;
    if (pairs > num_pair_tiles) pairs = num_pair_tiles;
    if (pairs < 1) pairs = 1;
    FusedAcc acc;
    acc.X = X; acc.ldx = ldx; acc.d = d; acc.k = k; acc.S = T.S; acc.cnt = T.cnt; acc.qscale = T.qscale;
    acc.lds = acc_ld(d);
    acc.reps = T.reps; acc.rep_words = T.rep_words;
    acc.vec = (d % 4 == 0) ? 1 : 0;              // X rows are 16-byte aligned here (tc::shape_ok)
    static const int dbg = tc::env_int("SOM_B200_DBG");
    acc.dbg = dbg;
    const float *bias_p = reinterpret_cast<const float *>(ws + L.bias_off);
    const float *wsinv_p = reinterpret_cast<const float *>(ws + L.wsinv_off);
    const unsigned int *gstat_p = reinterpret_cast<const unsigned int *>(ws + L.gstat_off);
    const dim3 grid(2 * pairs), block(NUM_THREADS);
    cudaError_t e;
#define SOM_LAUNCH_TC3(M, F)                                                                                              \
    e = launch_pdl(bmu_tc3_kernel<M, F>, grid, block, SMEM_BYTES, st, mx, mhi, mlo, mfold, bias_p, wsinv_p, gstat_p, xscale, n, \
                   num_pair_tiles, num_n_tiles, num_k_blocks, bmu, best, acc)
    if (packed)                SOM_LAUNCH_TC3(MODE_PACKED, true);
    else if (resident && fold) SOM_LAUNCH_TC3(MODE_RESIDENT, true);
    else if (resident)         SOM_LAUNCH_TC3(MODE_RESIDENT, false);
    else                       SOM_LAUNCH_TC3(MODE_STREAM, false);
#undef SOM_LAUNCH_TC3
    return check_cuda(e, "bmu_tc3_kernel launch");
}

}  // namespace tc3
}  // namespace somb200
