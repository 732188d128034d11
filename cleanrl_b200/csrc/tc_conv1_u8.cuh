// conv1 straight from the uint8 rollout (no 16-bit copy of the frames anywhere).
//
// The rollout keeps every frame ONCE as uint8 space-to-depth(4) pixels, 28 224 B per frame (the algorithmic
// minimum; reference: a 14.8 GB fp32 buffer, ppo_atari_envpool.py:203):
//     frames    u8 [img][441 grid rows][64 ch]   row-major    -> forward  (K-major operand rows)
//     frames_t  u8 [img][64 ch][448 grid rows]   channel-major -> weight gradient (lanes = channels, K = rows)
//
// Forward  tc_conv1_i8: integer tensor cores (tcgen05.mma kind::i8, u8 x s8 -> s32, accumulators in TMEM).  The
//   pixels are EXACT (0..255 are integers); the fp32 master weights are split per output channel into two signed
//   8-bit limbs  w ~= s_co * (l1 / 2^7 + l2 / 2^14)  (|error| <= s_co * 2^-15, i.e. 15 bits relative to the row
//   maximum -- tighter than the 8 bits of a bf16 weight), the two limbs are 2 x 32 = 64 GEMM columns, and the
//   epilogue recombines  y = acc1 * s/2^7/255 + acc2 * s/2^14/255 + bias.  Integer accumulation is exact, so the
//   result does not depend on the order of the 256-term dot products.  A 1-byte operand also halves the
//   shared-memory operand traffic that bounded the bf16 kernel (N = 32 is too narrow to amortise the 128-row A tile).
//   Same window scheme as tc_conv_win: one TMA box of 128 + 22 rows per tile, the four 2x2 taps are descriptors
//   shifted by whole 64-byte rows of a SWIZZLE_64B image.
//
// Weight gradient  tc_conv1_wgrad_u8: dW^T[(tap, c), co] = sum_r X[r + shift_tap, c] * dY[r, co].  The pixels go
//   uint8 (shared memory, channel-major TMA box) -> fp16 pairs in REGISTERS (one PRMT per two pixels builds
//   1024 + x, one HSUB2 removes the 1024) -> tensor memory (tcgen05.st), and are consumed as the A operand
//   straight from TMEM (tcgen05.mma with A in TMEM); dY rows are a SWIZZLE_64B TMA box used as an MN-major B
//   operand with N = 32.  No 16-bit image of the frames ever exists in shared or global memory.
#pragma once
#include "tc_base.cuh"
#include <cuda_fp16.h>
#include "tc_conv_win.cuh"

namespace b200rl {
using namespace tc;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory (lane = row m; 32-bit column j holds K elements 2j, 2j+1), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr uint64_t kDescSwizzle64 = 4ull << 61;      // LayoutType::SWIZZLE_64B
// K-major SWIZZLE_64B operand: rows 64 B apart, 8-row atoms 512 B apart (SBO)
__device__ __forceinline__ uint64_t desc_kmajor_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | kDescVersion | kDescSwizzle64;
}
// MN-major SWIZZLE_64B operand with ONE 32-element (64-byte) MN atom: rows = K index, 8-row K groups 512 B apart
__device__ __forceinline__ uint64_t desc_mnmajor_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | kDescVersion | kDescSwizzle64;
}
// byte offset of 16-byte chunk c (0..3) of row r in a SWIZZLE_64B image (address bits [4,6) ^= bits [7,9))
__device__ __forceinline__ uint32_t img64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }
// u8 x s8 -> s32
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// A = fp16, B = bf16 or fp16 -> fp32;  b_mn_major = 1: B rows are the K index
__host__ __device__ constexpr uint32_t make_idesc_f16ts(int M, int N, int b_bf16, int b_mn_major) {
    return (1u << 4) | (0u << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------ frame conversion (once per env step)
// uint8 frames [n,4,84,84] (NCHW, as the env delivers them) -> row-major u8 [n,441,64] and channel-major u8 [n,64,448]
// space-to-depth(4) pixels: channel = c*16 + sy*4 + sx of source pixel (4Y+sy, 4X+sx), grid row = Y*21 + X.
// One block per (frame, colour plane c): the 84x84 plane is staged in shared memory (coalesced 16-byte reads).
__global__ void __launch_bounds__(256) tc_frames_to_s2d_u8(const uint8_t* __restrict__ obs, const int64_t* __restrict__ rows, int64_t n,
                                                           uint8_t* __restrict__ out_rm, uint8_t* __restrict__ out_cm) {
    __shared__ __align__(16) uint8_t plane[7056];
    const int64_t i = blockIdx.x >> 2;
    const int c = blockIdx.x & 3;
    const int64_t img = rows ? rows[i] : i;
    const int4* src = reinterpret_cast<const int4*>(obs + img * 28224 + c * 7056);
    for (int t = threadIdx.x; t < 441; t += 256) reinterpret_cast<int4*>(plane)[t] = __ldg(src + t);
    __syncthreads();
    // row-major: 16 bytes (4 rows of 4 pixels) per grid position
    for (int pos = threadIdx.x; pos < 441; pos += 256) {
        const int Y = pos / 21, X = pos - Y * 21;
        const uint8_t* p = plane + (Y * 4) * 84 + X * 4;
        int4 v;
        v.x = *reinterpret_cast<const int*>(p); v.y = *reinterpret_cast<const int*>(p + 84);
        v.z = *reinterpret_cast<const int*>(p + 168); v.w = *reinterpret_cast<const int*>(p + 252);
        *reinterpret_cast<int4*>(out_rm + (i * 441 + pos) * 64 + c * 16) = v;
    }
    // channel-major: channel (sy, sx) of this plane, 4 consecutive grid rows per 32-bit store (448-byte rows, zero tail)
    for (int t = threadIdx.x; t < 16 * 112; t += 256) {
        const int ch = t / 112, q = t - ch * 112;
        const int sy = ch >> 2, sx = ch & 3;
        uint32_t w = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int pos = q * 4 + e;
            if (pos < 441) {
                const int Y = pos / 21, X = pos - Y * 21;
                w |= (uint32_t)plane[(Y * 4 + sy) * 84 + X * 4 + sx] << (8 * e);
            }
        }
        *reinterpret_cast<uint32_t*>(out_cm + (i * 64 + c * 16 + ch) * 448 + q * 4) = w;
    }
}

// ------------------------------------------------------------------------------------ conv1 weight limbs
// w[co][c][ky][kx] (fp32) -> s8 limbs L[(limb, co)][tap (a,b)][c*16 + sy*4 + sx] with ky = 4a+sy, kx = 4b+sx, and the
// per-column output scales sc[limb*32 + co] = s_co / 2^(7*(limb+1)) / 255 (the /255 of ppo_atari_envpool.py:144).
__global__ void __launch_bounds__(256) tc_pack_conv1_i8(const float* __restrict__ w, int8_t* __restrict__ limbs, float* __restrict__ sc) {
    __shared__ float red[32];
    const int co = blockIdx.x, idx = threadIdx.x;                  // idx = (c, ky, kx)
    const float v = w[co * 256 + idx];
    float m = fabsf(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((idx & 31) == 0) red[idx >> 5] = m;
    __syncthreads();
    float mx = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) mx = fmaxf(mx, red[k]);
    const float s = mx > 0.f ? mx * (128.0f / 127.0f) : 1.0f;
    const float u = v / s * 128.0f;                                // |u| <= 127
    float l1 = rintf(u);
    l1 = fminf(fmaxf(l1, -127.f), 127.f);
    float l2 = rintf((u - l1) * 128.0f);
    l2 = fminf(fmaxf(l2, -127.f), 127.f);
    const int kx = idx & 7, ky = (idx >> 3) & 7, c = idx >> 6;
    const int a = ky >> 2, sy = ky & 3, b = kx >> 2, sx = kx & 3;
    const int k = (a * 2 + b) * 64 + c * 16 + sy * 4 + sx;
    limbs[co * 256 + k] = (int8_t)l1;
    limbs[(32 + co) * 256 + k] = (int8_t)l2;
    if (idx == 0) {
        sc[co] = s / 128.0f / 255.0f;
        sc[32 + co] = s / 16384.0f / 255.0f;
    }
}

// ------------------------------------------------------------------------------------ conv1 forward (kind::i8)
struct Conv1U8Params {
    const int64_t* rows;     // optional image gather (minibatch rows of the rollout)
    int n;                   // images in this launch
    int64_t n_images;        // images addressable through `rows`
    const int8_t* limbs;     // [64][256] s8 (tc_pack_conv1_i8)
    const float* sc;         // [64] column scales
    const float* bias;       // [32]
    bf16* out;               // act1 as 2x2 cells [n,100,128]
    uint32_t* mask_out;      // act1 > 0 bits: [n,100 cells] x 4 words
};

template <int STAGES>
__global__ void __launch_bounds__(kConvWinThreads, 1) tc_conv1_i8(const __grid_constant__ CUtensorMap tmA, const Conv1U8Params p, int total_tiles) {
    constexpr int BN = 64, WR = 152, NTAPS = 4;
    constexpr int STAGE_BYTES = WR * 64;            // 9728: a multiple of 512, so every stage is atom-aligned
    constexpr int B_CHUNK = BN * 64;                // one tap of the limb image: 64 rows x 64 B
    constexpr uint32_t TMEM_COLS = 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_sc[64], s_bias[32];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sW = smem;                             // 4 taps x 4096 B
    uint8_t* sRing = smem + NTAPS * B_CHUNK;        // 16384: 1024-aligned

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
    for (int idx = tid; idx < NTAPS * BN * 4; idx += blockDim.x) {          // 16-byte chunks of the limb image
        const int c16 = idx & 3;
        const int t = (idx >> 2) & 3;
        const int r = idx >> 4;
        *reinterpret_cast<int4*>(sW + t * B_CHUNK + img64_off(r, c16)) = ldg16(p.limbs + r * 256 + t * 64 + c16 * 16);
    }
    if (tid < 64) s_sc[tid] = p.sc[tid];
    if (tid < 32) s_bias[tid] = p.bias[tid];
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = tmem_base_smem;
    const int tile_begin = (int)(((int64_t)total_tiles * blockIdx.x) / gridDim.x);
    const int tile_end = (int)(((int64_t)total_tiles * (blockIdx.x + 1)) / gridDim.x);

    if (warp == 0) {
        // ======================= TMA producer: one [152 rows x 64 B] box of one image per tile (4 tiles per image);
        // the image coordinate is the minibatch gather, fetched one tile ahead
        if (tid == 0) {
            uint32_t q = 0;
            int z_next = 0;
            if (tile_begin < tile_end) {
                const int img = tile_begin >> 2;
                z_next = p.rows ? (int)__ldg(p.rows + img) : img;
            }
            for (int tile = tile_begin; tile < tile_end; ++tile, ++q) {
                const uint32_t s = q % STAGES;
                const int z = z_next;
                if (tile + 1 < tile_end) {
                    const int img = (tile + 1) >> 2;
                    z_next = p.rows ? (int)__ldg(p.rows + img) : img;
                }
                if (q >= (uint32_t)STAGES) mbar_wait(&empty_bar[s], ((q / STAGES) - 1) & 1);
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)STAGE_BYTES);
                tma_load_3d(smem_u32(sRing + (size_t)s * STAGE_BYTES), &tmA, 0, (tile & 3) * 128, z, &full_bar[s]);
            }
        }
    } else if (warp == 1 || warp == 10) {
        // ======================= MMA issuers (warp-uniform loop, one elected lane each; issuer ih owns the tiles of parity ih
        // = accumulator buffer ih, as in tc_conv_win): per tile 4 taps x 2 K-steps of 32 bytes
        const uint32_t ih = warp == 1 ? 0u : 1u;
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_i8(128, BN);
        const uint64_t desc_hi = desc_kmajor_sw64(0) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo_flags = (uint32_t)(desc_kmajor_sw64(0) & 0xFFFFFFFFull);
        const uint32_t w_lo = ((smem_u32(sW) & 0x3FFFFu) >> 4) | desc_lo_flags;
        constexpr int shift[4] = {0, 1, 21, 22};
        for (uint32_t q = ih; (int)q < tile_end - tile_begin; q += 2) {
            const uint32_t acc = ih, s = q % STAGES;
            if (q >= 2) mbar_wait(&tempty_bar[acc], ((q >> 1) - 1) & 1);
            mbar_wait(&full_bar[s], (q / STAGES) & 1);
            tc_fence_after_sync();
            if (leader) {
                const uint32_t d_addr = tmem_d + acc * BN;
                const uint32_t win_lo = ((smem_u32(sRing + (size_t)s * STAGE_BYTES) & 0x3FFFFu) >> 4) | desc_lo_flags;
#pragma unroll
                for (int t = 0; t < NTAPS; ++t) {
                    const uint32_t a_lo = win_lo + (uint32_t)shift[t] * 4u;          // whole 64-byte rows
                    const uint32_t b_lo = w_lo + (uint32_t)((t * B_CHUNK) >> 4);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
                        umma_i8(d_addr, desc_hi | (uint64_t)(a_lo + 2 * kk), desc_hi | (uint64_t)(b_lo + 2 * kk), idesc,
                                (t | kk) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
                umma_commit(&tfull_bar[acc]);
            }
            __syncwarp();
        }
    } else {
        // ======================= epilogue: warps 2-9 = two groups of four (one warp per TMEM lane quadrant); group h owns
        // accumulator buffer h = every other tile.  Row -> (image, Y, X) -> 2x2-cell offset once per tile and thread.
        const int ew = warp & 3;
        const int h = (warp - 2) >> 2;
        const int lrow = ew * 32 + (tid & 31);
        const uint32_t lane_addr = tmem_d + h * BN + ((uint32_t)(ew * 32) << 16);
        uint32_t k = 0;
        for (int tile = tile_begin + h; tile < tile_end; tile += 2, ++k) {
            const int i = tile >> 2;
            const int rem = ((tile & 3) << 7) + lrow;
            const int Y = (rem * 3121) >> 16, X = rem - Y * 21;          // rem / 21 for rem < 512
            const bool valid = rem < 441 && Y < 20 && X < 20;
            const int64_t cell = ((int64_t)i * 10 + (Y >> 1)) * 10 + (X >> 1);
            const int cls = (Y & 1) * 2 + (X & 1);
            mbar_wait(&tfull_bar[h], k & 1);
            tc_fence_after_sync();
            uint32_t a1[32], a2[32];
            tmem_ld32(lane_addr, a1);
            tmem_ld32(lane_addr + 32, a2);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&tempty_bar[h]);              // accumulator drained
            if (!valid) continue;
            uint32_t bits = 0u;
            int4 w[4];
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                float f0 = fmaf((float)(int)a1[e], s_sc[e], fmaf((float)(int)a2[e], s_sc[32 + e], s_bias[e]));
                float f1 = fmaf((float)(int)a1[e + 1], s_sc[e + 1], fmaf((float)(int)a2[e + 1], s_sc[33 + e], s_bias[e + 1]));
                const bool p0 = f0 > 0.f, p1 = f1 > 0.f;
                bits |= (p0 ? 1u : 0u) << e;
                bits |= (p1 ? 1u : 0u) << (e + 1);
                pk[e >> 1] = pack_bf16x2(p0 ? f0 : 0.f, p1 ? f1 : 0.f);
            }
            p.mask_out[cell * 4 + cls] = bits;
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = make_int4((int)pk[4 * e], (int)pk[4 * e + 1], (int)pk[4 * e + 2], (int)pk[4 * e + 3]);
            int4* dst = reinterpret_cast<int4*>(p.out + cell * 128 + cls * 32);
            dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2]; dst[3] = w[3];
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
}


// ------------------------------------------------------------------------------------ conv1 weight gradient (uint8 frames -> TMEM)
// dW^T[(tap, c), co] = sum_r X[r + shift_tap, c] * dY[r, co] over the grid rows r of a step of 128 rows (4 steps per image).
//   X   : channel-major frames [img][64 ch][448 rows] u8: one TMA box [64 ch][176 rows] per step (row pitch 176 B in shared
//         memory: 16-byte reads of 32 consecutive channels are bank-conflict free).
//   dY  : d(act1) on the 21x21 grid, fp16 [img][441][32] scaled by 2^12 (written by conv2's data gradient with a
//         saturating conversion): one SWIZZLE_64B TMA box of 128 rows per step = MN-major B operand, N = 32.
//   A   : 2 M-tiles of 128 lanes; lane m of tile j = (tap = 2 (m >> 6) + j, channel m & 63); K = the 128 rows of the step
//         = 64 TMEM columns per tile, double-buffered.  Eight convert warps (lane quadrant = warp % 4, K half = warp / 4)
//         read their channel's bytes with 16-byte loads, funnel-shift the two taps' windows (PRMT), expand uint8 -> fp16 with
//         one PRMT per two pixels (bytes (x, 0x64) = fp16 1024 + x) and one HSUB2, and tcgen05.st them into tensor memory.
//   Two issuer warps (one per M-tile) run tcgen05.mma with A in TMEM; four more warps accumulate the bias gradient
//   (column sums of dY) from the staged tiles and drain the accumulators at the end.  Partial tiles go to
//   ws[cta][256][64] / wsb[cta][64] (first 32 columns used) and are folded in fixed order by tc_fold_win.
struct Conv1WgradU8Params {
    const int64_t* rows;       // optional image gather (minibatch rows of the rollout)
    int n;                     // images of the minibatch
    int64_t rows_per_cta;      // multiple of 128 grid rows (M = n * 512)
    float* ws;
    float* wsb;
};
constexpr int kC1WStages = 6;
constexpr int kC1WXBytes = 64 * 176, kC1WYBytes = 128 * 64, kC1WStageBytes = kC1WXBytes + kC1WYBytes;    // 19456 = 19 * 1024
constexpr float kDact1Scale = 4096.0f;

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t hsub2_1024(uint32_t h) {         // (1024 + x) - 1024, exact
    uint32_t d;
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(h), "r"(0x64006400u));
    return d;
}
// 64 pixels (16 words starting at byte offset `b` of W[w0]) -> 32 packed fp16 pairs, K order preserved
template <int B>
__device__ __forceinline__ void u8_window_to_f16(const uint32_t* W, uint32_t (&out)[32]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        uint32_t x = W[j];
        if (B == 1) x = prmt(W[j], W[j + 1], 0x4321u);
        if (B == 2) x = prmt(W[j], W[j + 1], 0x5432u);
        if (B == 3) x = prmt(W[j], W[j + 1], 0x6543u);
        out[2 * j] = hsub2_1024(prmt(x, 0x64646464u, 0x5140u));
        out[2 * j + 1] = hsub2_1024(prmt(x, 0x64646464u, 0x7362u));
    }
}

__global__ void __launch_bounds__(512, 1) tc_conv1_wgrad_u8(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                                                            const Conv1WgradU8Params p) {
    constexpr int STAGES = kC1WStages;
    constexpr uint32_t TMEM_COLS = 512, COL_D = 0, COL_A = 64;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], a_full[2], a_empty[2], done_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float sRed[32 * 32];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 8 + 2 + 4); }
        for (int b = 0; b < 2; ++b) { mbar_init(&a_full[b], 8); mbar_init(&a_empty[b], 2); }
        mbar_init(&done_bar, 2);
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmY);
    }
    if (warp == 3) tmem_alloc(&tmem_base_smem, TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem0 = tmem_base_smem;
    const int64_t M = (int64_t)p.n * 512;
    const int64_t m_begin = (int64_t)blockIdx.x * p.rows_per_cta;
    int64_t m_end = m_begin + p.rows_per_cta;
    if (m_end > M) m_end = M;
    const int nsteps = m_end > m_begin ? (int)((m_end - m_begin) >> 7) : 0;
    const int64_t g0 = m_begin >> 7;                       // first global step (4 steps per image)

    if (warp == 0) {
        // ======================= TMA producer: dY rows (SW64) + the channel-major frame window of the step
        if (lane == 0) {
            int z_next = 0;
            if (nsteps > 0) {
                const int64_t img = g0 >> 2;
                z_next = p.rows ? (int)__ldg(p.rows + img) : (int)img;
            }
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % STAGES;
                const int64_t g = g0 + it;
                const int z = z_next;
                if (it + 1 < nsteps) {
                    const int64_t img1 = (g + 1) >> 2;
                    z_next = p.rows ? (int)__ldg(p.rows + img1) : (int)img1;
                }
                if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
                const uint32_t dst = smem_u32(smem + (size_t)s * kC1WStageBytes);
                const int t_in = (int)(g & 3);
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)kC1WStageBytes);
                tma_load_3d(dst, &tmY, 0, t_in * 128, (int)(g >> 2), &full_bar[s]);
                tma_load_3d(dst + kC1WYBytes, &tmX, t_in * 128, 0, z, &full_bar[s]);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ======================= MMA issuers: M-tile j = warp - 1; 8 K-steps of 16 rows per step, A from tensor memory
        const int j = warp - 1;
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_f16ts(128, 32, 0, 1);
        const uint64_t desc_hi = desc_mnmajor_sw64(0) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo_flags = (uint32_t)(desc_mnmajor_sw64(0) & 0xFFFFFFFFull);
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % STAGES, buf = it & 1;
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            mbar_wait(&a_full[buf], (it >> 1) & 1);
            tc_fence_after_sync();
            if (leader) {
                const uint32_t y_lo = ((smem_u32(smem + (size_t)s * kC1WStageBytes) & 0x3FFFFu) >> 4) | desc_lo_flags;
                const uint32_t a_col = tmem0 + COL_A + (uint32_t)(buf * 128 + j * 64);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_f16_ts(tmem0 + COL_D + (uint32_t)(j * 32), a_col + 8 * kk, desc_hi | (uint64_t)(y_lo + 64 * kk), idesc,
                                (it | kk) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[s]);
                umma_commit(&a_empty[buf]);
            }
            __syncwarp();
        }
        if (leader) umma_commit(&done_bar);
        __syncwarp();
    } else if (warp >= 4 && warp < 12) {
        // ======================= convert warps: uint8 channel rows -> fp16 pairs -> tensor memory
        const int q = warp & 3, kh = (warp - 4) >> 2;
        const int tapslot = q >> 1;
        const int c = (q & 1) * 32 + lane;
        const uint32_t lane_base = tmem0 + COL_A + ((uint32_t)(q * 32) << 16) + (uint32_t)(kh * 32);
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % STAGES, buf = it & 1;
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            const uint8_t* src = smem + (size_t)s * kC1WStageBytes + kC1WYBytes + c * 176 + (tapslot ? 16 : 0) + kh * 64;
            uint32_t W[24];
#pragma unroll
            for (int v = 0; v < 6; ++v) {
                const int4 t = *reinterpret_cast<const int4*>(src + 16 * v);
                W[4 * v] = (uint32_t)t.x; W[4 * v + 1] = (uint32_t)t.y; W[4 * v + 2] = (uint32_t)t.z; W[4 * v + 3] = (uint32_t)t.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);             // this warp's bytes are in registers
            if (it >= 2) mbar_wait(&a_empty[buf], ((it >> 1) - 1) & 1);
            tc_fence_after_sync();
            uint32_t o[32];
            // tile 0 = first tap of the slot (shift 0 / 21), tile 1 = second tap (shift 1 / 22); tapslot 1 reads from byte 16,
            // so its windows start 5 / 6 bytes in: word 1, byte 1 / 2
            if (tapslot == 0) u8_window_to_f16<0>(W, o); else u8_window_to_f16<1>(W + 1, o);
            tmem_st32(lane_base + (uint32_t)(buf * 128), o);
            if (tapslot == 0) u8_window_to_f16<1>(W, o); else u8_window_to_f16<2>(W + 1, o);
            tmem_st32(lane_base + (uint32_t)(buf * 128 + 64), o);
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[buf]);
        }
    } else if (warp >= 12) {
        // ======================= dY warps: bias gradient = column sums of dY from the staged tiles (fp32, fixed order)
        const int tb = tid - 384;
        const int rq = tb >> 2, c16 = tb & 3;
        float bsum[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) bsum[e] = 0.f;
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % STAGES;
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            const uint8_t* sY = smem + (size_t)s * kC1WStageBytes;
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                const int rr = ps * 32 + rq;
                const int4 v = *reinterpret_cast<const int4*>(sY + img64_off(rr, c16));
                const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                    bsum[2 * e] += f.x;
                    bsum[2 * e + 1] += f.y;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) sRed[rq * 32 + c16 * 8 + e] = bsum[e];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tb < 32) {
            float t = 0.f;
#pragma unroll
            for (int l = 0; l < 32; ++l) t += sRed[l * 32 + tb];
            p.wsb[(int64_t)blockIdx.x * 64 + tb] = t;
        }
        // drain the two accumulator tiles
        if (nsteps > 0) {
            mbar_wait(&done_bar, 0);
            tc_fence_after_sync();
        }
        const int lrow = (warp & 3) * 32 + lane;
        float* wsc = p.ws + (int64_t)blockIdx.x * 256 * 64;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint32_t v[32];
            if (nsteps > 0) {
                tmem_ld32(tmem0 + COL_D + (uint32_t)(j * 32) + ((uint32_t)((warp & 3) * 32) << 16), v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = 0u;
            }
            float4* dst = reinterpret_cast<float4*>(wsc + (int64_t)(j * 128 + lrow) * 64);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                     __uint_as_float(v[4 * e + 3]));
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tmem0, TMEM_COLS);
}

static int launch_conv1_wgrad_u8(const Conv1WgradU8Params& p, const void* frames_cm, int64_t n_images, const void* dact1_f16, int ctas,
                                 cudaStream_t s, const char* what) {
    const size_t smem = (size_t)kC1WStages * kC1WStageBytes + 1024;
    static SmemAttrCache attr;
    int rc;
    if ((rc = attr.ensure(tc_conv1_wgrad_u8, smem, what))) return rc;
    if (p.rows_per_cta % 128 != 0) return fail(B200RL_ERR_INVALID_ARGUMENT, "%s: rows per CTA must be a multiple of 128", what);
    CUtensorMap tmX, tmY;
    memset(&tmX, 0, sizeof(tmX)); memset(&tmY, 0, sizeof(tmY));
    // frames [img][64 ch][448 rows] u8: box = [64 ch][176 rows]; dY [img][441 rows][32 co] fp16 = 64-byte rows: box [128 rows][64 B]
    if ((rc = make_tmap_3d_u8(&tmX, frames_cm, n_images, 64, 448, 448, 64, 176, what))) return rc;
    if ((rc = make_tmap_3d_u8(&tmY, dact1_f16, p.n, 441, 64, 64, 128, 64, what))) return rc;
    tc_conv1_wgrad_u8<<<ctas, 512, smem, s>>>(tmX, tmY, p);
    return check_launch(what);
}

constexpr int kConv1I8Stages = 12;
static int launch_conv1_i8(const Conv1U8Params& p, const void* frames_rm, cudaStream_t s, const char* what) {
    constexpr int STAGES = kConv1I8Stages;
    const size_t smem = (size_t)4 * 64 * 64 + (size_t)STAGES * 152 * 64 + 1024;
    static SmemAttrCache attr;
    if (int rc = attr.ensure(tc_conv1_i8<STAGES>, smem, what)) return rc;
    const int total = p.n * 4;                     // 4 tiles of 128 grid rows per image (441 used)
    int grid = num_sms();
    if (grid > total) grid = total;
    CUtensorMap tmA;
    memset(&tmA, 0, sizeof(tmA));
    int rc = make_tmap_3d_u8(&tmA, frames_rm, p.n_images, 441, 64, 64, 152, 64, what);
    if (rc) return rc;
    tc_conv1_i8<STAGES><<<grid, kConvWinThreads, smem, s>>>(tmA, p, total);
    return check_launch(what);
}

}  // namespace b200rl
