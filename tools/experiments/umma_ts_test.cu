// Hardware experiment: tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form) fed by tcgen05.st from
// registers -- the path conv1 uses to consume uint8 frames (u8 -> fp16 in registers -> TMEM, never through a
// 16-bit shared-memory image).  Checks
//   1. numerics of D[128 x N] = A[128 x 64] * B[N x 64]^T with A written by tcgen05.st.32x32b (lane = row m,
//      column j holds K elements 2j (low half) and 2j+1 (high half)), fp16 operands, fp32 accumulate;
//   2. issue rate: cycles per M128 x N x K16 MMA for the SS form (A from a SWIZZLE_128B shared-memory image)
//      and the TS form, N = 32 / 64 / 128 (the SS form of a narrow-N MMA is bound by the 4 KB of A it re-reads
//      from shared memory; the TS form is not).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../cleanrl_b200/csrc/tc_common.cuh"
using namespace b200rl::tc;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
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

// fp16 x fp16 -> fp32 (a_format = b_format = 0), K-major A and B
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode 0: TS numerics.  A [128][64] fp16 (global, row-major), B [N][64] fp16; D [128][N] fp32
// mode 1/2: timing of `reps` MMAs (1 = SS, 2 = TS), result cycles in D[0]
__global__ void k(const __half* A, const __half* B, float* D, int N, int mode, int reps, int swap_halves) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 128 rows x 128 B (SS form only)
    uint8_t* sB = smem + 128 * 128;     // up to 128 rows x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(&tbase, 256);
    for (int i = tid; i < 128 * 8; i += blockDim.x) { int r = i >> 3, c = i & 7; *(int4*)(sA + img_off(r, c)) = *(const int4*)(A + r * 64 + c * 8); }
    for (int i = tid; i < N * 8; i += blockDim.x) { int r = i >> 3, c = i & 7; *(int4*)(sB + img_off(r, c)) = *(const int4*)(B + r * 64 + c * 8); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t td = tbase;            // D: columns [0, N)
    const uint32_t ta = tbase + 128;      // A: columns [128, 160): 32 columns = 64 fp16 per lane
    // every thread = one row m of A: 64 fp16 -> 32 packed registers -> its TMEM lane
    {
        uint32_t v[32];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A + tid * 64);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            uint32_t w = src[j];
            if (swap_halves) w = (w >> 16) | (w << 16);
            v[j] = w;
        }
        tmem_st32(ta + ((uint32_t)(warp * 32) << 16), v);
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = idesc_f16(128, N);
        const uint64_t ad = desc_kmajor(smem_u32(sA)), bd = desc_kmajor(smem_u32(sB));
        if (mode == 0) {
            for (int kk = 0; kk < 4; ++kk) umma_f16_ts(td, ta + 8 * kk, bd + 2 * kk, idesc, kk != 0);
        } else {
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (mode == 1) umma_bf16(td, ad + 2 * kk, bd + 2 * kk, idesc, 1u);
                    else umma_f16_ts(td, ta + 8 * kk, bd + 2 * kk, idesc, 1u);
                }
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    if (tid == 0 && mode != 0) { t1 = clock64(); D[0] = (float)(t1 - t0) / (float)(4 * reps); }
    if (mode == 0) {
        uint32_t v[16];
        for (int c0 = 0; c0 < N; c0 += 16) {
            tmem_ld16(td + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int e = 0; e < 16; ++e) D[tid * N + c0 + e] = __uint_as_float(v[e]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(td, 256);
}

// ---------------------------------------------------------------------------------------------------------
// kind::i8: D[128 x N] (s32) = A[shift + m][k] (u8) * B[n][k] (s8), K = 64 bytes per row, SWIZZLE_64B images
// (row pitch 64 B, 8-row atoms of 512 B), K-major; the A descriptor starts `shift` rows into a 160-row window.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {      // c = s32 (2), a = u8 (0), b = s8 (1)
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t desc_kmajor_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t img64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__global__ void k_i8(const uint8_t* A, const int8_t* B, int* D, int N, int shift, int reps, float* cyc) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 160 rows x 64 B
    uint8_t* sB = smem + 160 * 64;      // N rows x 64 B (160*64 = 10240 = 10 x 1024: still 1024-aligned)
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(&tbase, 128);
    for (int i = tid; i < 160 * 4; i += blockDim.x) { int r = i >> 2, c = i & 3; *(int4*)(sA + img64_off(r, c)) = *(const int4*)(A + r * 64 + c * 16); }
    for (int i = tid; i < N * 4; i += blockDim.x) { int r = i >> 2, c = i & 3; *(int4*)(sB + img64_off(r, c)) = *(const int4*)(B + r * 64 + c * 16); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t td = tbase;
    long long t0 = 0;
    if (tid == 0) {
        const uint32_t idesc = idesc_i8(128, N);
        const uint64_t ad = desc_kmajor_sw64(smem_u32(sA) + shift * 64), bd = desc_kmajor_sw64(smem_u32(sB));
        t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int kk = 0; kk < 2; ++kk) umma_i8(td, ad + 2 * kk, bd + 2 * kk, idesc, (r | kk) != 0);   // K = 32 bytes per MMA
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    if (tid == 0) cyc[0] = (float)(clock64() - t0) / (float)(2 * reps);
    uint32_t v[16];
    for (int c0 = 0; c0 < N; c0 += 16) {
        tmem_ld16(td + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int e = 0; e < 16; ++e) D[tid * N + c0 + e] = (int)v[e];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(td, 128);
}

static void run_i8() {
    std::vector<uint8_t> A(160 * 64); std::vector<int8_t> B(128 * 64);
    for (int r = 0; r < 160; ++r) for (int c = 0; c < 64; ++c) A[r * 64 + c] = (uint8_t)((r * 37 + c * 11 + (r * c) % 5) & 255);
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 64; ++c) B[r * 64 + c] = (int8_t)(((r * 13 + c * 7) % 255) - 127);
    uint8_t* Ad; int8_t* Bd; int* Dd; float* Cd;
    cudaMalloc(&Ad, A.size()); cudaMalloc(&Bd, B.size()); cudaMalloc(&Dd, 128 * 128 * 4); cudaMalloc(&Cd, 4);
    cudaMemcpy(Ad, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(Bd, B.data(), B.size(), cudaMemcpyHostToDevice);
    const size_t smem = 160 * 64 + 128 * 64 + 2048;
    for (int N = 32; N <= 96; N += 64) {
        const int shifts[] = {0, 1, 2, 5, 21, 22};
        for (int si = 0; si < 6; ++si) {
            cudaMemset(Dd, 0, 128 * 128 * 4);
            k_i8<<<1, 128, smem>>>(Ad, Bd, Dd, N, shifts[si], 1, Cd);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<int> D(128 * N);
            cudaMemcpy(D.data(), Dd, D.size() * 4, cudaMemcpyDeviceToHost);
            long long maxerr = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
                long long ref = 0;
                for (int kk = 0; kk < 64; ++kk) ref += (long long)A[(shifts[si] + m) * 64 + kk] * B[n * 64 + kk];
                long long d = llabs(ref - D[m * N + n]); if (d > maxerr) maxerr = d;
            }
            printf("i8 SS K-major SW64 N=%2d row shift %2d : %s maxerr %lld %s\n", N, shifts[si], maxerr == 0 ? "EXACT" : "BAD", maxerr,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
            if (e != cudaSuccess) return;
        }
        k_i8<<<1, 128, smem>>>(Ad, Bd, Dd, N, 0, 2048, Cd);
        cudaError_t e = cudaDeviceSynchronize();
        float cyc = 0; cudaMemcpy(&cyc, Cd, 4, cudaMemcpyDeviceToHost);
        printf("issue rate  i8 SS N=%2d : %.1f cycles per M128xN%dxK32 MMA %s\n", N, cyc, N, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    std::vector<float> Af(128 * 64), Bf(128 * 64);
    std::vector<__half> Ah(128 * 64), Bh(128 * 64);
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 64; ++c) { float v = (float)((r * 7 + c * 3) % 256); Af[r * 64 + c] = v; Ah[r * 64 + c] = __float2half(v); }
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 64; ++c) { float v = (float)(((r * 5 + c) % 7) - 3) * 0.125f; Bf[r * 64 + c] = v; Bh[r * 64 + c] = __float2half(v); }
    __half *Ad, *Bd; float* Dd;
    cudaMalloc(&Ad, Ah.size() * 2); cudaMalloc(&Bd, Bh.size() * 2); cudaMalloc(&Dd, 128 * 128 * 4);
    cudaMemcpy(Ad, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(Bd, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = 128 * 128 * 2 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int N = 32; N <= 128; N *= 2) {
        for (int sw = 0; sw < 2; ++sw) {
            cudaMemset(Dd, 0, 128 * 128 * 4);
            k<<<1, 128, smem>>>(Ad, Bd, Dd, N, 0, 0, sw);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> D(128 * N);
            cudaMemcpy(D.data(), Dd, D.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int kk = 0; kk < 64; ++kk) ref += (double)Af[m * 64 + kk] * Bf[n * 64 + kk];
                double d = fabs(ref - D[m * N + n]); if (d > maxerr) maxerr = d;
            }
            printf("TS numerics N=%3d halves %s : %s maxerr %.3f %s\n", N, sw ? "swapped" : "natural (low = even k)",
                   maxerr < 1e-2 ? "OK " : "BAD", maxerr, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
        for (int mode = 1; mode <= 2; ++mode) {
            k<<<1, 128, smem>>>(Ad, Bd, Dd, N, mode, 2048, 0);
            cudaError_t e = cudaDeviceSynchronize();
            float cyc = 0; cudaMemcpy(&cyc, Dd, 4, cudaMemcpyDeviceToHost);
            printf("issue rate  N=%3d %s : %.1f cycles per M128xN%dxK16 MMA %s\n", N, mode == 1 ? "SS (A in smem)" : "TS (A in TMEM)", cyc, N,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    run_i8();
    return 0;
}
