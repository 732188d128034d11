// Hardware experiment: what does one tcgen05.mma cost when consecutive MMAs accumulate into the SAME TMEM tile
// (a K loop) versus round-robin over several independent accumulators?  Every narrow-N convolution kernel of this
// library measured 65-79 SM cycles per M128 x N<=64 x K16 MMA regardless of its byte traffic; umma_ts_test.cu
// measured ~93 cycles for a dependent chain at N = 32..128 in both the SS and the TS form.  If the limit is the
// accumulate latency of a dependent chain, interleaving independent chains divides it.
//   part 1: cycles per MMA for chains interleaved over 1 / 2 / 4 accumulators, SS f16, TS f16 and SS i8, N = 32..256
//   part 2: numerics of a mixed-format MMA (A = fp16 in TMEM, B = bf16 in shared memory, MN-major SWIZZLE_64B with
//           N = 32): the operand forms the uint8 weight-gradient kernel needs.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../cleanrl_b200/csrc/tc_common.cuh"
using namespace b200rl::tc;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_bf16, int b_bf16, int b_mn) {
    return (1u << 4) | ((uint32_t)a_bf16 << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t smem_addr) {      // K-major or single-atom MN-major SWIZZLE_64B
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t img64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// ---- part 1: issue-rate.  kind: 0 = SS f16, 1 = TS f16, 2 = SS i8.  nacc accumulators of N columns each.
__global__ void k_rate(int kind, int N, int nacc, int reps, float* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                   // 4 windows of 128 rows x 128 B (64 KB), contents irrelevant (zeros)
    uint8_t* sB = smem + 65536;           // 256 rows x 128 B (32 KB)
    const int tid = threadIdx.x;
    for (int i = tid; i < (65536 + 32768) / 16; i += blockDim.x) reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(&tbase, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t td = tbase;
    const uint32_t ta = tbase + 448;      // 64 columns of A (only read)
    long long t0 = 0;
    if (tid == 0) {
        const uint32_t idf = idesc_f16(128, N, 0, 0, 0), idi = idesc_i8(128, N);
        const uint64_t bd = desc_kmajor(smem_u32(sB)), bd64 = desc_sw64(smem_u32(sB));
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll 4
            for (int kk = 0; kk < 4; ++kk) {
                for (int j = 0; j < nacc; ++j) {
                    const uint32_t d = td + (uint32_t)(j * N);
                    const uint32_t a_addr = smem_u32(sA) + (uint32_t)j * 16384u;
                    if (kind == 0) umma_bf16(d, desc_kmajor(a_addr) + 2 * kk, bd + 2 * kk, idf, 1u);
                    else if (kind == 1) umma_f16_ts(d, ta + 8 * kk, bd + 2 * kk, idf, 1u);
                    else umma_i8(d, desc_sw64(a_addr) + 2 * (kk & 1), bd64 + 2 * (kk & 1), idi, 1u);
                }
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    if (tid == 0) out[0] = (float)(clock64() - t0) / (float)(4 * reps * nacc);
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(td, 512);
}

// ---- part 2: D[128 x 32] = A[128 x 32 (K)] (fp16, TMEM) * B[K = 32 rows][N = 32] (bf16 or fp16, MN-major SW64 in smem)
__global__ void k_mixed(const __half* A, const void* B, float* D, int b_bf16) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;                   // 32 rows (K) x 64 B (32 x 16-bit N)
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(&tbase, 64);
    for (int i = tid; i < 32 * 4; i += blockDim.x) { int r = i >> 2, c = i & 3; *(int4*)(sB + img64_off(r, c)) = *(const int4*)((const uint8_t*)B + r * 64 + c * 16); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t td = tbase, ta = tbase + 32;
    {
        uint32_t v[32];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A + tid * 32);     // 32 fp16 = 16 words; upper 16 unused
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = j < 16 ? src[j] : 0u;
        tmem_st32(ta + ((uint32_t)(warp * 32) << 16), v);
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = idesc_f16(128, 32, 0, b_bf16, 1);
        for (int kk = 0; kk < 2; ++kk)           // K step of 16 rows = 2 swizzle atoms of 8 rows = 1024 B
            umma_f16_ts(td, ta + 8 * kk, desc_sw64(smem_u32(sB) + kk * 1024), idesc, kk != 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    uint32_t v[16];
    for (int c0 = 0; c0 < 32; c0 += 16) {
        tmem_ld16(td + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int e = 0; e < 16; ++e) D[tid * 32 + c0 + e] = __uint_as_float(v[e]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(td, 64);
}

// ---- part 3: B = MN-major SWIZZLE_128B rows of 64 fp16 (128 B); the MMA uses N = 32 starting at column `col0` (0 or 32):
// D[128 x 32] = A[128 x 32 (K)] (fp16, TMEM) * B[K = 32 rows][col0 .. col0 + 32)
__global__ void k_halfrow(const __half* A, const __half* B, float* D, int col0) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;                   // 32 rows (K) x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(&tbase, 64);
    for (int i = tid; i < 32 * 8; i += blockDim.x) { int r = i >> 3, c = i & 7; *(int4*)(sB + img_off(r, c)) = *(const int4*)((const uint8_t*)B + r * 128 + c * 16); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t td = tbase, ta = tbase + 32;
    {
        uint32_t v[32];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A + tid * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = j < 16 ? src[j] : 0u;
        tmem_st32(ta + ((uint32_t)(warp * 32) << 16), v);
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = idesc_f16(128, 32, 0, 0, 1);
        for (int kk = 0; kk < 2; ++kk)           // K step of 16 rows = 2 x 8-row groups of 1024 B
            umma_f16_ts(td, ta + 8 * kk, desc_mnmajor(smem_u32(sB) + kk * 2048 + col0 * 2, 64 * 128), idesc, kk != 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    uint32_t v[16];
    for (int c0 = 0; c0 < 32; c0 += 16) {
        tmem_ld16(td + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int e = 0; e < 16; ++e) D[tid * 32 + c0 + e] = __uint_as_float(v[e]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(td, 64);
}

int main() {
    float* out; cudaMalloc(&out, 4);
    const size_t smem = 65536 + 32768 + 1024;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[3] = {"SS f16 (K16)", "TS f16 (K16)", "SS i8  (K32)"};
    for (int kind = 0; kind < 3 && getenv("UMMA_RATES"); ++kind)
        for (int N = 32; N <= 256; N *= 2)
            for (int nacc = 1; nacc <= 4; nacc *= 2) {
                if (N * nacc > 448) continue;
                k_rate<<<1, 128, smem>>>(kind, N, nacc, 1024, out);
                cudaError_t e = cudaDeviceSynchronize();
                float cyc = 0; cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
                printf("rate %s N=%3d accumulators=%d : %6.1f cycles per MMA (ideal %3d) %s\n", names[kind], N, nacc, cyc, N / 2,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    // part 2
    std::vector<float> Af(128 * 32), Bf(32 * 32);
    std::vector<__half> Ah(128 * 32), Bh(32 * 32);
    std::vector<__nv_bfloat16> Bb(32 * 32);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) { float v = (float)((m * 7 + k * 3) % 256); Af[m * 32 + k] = v; Ah[m * 32 + k] = __float2half(v); }
    for (int k = 0; k < 32; ++k) for (int n = 0; n < 32; ++n) { float v = (float)(((k * 5 + n) % 9) - 4) * 0.0625f; Bf[k * 32 + n] = v; Bh[k * 32 + n] = __float2half(v); Bb[k * 32 + n] = __float2bfloat16(v); }
    __half* Ad; void* Bd; float* Dd;
    cudaMalloc(&Ad, Ah.size() * 2); cudaMalloc(&Bd, 32 * 32 * 2); cudaMalloc(&Dd, 128 * 32 * 4);
    cudaMemcpy(Ad, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
    for (int bb = 0; bb < (getenv("UMMA_MIXED") ? 2 : 1); ++bb) {      // bb = 1 (bf16 B with fp16 A) raises 'illegal instruction' on B200
        cudaMemcpy(Bd, bb ? (void*)Bb.data() : (void*)Bh.data(), 32 * 32 * 2, cudaMemcpyHostToDevice);
        cudaMemset(Dd, 0, 128 * 32 * 4);
        k_mixed<<<1, 128, 8192>>>(Ad, Bd, Dd, bb);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> D(128 * 32);
        cudaMemcpy(D.data(), Dd, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += (double)Af[m * 32 + k] * Bf[k * 32 + n];
            maxerr = fmax(maxerr, fabs(ref - D[m * 32 + n]));
        }
        printf("TS A=fp16(TMEM) x B=%s (MN-major SW64, N=32): %s maxerr %.4f %s\n", bb ? "bf16" : "fp16", maxerr < 1e-2 ? "OK " : "BAD", maxerr,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    // part 3
    {
        std::vector<float> Bw(32 * 64);
        std::vector<__half> Bwh(32 * 64);
        for (int k = 0; k < 32; ++k) for (int n = 0; n < 64; ++n) { float v = (float)(((k * 5 + n * 3) % 11) - 5) * 0.0625f; Bw[k * 64 + n] = v; Bwh[k * 64 + n] = __float2half(v); }
        __half* Bd2; cudaMalloc(&Bd2, 32 * 64 * 2);
        cudaMemcpy(Bd2, Bwh.data(), 32 * 64 * 2, cudaMemcpyHostToDevice);
        for (int col0 = 0; col0 <= 32; col0 += 32) {
            cudaMemset(Dd, 0, 128 * 32 * 4);
            k_halfrow<<<1, 128, 8192>>>(Ad, Bd2, Dd, col0);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> D(128 * 32);
            cudaMemcpy(D.data(), Dd, D.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
                double ref = 0;
                for (int k = 0; k < 32; ++k) ref += (double)Af[m * 32 + k] * Bw[k * 64 + col0 + n];
                maxerr = fmax(maxerr, fabs(ref - D[m * 32 + n]));
            }
            printf("TS A x B MN-major SW128 half row (N=32 at column %2d of 64): %s maxerr %.4f %s\n", col0, maxerr < 1e-2 ? "OK " : "BAD", maxerr,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
