// Hardware experiment: may the start address of an MN-major SWIZZLE_64B B operand (N = 32 fp16, rows = K index, 64 B
// each) be shifted by whole K rows that are NOT a multiple of the 8-row swizzle atom?  (K-major operands may:
// umma_shift_test.cu / umma_ts_test.cu.)  If yes, tc_conv1_wgrad_u8 can read the dY rows of the tap group shifted by
// 21 grid rows from the same staged tile as the unshifted group (plus a 24-row halo) instead of a second TMA box.
//   D[128 x 32] = A[128 x 32 (K)] (fp16, TMEM) * B[shift .. shift + 32)[N = 32]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../cleanrl_b200/csrc/tc_common.cuh"
using namespace b200rl::tc;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
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
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int b_mn) {
    return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t img64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__global__ void k_shift(const __half* A, const __half* B, float* D, int shift) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    uint8_t* sB = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));       // 64 rows (K) x 64 B
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(&tbase, 64);
    for (int i = tid; i < 64 * 4; i += blockDim.x) { int r = i >> 2, c = i & 3; *(int4*)(sB + img64_off(r, c)) = *(const int4*)((const uint8_t*)B + r * 64 + c * 16); }
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
        const uint32_t idesc = idesc_f16(128, 32, 1);
        for (int kk = 0; kk < 2; ++kk)
            umma_f16_ts(td, ta + 8 * kk, desc_sw64(smem_u32(sB) + shift * 64 + kk * 1024), idesc, kk != 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
                 "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(td + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 32; ++e) D[tid * 32 + e] = __uint_as_float(v[e]);
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(td, 64);
}

int main() {
    std::vector<float> Af(128 * 32), Bf(64 * 32);
    std::vector<__half> Ah(128 * 32), Bh(64 * 32);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) { float v = (float)((m * 7 + k * 3) % 17) - 8.f; Af[m * 32 + k] = v; Ah[m * 32 + k] = __float2half(v); }
    for (int k = 0; k < 64; ++k) for (int n = 0; n < 32; ++n) { float v = (float)(((k * 5 + n * 3 + (k * n) % 7) % 13) - 6) * 0.0625f; Bf[k * 32 + n] = v; Bh[k * 32 + n] = __float2half(v); }
    __half *Ad, *Bd; float* Dd;
    cudaMalloc(&Ad, Ah.size() * 2); cudaMalloc(&Bd, Bh.size() * 2); cudaMalloc(&Dd, 128 * 32 * 4);
    cudaMemcpy(Ad, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(Bd, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
    const int shifts[] = {0, 8, 16, 1, 2, 3, 4, 5, 7, 11, 21, 27};
    for (int s : shifts) {
        cudaMemset(Dd, 0, 128 * 32 * 4);
        k_shift<<<1, 128, 8192>>>(Ad, Bd, Dd, s);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> D(128 * 32);
        cudaMemcpy(D.data(), Dd, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += (double)Af[m * 32 + k] * Bf[(k + s) * 32 + n];
            maxerr = fmax(maxerr, fabs(ref - D[m * 32 + n]));
        }
        printf("MN-major SW64 B (N=32), start shifted by %2d K rows: %s maxerr %.4f %s\n", s, maxerr < 1e-2 ? "OK " : "BAD", maxerr,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
        if (e != cudaSuccess) break;
    }
    return 0;
}
