// Hardware experiment: can a SWIZZLE_128B UMMA operand start at a row offset that is not a multiple of 8
// (i.e. a start address that is 128-B but not 1024-B aligned)?  Tries base_offset = 0 and
// base_offset = (addr >> 7) & 7 for K-major A (M-direction shift) and MN-major A (K-direction shift).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../cleanrl_b200/csrc/tc_common.cuh"
using namespace b200rl::tc;
typedef __nv_bfloat16 bf16;

// mode 0: K-major:  D[m][n] = sum_k A[shift+m][k] * B[n][k]          (A rows = M index), K = 64
// mode 1: MN-major: D[m][n] = sum_r X[shift+r][m] * Y[r][n], r < 32   (rows = reduction index), M = 64 cols of X.. use M=64? -> use 128 via 2 images
__global__ void k(const bf16* A, const bf16* B, float* D, int shift, int use_base_off, int mode) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 256 rows x 128 B (mode 1: two images of 128 rows each: cols 0-63 / 64-127)
    uint8_t* sB = smem + 256 * 128;     // 64 rows x 128 B
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (tid < 32) tmem_alloc(&tbase, 32);
    // stage
    if (mode == 0) {
        for (int i = tid; i < 256 * 8; i += blockDim.x) { int r = i >> 3, c = i & 7; *(int4*)(sA + img_off(r, c)) = *(const int4*)(A + r * 64 + c * 8); }
        for (int i = tid; i < 32 * 8; i += blockDim.x) { int r = i >> 3, c = i & 7; *(int4*)(sB + img_off(r, c)) = *(const int4*)(B + r * 64 + c * 8); }
    } else {
        // X: [128 rows][128 cols] -> image0 = cols 0..63, image1 = cols 64..127 (each 128 rows x 128 B)
        for (int i = tid; i < 128 * 16; i += blockDim.x) { int r = i >> 4, c = i & 15; *(int4*)(sA + (c >> 3) * 128 * 128 + img_off(r, c & 7)) = *(const int4*)(A + r * 128 + c * 8); }
        // Y: [64 rows][64 cols (only 32 used as N)]
        for (int i = tid; i < 64 * 8; i += blockDim.x) { int r = i >> 3, c = i & 7; *(int4*)(sB + img_off(r, c)) = *(const int4*)(B + r * 64 + c * 8); }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t td = tbase;
    if (tid == 0) {
        if (mode == 0) {
            const uint32_t idesc = make_idesc(128, 32, 0, 0);
            uint32_t a_addr = smem_u32(sA) + shift * 128;
            uint64_t ad = desc_kmajor(a_addr), bd = desc_kmajor(smem_u32(sB));
            if (use_base_off) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
            for (int kk = 0; kk < 4; ++kk) umma_bf16(td, ad + 2 * kk, bd + 2 * kk, idesc, kk != 0);
        } else {
            const uint32_t idesc = make_idesc(128, 32, 1, 1);
            for (int kk = 0; kk < 2; ++kk) {   // 32 reduction rows = 2 x K16
                uint32_t a_addr = smem_u32(sA) + (shift + kk * 16) * 128;
                uint32_t b_addr = smem_u32(sB) + (kk * 16) * 128;
                uint64_t ad = desc_mnmajor(a_addr, 128 * 128), bd = desc_mnmajor(b_addr, 64 * 128);
                if (use_base_off) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
                umma_bf16(td, ad, bd, idesc, kk != 0);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    if (tid < 128) {
        const int warp = tid >> 5;
        uint32_t v[16];
        for (int c0 = 0; c0 < 32; c0 += 16) {
            tmem_ld16(td + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int e = 0; e < 16; ++e) D[tid * 32 + c0 + e] = __uint_as_float(v[e]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(td, 32);
}

int main() {
    std::vector<float> Af(256 * 128), Bf(64 * 64);
    std::vector<bf16> Ah(256 * 128), Bh(64 * 64);
    bf16 *Ad, *Bd; float* Dd;
    cudaMalloc(&Ad, Ah.size() * 2); cudaMalloc(&Bd, Bh.size() * 2); cudaMalloc(&Dd, 128 * 32 * 4);
    const size_t smem = 256 * 128 + 64 * 128 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 2; ++mode) {
        const int lda = mode == 0 ? 64 : 128;
        for (int r = 0; r < 256; ++r) for (int c = 0; c < lda; ++c) { float v = (float)(((r * 7 + c * 3) % 13) - 6); Af[r * lda + c] = v; Ah[r * lda + c] = __float2bfloat16(v); }
        for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c) { float v = (float)(((r * 5 + c) % 7) - 3); Bf[r * 64 + c] = v; Bh[r * 64 + c] = __float2bfloat16(v); }
        cudaMemcpy(Ad, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(Bd, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
        const int shifts[] = {0, 1, 2, 3, 5, 8, 13, 21, 22, 40};
        for (int si = 0; si < 10; ++si) {
            const int shift = shifts[si];
            for (int ub = 0; ub < 2; ++ub) {
                cudaMemset(Dd, 0, 128 * 32 * 4);
                k<<<1, 128, smem>>>(Ad, Bd, Dd, shift, ub, mode);
                cudaError_t e = cudaDeviceSynchronize();
                std::vector<float> D(128 * 32);
                cudaMemcpy(D.data(), Dd, D.size() * 4, cudaMemcpyDeviceToHost);
                double maxerr = 0;
                for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
                    double ref = 0;
                    if (mode == 0) { for (int kk = 0; kk < 64; ++kk) ref += Af[(shift + m) * 64 + kk] * Bf[n * 64 + kk]; }
                    else { for (int r = 0; r < 32; ++r) ref += Af[(shift + r) * 128 + m] * Bf[r * 64 + n]; }
                    double d = fabs(ref - D[m * 32 + n]); if (d > maxerr) maxerr = d;
                }
                printf("mode %d (%s) shift %2d base_offset=%s : %s maxerr %.1f %s\n", mode, mode ? "MN-major, K shift" : "K-major, M shift", shift,
                       ub ? "(addr>>7)&7" : "0", maxerr < 0.5 ? "OK " : "BAD", maxerr, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
