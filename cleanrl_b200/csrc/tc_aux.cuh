// Frame conversion (uint8 NCHW -> space-to-depth bf16) and weight packing kernels.
#pragma once
#include "tc_base.cuh"

namespace b200rl {
using namespace tc;

// uint8 frames [n,4,84,84] (NCHW, as envpool delivers them) -> space-to-depth bf16 [n,21,21,64] with
// channel = c*16 + sy*4 + sx for source pixel (4Y+sy, 4X+sx).  conv1 (8x8, stride 4) becomes a 2x2,
// stride-1 convolution over 64-channel NHWC pixels, i.e. the same 128-byte-per-tap gather as conv2/conv3.
// Done ONCE per environment step; the minibatch updates then read the bf16 rollout directly.
__global__ void __launch_bounds__(256) tc_frames_to_s2d(const uint8_t* __restrict__ obs, const int64_t* __restrict__ rows,
                                                        int64_t n, bf16* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;    // ((i*21 + Y)*21 + X)*4 + c
    if (idx >= n * 21 * 21 * 4) return;
    const int c = (int)(idx & 3);
    int64_t t = idx >> 2;
    const int X = (int)(t % 21); t /= 21;
    const int Y = (int)(t % 21);
    const int64_t i = t / 21;
    const int64_t img = rows ? rows[i] : i;
    const uint8_t* src = obs + img * 28224 + c * 7056 + (Y * 4) * 84 + X * 4;
    uint32_t o[8];
#pragma unroll
    for (int sy = 0; sy < 4; ++sy) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(src + sy * 84));
        o[2 * sy] = pack_bf16x2((float)(w & 0xFF), (float)((w >> 8) & 0xFF));
        o[2 * sy + 1] = pack_bf16x2((float)((w >> 16) & 0xFF), (float)(w >> 24));
    }
    int4* dst = reinterpret_cast<int4*>(out + (idx >> 2) * 64 + c * 16);
    dst[0] = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
    dst[1] = make_int4((int)o[4], (int)o[5], (int)o[6], (int)o[7]);
}

// ------------------------------------------------------------------ weight packing (fp32 master -> bf16 GEMM operands)
// conv weight w[co][c][ky][kx] -> fwd[co][(ky,kx,c)] (nhwc_k) or [co][(c,ky,kx)] (conv1), and
// dgrad[c][(ky,kx,co)] with taps FLIPPED implicitly by the loader's negative offsets (no flip needed here).
__global__ void tc_pack_conv(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int nchw_k,
                             bf16* __restrict__ fwd, bf16* __restrict__ dgrad) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)Cout * Cin * KH * KW;
    if (idx >= total) return;
    int kx = (int)(idx % KW); int64_t t = idx / KW;
    int ky = (int)(t % KH); t /= KH;
    int c = (int)(t % Cin); int co = (int)(t / Cin);
    const bf16 v = __float2bfloat16(w[idx]);
    const int K = Cin * KH * KW;
    if (nchw_k) fwd[(int64_t)co * K + (c * KH + ky) * KW + kx] = v;
    else fwd[(int64_t)co * K + (ky * KW + kx) * Cin + c] = v;
    if (dgrad) dgrad[(int64_t)c * (KH * KW * Cout) + (ky * KW + kx) * Cout + co] = v;
}
// conv1 weight w[co][c][ky][kx] (8x8) -> [co][(a,b), c*16 + sy*4 + sx] with ky = 4a+sy, kx = 4b+sx
// (K order of the space-to-depth frames)
__global__ void tc_pack_conv1_s2d(const float* __restrict__ w, bf16* __restrict__ fwd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 32 * 256) return;
    const int kx = idx & 7, ky = (idx >> 3) & 7, c = (idx >> 6) & 3, co = idx >> 8;
    const int a = ky >> 2, sy = ky & 3, b = kx >> 2, sx = kx & 3;
    fwd[co * 256 + (a * 2 + b) * 64 + c * 16 + sy * 4 + sx] = __float2bfloat16(w[idx]);
}
// conv2 weight w[co][c][ky][kx] (4x4, stride 2) -> [co][(a,b) tap][(py,px,c)] with ky = 2a+py, kx = 2b+px:
// K order of the 2x2-cell (space-to-depth 2) activations [n,10,10,128]
__global__ void tc_pack_conv2_cells(const float* __restrict__ w, bf16* __restrict__ fwd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 64 * 512) return;
    const int kx = idx & 3, ky = (idx >> 2) & 3, c = (idx >> 4) & 31, co = idx >> 9;
    const int a = ky >> 1, py = ky & 1, b = kx >> 1, px = kx & 1;
    fwd[co * 512 + (a * 2 + b) * 128 + (py * 2 + px) * 32 + c] = __float2bfloat16(w[idx]);
}
// conv2 data-gradient weights per stride-parity class: dg[cls][c][(a,b,co)] = w[co][c][py+2a][px+2b]
__global__ void tc_pack_conv_s2_classes(const float* __restrict__ w, int Cout, int Cin, bf16* __restrict__ dg) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)4 * Cin * 4 * Cout;
    if (idx >= total) return;
    int co = (int)(idx % Cout); int64_t t = idx / Cout;
    int ab = (int)(t % 4); t /= 4;
    int c = (int)(t % Cin); int cls = (int)(t / Cin);
    const int py = cls >> 1, px = cls & 1, a = ab >> 1, b = ab & 1;
    dg[idx] = __float2bfloat16(w[(((int64_t)co * Cin + c) * 4 + (py + 2 * a)) * 4 + (px + 2 * b)]);
}
// fc weight w[o][c*49+p] -> fwd[o][p*64+c]; dgrad[p*64+c][o].  Block = 8 output rows x 7 pixels x all 64 channels,
// staged through shared memory so that both packed layouts are written with 16-byte stores
// (fwd: 8 consecutive c of one (o, p); dgrad: the 8 o of one (p, c)).  Requires O % 8 == 0, PP % 7 == 0, C == 64.
__global__ void __launch_bounds__(256) tc_pack_fc(const float* __restrict__ w, int O, int PP, bf16* __restrict__ fwd,
                                                  bf16* __restrict__ dgrad) {
    constexpr int C = 64, PS = 7, R = 8;
    __shared__ float sw[R * C * PS];                    // [r][c][pl]
    const int K = C * PP;
    const int o0 = blockIdx.x * R, p0 = blockIdx.y * PS;
    for (int i = threadIdx.x; i < R * C * PS; i += blockDim.x) {
        const int pl = i % PS, rc = i / PS;             // rc = r*64 + c
        sw[i] = w[(int64_t)(o0 + (rc >> 6)) * K + (rc & 63) * PP + p0 + pl];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * PS * (C / 8); i += blockDim.x) {      // fwd[o][p*64 + c0 .. c0+7]
        const int g = i & 7, pl = (i >> 3) % PS, r = i / (8 * PS);
        const float* src = sw + (r * C + g * 8) * PS + pl;
        int4 v;
        v.x = (int)pack_bf16x2(src[0 * PS], src[1 * PS]); v.y = (int)pack_bf16x2(src[2 * PS], src[3 * PS]);
        v.z = (int)pack_bf16x2(src[4 * PS], src[5 * PS]); v.w = (int)pack_bf16x2(src[6 * PS], src[7 * PS]);
        *reinterpret_cast<int4*>(fwd + (int64_t)(o0 + r) * K + (p0 + pl) * C + g * 8) = v;
    }
    for (int i = threadIdx.x; i < PS * C; i += blockDim.x) {                 // dgrad[p*64 + c][o0 .. o0+7]
        const int c = i & 63, pl = i >> 6;
        const float* src = sw + c * PS + pl;
        int4 v;
        v.x = (int)pack_bf16x2(src[0 * C * PS], src[1 * C * PS]); v.y = (int)pack_bf16x2(src[2 * C * PS], src[3 * C * PS]);
        v.z = (int)pack_bf16x2(src[4 * C * PS], src[5 * C * PS]); v.w = (int)pack_bf16x2(src[6 * C * PS], src[7 * C * PS]);
        *reinterpret_cast<int4*>(dgrad + ((int64_t)(p0 + pl) * C + c) * O + o0) = v;
    }
}

}  // namespace b200rl
