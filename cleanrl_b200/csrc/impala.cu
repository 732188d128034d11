// Elementwise / pooling kernels of the IMPALA-CNN agent (reference: cleanrl/ppo_procgen.py:89-150): the 3x3 convolutions
// themselves are the padded fp32 convolution kernels of net_f32.cu.
//   max_pool 3x3, stride 2, padding 1 (ConvSequence.forward, :113): forward keeps the arg-max (0..8, first maximum in
//       row-major window order, as torch), backward gathers dy from the <= 4 windows that contain an input element.
//   relu / relu_bwd / add: the pre-activation residual blocks x + conv1(relu(conv0(relu(x)))) (:96-102).
//   nhwc_u8_to_nchw: procgen frames arrive [n, 64, 64, 3]; the network sees x.permute(0, 3, 1, 2) (:143).
#include "common.cuh"

namespace b200rl {

__global__ void __launch_bounds__(256) maxpool3s2_fwd_kernel(const float* __restrict__ x, int64_t nc, int H, int W, int OH, int OW,
                                                             float* __restrict__ y, uint8_t* __restrict__ arg) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nc * OH * OW) return;
    const int ox = (int)(idx % OW);
    const int64_t t = idx / OW;
    const int oy = (int)(t % OH);
    const int64_t c = t / OH;
    const float* xp = x + c * H * W;
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int iy = oy * 2 + ky - 1, ix = ox * 2 + kx - 1;
            if (iy < 0 || ix < 0 || iy >= H || ix >= W) continue;
            const float v = xp[iy * W + ix];
            if (v > best || (v != v && !(best != best))) { best = v; bi = ky * 3 + kx; }   // first maximum; NaN propagates
        }
    y[idx] = best;
    arg[idx] = (uint8_t)bi;
}

__global__ void __launch_bounds__(256) maxpool3s2_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ arg, int64_t nc,
                                                             int H, int W, int OH, int OW, float* __restrict__ dx) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nc * H * W) return;
    const int ix = (int)(idx % W);
    const int64_t t = idx / W;
    const int iy = (int)(t % H);
    const int64_t c = t / H;
    float s = 0.f;
    // windows (oy, ox) with oy*2 - 1 <= iy <= oy*2 + 1
    for (int oy = (iy) / 2; oy <= (iy + 1) / 2; ++oy) {
        if (oy < 0 || oy >= OH) continue;
        const int ky = iy - (oy * 2 - 1);
        if (ky < 0 || ky > 2) continue;
        for (int ox = (ix) / 2; ox <= (ix + 1) / 2; ++ox) {
            if (ox < 0 || ox >= OW) continue;
            const int kx = ix - (ox * 2 - 1);
            if (kx < 0 || kx > 2) continue;
            const int64_t o = (c * OH + oy) * OW + ox;
            if (arg[o] == ky * 3 + kx) s += dy[o];
        }
    }
    dx[idx] = s;
}

__global__ void __launch_bounds__(256) relu_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fmaxf(x[i], 0.f);
}
// dx = dy * (x > 0) [+ extra]
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ extra,
                                                       int64_t n, float* __restrict__ dx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dx[i] = (x[i] > 0.f ? dy[i] : 0.f) + (extra ? extra[i] : 0.f);
}
__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = a[i] + b[i];
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_u8_kernel(const uint8_t* __restrict__ x, const int64_t* __restrict__ rows, int64_t n, int H,
                                                              int W, int C, uint8_t* __restrict__ y) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // output index ((i*C + c)*H + h)*W + w
    if (idx >= n * C * H * W) return;
    const int w = (int)(idx % W);
    int64_t t = idx / W;
    const int h = (int)(t % H); t /= H;
    const int c = (int)(t % C);
    const int64_t i = t / C;
    const int64_t src = rows ? rows[i] : i;
    y[idx] = x[((src * H + h) * W + w) * C + c];
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_maxpool3s2_fwd_f32(const float* x, int64_t nc, int H, int W, float* y, uint8_t* argmax, void* stream) {
    B200RL_REQUIRE(nc >= 0 && H >= 1 && W >= 1, "maxpool_fwd: bad sizes");
    if (nc == 0) return B200RL_OK;
    B200RL_REQUIRE(x && y && argmax, "maxpool_fwd: null pointer");
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    maxpool3s2_fwd_kernel<<<(unsigned)ceil_div(nc * OH * OW, 256), 256, 0, (cudaStream_t)stream>>>(x, nc, H, W, OH, OW, y, argmax);
    return check_launch("maxpool_fwd");
}
extern "C" int b200rl_maxpool3s2_bwd_f32(const float* dy, const uint8_t* argmax, int64_t nc, int H, int W, float* dx, void* stream) {
    B200RL_REQUIRE(nc >= 0 && H >= 1 && W >= 1, "maxpool_bwd: bad sizes");
    if (nc == 0) return B200RL_OK;
    B200RL_REQUIRE(dy && dx && argmax, "maxpool_bwd: null pointer");
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    maxpool3s2_bwd_kernel<<<(unsigned)ceil_div(nc * H * W, 256), 256, 0, (cudaStream_t)stream>>>(dy, argmax, nc, H, W, OH, OW, dx);
    return check_launch("maxpool_bwd");
}
extern "C" int b200rl_relu_f32(const float* x, int64_t n, float* y, void* stream) {
    B200RL_REQUIRE(n >= 0, "relu: negative n");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(x && y, "relu: null pointer");
    relu_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, y);
    return check_launch("relu");
}
extern "C" int b200rl_relu_bwd_f32(const float* dy, const float* x, const float* extra, int64_t n, float* dx, void* stream) {
    B200RL_REQUIRE(n >= 0, "relu_bwd: negative n");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(dy && x && dx, "relu_bwd: null pointer");
    relu_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, x, extra, n, dx);
    return check_launch("relu_bwd");
}
extern "C" int b200rl_add_f32(const float* a, const float* b, int64_t n, float* y, void* stream) {
    B200RL_REQUIRE(n >= 0, "add: negative n");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(a && b && y, "add: null pointer");
    add_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, y);
    return check_launch("add");
}
extern "C" int b200rl_nhwc_to_nchw_u8(const uint8_t* x, const int64_t* rows, int64_t n, int H, int W, int C, uint8_t* y, void* stream) {
    B200RL_REQUIRE(n >= 0 && H >= 1 && W >= 1 && C >= 1, "nhwc_to_nchw: bad sizes");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(x && y, "nhwc_to_nchw: null pointer");
    nhwc_to_nchw_u8_kernel<<<(unsigned)ceil_div(n * C * H * W, 256), 256, 0, (cudaStream_t)stream>>>(x, rows, n, H, W, C, y);
    return check_launch("nhwc_to_nchw");
}
