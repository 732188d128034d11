// LSTM cell kernels for the recurrent PPO agent (reference: cleanrl/ppo_atari_lstm.py:117-160, nn.LSTM(512, 128), one
// layer, gate order i, f, g, o as torch.nn.LSTM).  The four gate pre-activations come from two fp32 GEMMs of this library
// (x W_ih^T + b_ih for ALL steps of a sequence at once, h' W_hh^T + b_hh per step); these kernels are the elementwise parts:
//   mask     : (h', c') = (1 - done) * (h, c)                -- the reference resets the state BEFORE the cell (:137-142)
//   cell fwd : i, f, o = sigmoid, g = tanh;  c = f c' + i g;  h = o tanh(c); keeps (i, f, g, o, tanh c) for the backward
//   cell bwd : one step of back-propagation through time, given dL/dh from the heads, the recurrent dL/dh (raw, from the
//              NEXT step's W_hh GEMM) and the recurrent dL/dc.
// All tensors fp32, hidden index contiguous; one thread per (row, hidden unit).
#include "common.cuh"

namespace b200rl {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256) lstm_mask_kernel(const float* __restrict__ h, const float* __restrict__ c,
                                                        const float* __restrict__ done, int64_t n, int H,
                                                        float* __restrict__ hm, float* __restrict__ cm) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * H) return;
    const float keep = 1.0f - done[idx / H];
    hm[idx] = keep * h[idx];
    cm[idx] = keep * c[idx];
}

__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ gh,
                                                            const float* __restrict__ cm, int64_t n, int H,
                                                            float* __restrict__ h_out, float* __restrict__ c_out,
                                                            float* __restrict__ save) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * H) return;
    const int64_t r = idx / H;
    const int j = (int)(idx - r * H);
    const float* a = gx + r * 4 * H;
    const float* b = gh + r * 4 * H;
    const float i = sigmoidf_(a[j] + b[j]);
    const float f = sigmoidf_(a[H + j] + b[H + j]);
    const float g = tanhf(a[2 * H + j] + b[2 * H + j]);
    const float o = sigmoidf_(a[3 * H + j] + b[3 * H + j]);
    const float c = f * cm[idx] + i * g;
    const float tc = tanhf(c);
    c_out[idx] = c;
    h_out[idx] = o * tc;
    if (save) {
        float* s = save + r * 5 * H;
        s[j] = i; s[H + j] = f; s[2 * H + j] = g; s[3 * H + j] = o; s[4 * H + j] = tc;
    }
}

// dh = dh_heads + (1 - done_next) * dh_rec_raw;  dc = dc_rec + dh o (1 - tanh(c)^2)
// dgates (pre-activation) = [dc g i(1-i), dc c' f(1-f), dc i (1-g^2), dh tanh(c) o(1-o)];  dc_rec_out = (1 - done) dc f
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(const float* __restrict__ dh_heads, const float* __restrict__ dh_rec_raw,
                                                            const float* __restrict__ done_next, const float* __restrict__ dc_rec,
                                                            const float* __restrict__ save, const float* __restrict__ cm,
                                                            const float* __restrict__ done, int64_t n, int H,
                                                            float* __restrict__ dgates, float* __restrict__ dc_rec_out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * H) return;
    const int64_t r = idx / H;
    const int j = (int)(idx - r * H);
    float dh = dh_heads[idx];
    if (dh_rec_raw) dh += (1.0f - done_next[r]) * dh_rec_raw[idx];
    const float* s = save + r * 5 * H;
    const float i = s[j], f = s[H + j], g = s[2 * H + j], o = s[3 * H + j], tc = s[4 * H + j];
    float dc = dh * o * (1.0f - tc * tc);
    if (dc_rec) dc += dc_rec[idx];
    float* dg = dgates + r * 4 * H;
    dg[j] = dc * g * i * (1.0f - i);
    dg[H + j] = dc * cm[idx] * f * (1.0f - f);
    dg[2 * H + j] = dc * i * (1.0f - g * g);
    dg[3 * H + j] = dh * tc * o * (1.0f - o);
    dc_rec_out[idx] = (1.0f - done[r]) * dc * f;
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_lstm_mask_state_f32(const float* h, const float* c, const float* done, int64_t n, int H,
                                          float* h_masked, float* c_masked, void* stream) {
    B200RL_REQUIRE(n >= 0 && H >= 1, "lstm_mask_state: bad sizes");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(h && c && done && h_masked && c_masked, "lstm_mask_state: null pointer");
    lstm_mask_kernel<<<(unsigned)ceil_div(n * H, 256), 256, 0, (cudaStream_t)stream>>>(h, c, done, n, H, h_masked, c_masked);
    return check_launch("lstm_mask_state");
}

extern "C" int b200rl_lstm_cell_fwd_f32(const float* gates_x, const float* gates_h, const float* c_masked, int64_t n, int H,
                                        float* h_out, float* c_out, float* save, void* stream) {
    B200RL_REQUIRE(n >= 0 && H >= 1, "lstm_cell_fwd: bad sizes");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(gates_x && gates_h && c_masked && h_out && c_out, "lstm_cell_fwd: null pointer");
    lstm_cell_fwd_kernel<<<(unsigned)ceil_div(n * H, 256), 256, 0, (cudaStream_t)stream>>>(gates_x, gates_h, c_masked, n, H, h_out, c_out, save);
    return check_launch("lstm_cell_fwd");
}

extern "C" int b200rl_lstm_cell_bwd_f32(const float* dh_heads, const float* dh_rec_raw, const float* done_next, const float* dc_rec,
                                        const float* save, const float* c_masked, const float* done, int64_t n, int H,
                                        float* dgates, float* dc_rec_out, void* stream) {
    B200RL_REQUIRE(n >= 0 && H >= 1, "lstm_cell_bwd: bad sizes");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(dh_heads && save && c_masked && done && dgates && dc_rec_out, "lstm_cell_bwd: null pointer");
    B200RL_REQUIRE(!dh_rec_raw || done_next, "lstm_cell_bwd: dh_rec_raw needs done_next");
    lstm_cell_bwd_kernel<<<(unsigned)ceil_div(n * H, 256), 256, 0, (cudaStream_t)stream>>>(dh_heads, dh_rec_raw, done_next, dc_rec, save,
                                                                                             c_masked, done, n, H, dgates, dc_rec_out);
    return check_launch("lstm_cell_bwd");
}
