// DQN TD update head (config 5): td target from the target network's Q-values, the chosen action's Q-value,
// the loss, and dL/dQ -- one pass.  Replaces cleanrl/dqn_atari.py:220-224 (+ the head part of loss.backward()).
// The reference loss is F.mse_loss (dqn_atari.py:224; docs/rl-algorithms/dqn.md:102 calls that deliberate);
// huber = 1 selects smooth-L1 (delta = 1) as BASELINE.json's config text names it.
// Algorithmic bytes: 2 * 4A (both Q rows) + 8 + 4 + 4 in, 4A out per sample.  Deterministic reductions.
#include "common.cuh"

namespace b200rl {

constexpr int kTdThreads = 256;

struct TdParams {
    const float* q; int64_t ldq;
    const float* qt; int64_t ldqt;
    const int64_t* actions; const float* rewards; const float* dones;
    int64_t B; int A; float gamma; int huber;
    float* dq; int64_t lddq; float* stats; float* partials; unsigned int* ticket;
};

__global__ void __launch_bounds__(kTdThreads) dqn_td_loss_kernel(TdParams P) {
    __shared__ float red[32];
    __shared__ bool is_last;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f, qv = 0.f;
    if (i < P.B) {
        const float* qt = P.qt + i * P.ldqt;
        float mx = qt[0];
        for (int k = 1; k < P.A; ++k) mx = fmaxf(mx, qt[k]);
        // td_target = r + gamma * target_max * (1 - done)   (dqn_atari.py:222; separately rounded like torch)
        const float td = __fadd_rn(P.rewards[i], __fmul_rn(__fmul_rn(P.gamma, mx), __fsub_rn(1.f, P.dones[i])));
        int a = (int)P.actions[i];
        a = a < 0 ? 0 : (a >= P.A ? P.A - 1 : a);
        const float old = P.q[i * P.ldq + a];
        const float x = old - td;
        float g;
        if (P.huber) {
            const float ax = fabsf(x);
            l = ax < 1.f ? 0.5f * x * x : ax - 0.5f;
            g = fminf(fmaxf(x, -1.f), 1.f);
        } else {
            l = x * x;
            g = 2.f * x;
        }
        qv = old;
        const float invB = 1.0f / (float)P.B;
        float* d = P.dq + i * P.lddq;
        for (int k = 0; k < P.A; ++k) d[k] = (k == a) ? g * invB : 0.f;
    }
    const float sl = block_sum(l, red);
    const float sq = block_sum(qv, red);
    if (threadIdx.x == 0) {
        P.partials[2 * blockIdx.x] = sl;
        P.partials[2 * blockIdx.x + 1] = sq;
        __threadfence();
        is_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float a0 = 0.f, a1 = 0.f;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) { a0 += __ldcg(P.partials + 2 * b); a1 += __ldcg(P.partials + 2 * b + 1); }
    a0 = block_sum(a0, red);
    a1 = block_sum(a1, red);
    if (threadIdx.x == 0) {
        P.stats[0] = a0 / (float)P.B;      // losses/td_loss
        P.stats[1] = a1 / (float)P.B;      // losses/q_values (old_val.mean())
        *P.ticket = 0;
    }
}

// argmax_a Q[i, a] (first maximum, torch.argmax) for the greedy branch of the epsilon-greedy policy (dqn_atari.py:192-193)
__global__ void dqn_argmax_kernel(const float* __restrict__ q, int64_t ld, int64_t n, int A, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = q + i * ld;
    int best = 0; float bv = r[0];
    for (int k = 1; k < A; ++k) if (r[k] > bv) { bv = r[k]; best = k; }
    out[i] = best;
}

}  // namespace b200rl

extern "C" size_t b200rl_dqn_td_loss_workspace_bytes(int64_t B) {
    if (B < 0) return 0;
    return 16 + (size_t)b200rl::ceil_div(B > 0 ? B : 1, b200rl::kTdThreads) * 2 * sizeof(float);
}

extern "C" int b200rl_dqn_td_loss_f32(const float* q, int64_t ld_q, const float* q_target_next, int64_t ld_qt,
                                      const int64_t* actions, const float* rewards, const float* dones,
                                      int64_t B, int A, double gamma, int huber,
                                      float* dq, int64_t ld_dq, float* stats,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    using namespace b200rl;
    B200RL_REQUIRE(B >= 1, "dqn_td_loss: B must be >= 1");
    B200RL_REQUIRE(A >= 1 && A <= 64, "dqn_td_loss: A=%d outside [1,64]", A);
    B200RL_REQUIRE(q && q_target_next && actions && rewards && dones && dq && stats, "dqn_td_loss: null pointer");
    B200RL_REQUIRE(ld_q >= A && ld_qt >= A && ld_dq >= A, "dqn_td_loss: bad strides");
    B200RL_REQUIRE(workspace && aligned(workspace, 16), "dqn_td_loss: workspace null or misaligned");
    if (workspace_bytes < b200rl_dqn_td_loss_workspace_bytes(B))
        return fail(B200RL_ERR_WORKSPACE, "dqn_td_loss: workspace %zu < %zu", workspace_bytes, b200rl_dqn_td_loss_workspace_bytes(B));
    cudaStream_t s = (cudaStream_t)stream;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 16);
    ProfScope ps(s, "dqn_td_loss", 0, (double)B * (12.0 * A + 16));
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return fail(B200RL_ERR_CUDA, "dqn_td_loss: memset: %s", cudaGetErrorString(e));
    TdParams P{q, ld_q, q_target_next, ld_qt, actions, rewards, dones, B, A, (float)gamma, huber, dq, ld_dq, stats, partials, ticket};
    dqn_td_loss_kernel<<<(unsigned)ceil_div(B, kTdThreads), kTdThreads, 0, s>>>(P);
    return check_launch("dqn_td_loss");
}

extern "C" int b200rl_argmax_f32(const float* q, int64_t ld_q, int64_t n, int A, int64_t* out, void* stream) {
    using namespace b200rl;
    B200RL_REQUIRE(n >= 0 && A >= 1, "argmax: bad shape");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(q && out && ld_q >= A, "argmax: bad arguments");
    dqn_argmax_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(q, ld_q, n, A, out);
    return check_launch("argmax");
}
