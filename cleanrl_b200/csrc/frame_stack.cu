// Frame-stack aware observation upload (include/b200rl.h, "frame-stack delta upload").
//
// The Atari observation of cleanrl/ppo_atari_envpool.py:185-196,237-239 is a stack of the 4 newest 84x84 frames: from
// one env step to the next, planes 0..2 of an env's observation are planes 1..3 of its previous one unless the env was
// reset.  The reference uploads the whole stack as fp32 every step (115.6 MB at 1024 envs); the round-1/2 engine uploads
// it as uint8 (28.9 MB) and is PCIe-bound.  Here only the newest plane crosses PCIe (7.2 MB):
//
//   host   b200rl_stackdelta_*   a worker pool keeps a private mirror of the last observation of every env and VERIFIES
//                                on the host, asynchronously to the upload, that the shifted-stack property really holds
//                                for every env that is not flagged done (memcmp of 3 planes per env); envs flagged done
//                                are staged as full frames up front, envs that fail the check are reported so that the
//                                caller can re-stage them as full frames and redo the step (never a silent error).
//   device tc_frames_delta_s2d_u8   builds rollout slot t (both uint8 space-to-depth orientations, tc_conv1_u8.cuh) from
//                                slot t-1 (channel groups 1..3 -> 0..2) plus the newest plane (-> channel group 3), or
//                                from a full staged frame for the envs that have one.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b200rl {

// ------------------------------------------------------------------------------------------------ device
// grid = 4 n blocks; block (i, c) owns channel group c (= frame plane c) of frame i:
//   row-major    out_rm[i][pos][c*16 .. c*16+15]   (16 bytes = 4 rows x 4 pixels of grid position pos)
//   channel-major out_cm[i][c*16 + sy*4 + sx][pos]  (448-byte rows, positions 441..447 zero)
__global__ void __launch_bounds__(256) tc_frames_delta_s2d_u8(const uint8_t* __restrict__ newp, const int32_t* __restrict__ full_slot,
                                                              const uint8_t* __restrict__ full, const uint8_t* __restrict__ prev_rm,
                                                              const uint8_t* __restrict__ prev_cm, int64_t n,
                                                              uint8_t* __restrict__ out_rm, uint8_t* __restrict__ out_cm) {
    __shared__ __align__(16) uint8_t plane[7056];
    const int64_t i = blockIdx.x >> 2;
    const int c = blockIdx.x & 3;
    const int slot = full_slot ? full_slot[i] : -1;
    if (slot < 0 && c < 3) {
        // shifted stack: plane c of this observation is plane c+1 of the previous one (already in storage layout)
        const uint8_t* prm = prev_rm + i * 28224 + (c + 1) * 16;
        uint8_t* orm = out_rm + i * 28224 + c * 16;
        for (int pos = threadIdx.x; pos < 441; pos += 256)
            *reinterpret_cast<int4*>(orm + pos * 64) = __ldg(reinterpret_cast<const int4*>(prm + pos * 64));
        const int4* pcm = reinterpret_cast<const int4*>(prev_cm + (i * 64 + (c + 1) * 16) * 448);
        int4* ocm = reinterpret_cast<int4*>(out_cm + (i * 64 + c * 16) * 448);
        for (int t = threadIdx.x; t < 16 * 28; t += 256) ocm[t] = __ldg(pcm + t);
        return;
    }
    const uint8_t* srcp = slot >= 0 ? full + (int64_t)slot * 28224 + c * 7056 : newp + i * 7056;
    const int4* src = reinterpret_cast<const int4*>(srcp);
    for (int t = threadIdx.x; t < 441; t += 256) reinterpret_cast<int4*>(plane)[t] = __ldg(src + t);
    __syncthreads();
    for (int pos = threadIdx.x; pos < 441; pos += 256) {
        const int Y = pos / 21, X = pos - Y * 21;
        const uint8_t* p = plane + (Y * 4) * 84 + X * 4;
        int4 v;
        v.x = *reinterpret_cast<const int*>(p); v.y = *reinterpret_cast<const int*>(p + 84);
        v.z = *reinterpret_cast<const int*>(p + 168); v.w = *reinterpret_cast<const int*>(p + 252);
        *reinterpret_cast<int4*>(out_rm + (i * 441 + pos) * 64 + c * 16) = v;
    }
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

// ------------------------------------------------------------------------------------------------ host
// One tracker per vector env (or env group).  A small pool of worker threads (sleeping on a condition variable when idle)
// runs two kinds of passes over env ranges: PACK (synchronous: copy planes into pinned staging) and VERIFY
// (asynchronous: compare against the mirror, refresh the mirror).
struct StackDelta {
    int64_t n, plane_bytes;
    int planes;
    std::vector<uint8_t> mirror;     // [n][planes][plane_bytes]; logical plane j of the last observation = ring slot (base + j) % planes
    int base = 0;
    bool primed = false;

    enum Kind { NONE, PACK_NEW, PACK_FULL, VERIFY };
    struct Job {
        Kind kind = NONE;
        const uint8_t* obs = nullptr;
        int64_t env_stride = 0;
        uint8_t* out = nullptr;            // PACK_NEW: new_out, PACK_FULL: full_out
        const int32_t* slot = nullptr;
    } job;
    // ---- work distribution: one item = kGrain consecutive envs.  Everything a worker needs to claim an item (the job, the
    // next unclaimed item) and to report it (items done, mismatches) is guarded by `mu`, so a worker that wakes up late can
    // never run an item of a newer job with an older job's description, and a pass is complete when its ITEMS are done --
    // it never waits for every worker to have woken up.
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    int64_t items_total = 0, next_item = 0, items_done = 0;
    bool stop = false;
    bool verify_pending = false;
    std::vector<int32_t> mismatch;         // guarded by mu
    static constexpr int64_t kGrain = 16;  // envs per work item

    StackDelta(int64_t n_, int planes_, int64_t pb, int threads) : n(n_), plane_bytes(pb), planes(planes_) {
        mirror.resize((size_t)n * planes * pb);
        for (int t = 0; t < threads; ++t) workers.emplace_back([this] { loop(); });
    }
    ~StackDelta() {
        { std::lock_guard<std::mutex> g(mu); stop = true; }
        cv_work.notify_all();
        for (auto& w : workers) w.join();
    }
    uint8_t* mslot(int64_t i, int s) { return mirror.data() + ((size_t)i * planes + s) * plane_bytes; }

    void run_range(const Job& j, int64_t a, int64_t b, std::vector<int32_t>& bad) {
        const int P = planes;
        const int64_t pb = plane_bytes;
        for (int64_t i = a; i < b; ++i) {
            const uint8_t* cur = j.obs + i * j.env_stride;
            switch (j.kind) {
                case PACK_NEW: memcpy(j.out + i * pb, cur + (P - 1) * pb, (size_t)pb); break;
                case PACK_FULL: if (j.slot[i] >= 0) memcpy(j.out + (int64_t)j.slot[i] * P * pb, cur, (size_t)(P * pb)); break;
                case VERIFY: {
                    bool shifted = j.slot[i] < 0;
                    if (shifted)
                        for (int p = 0; p + 1 < P; ++p)
                            if (memcmp(cur + p * pb, mslot(i, (base + p + 1) % P), (size_t)pb) != 0) { shifted = false; break; }
                    if (shifted) {
                        memcpy(mslot(i, base), cur + (P - 1) * pb, (size_t)pb);
                    } else {
                        if (j.slot[i] < 0) bad.push_back((int32_t)i);
                        for (int p = 0; p < P; ++p) memcpy(mslot(i, (base + 1 + p) % P), cur + p * pb, (size_t)pb);
                    }
                    break;
                }
                default: break;
            }
        }
    }
    void loop() {
        std::vector<int32_t> bad;
        std::unique_lock<std::mutex> g(mu);
        for (;;) {
            cv_work.wait(g, [&] { return stop || next_item < items_total; });
            if (stop) return;
            const int64_t item = next_item++;
            const Job j = job;
            const bool more = next_item < items_total;
            g.unlock();
            if (more) cv_work.notify_one();          // chained wake-up: the submitting thread pays for ONE futex wake
            bad.clear();
            const int64_t a = item * kGrain;
            run_range(j, a, a + kGrain < n ? a + kGrain : n, bad);
            g.lock();
            if (!bad.empty()) mismatch.insert(mismatch.end(), bad.begin(), bad.end());
            if (++items_done == items_total) cv_done.notify_all();
        }
    }
    void start(const Job& j) {               // no pass may be in flight
        if (workers.empty()) {
            std::vector<int32_t> bad;
            run_range(j, 0, n, bad);
            mismatch.insert(mismatch.end(), bad.begin(), bad.end());
            return;
        }
        {
            std::lock_guard<std::mutex> g(mu);
            job = j;
            next_item = 0;
            items_done = 0;
            items_total = (n + kGrain - 1) / kGrain;
        }
        cv_work.notify_one();
    }
    bool join(double timeout_s = 60.0) {     // false = the pass did not finish in time (a bug: report, never hang)
        if (workers.empty()) return true;
        std::unique_lock<std::mutex> g(mu);
        return cv_done.wait_for(g, std::chrono::duration<double>(timeout_s), [&] { return items_done == items_total; });
    }
    bool finish_verify() {
        if (!verify_pending) return true;
        if (!join()) return false;
        base = (base + 1) % planes;
        primed = true;
        verify_pending = false;
        return true;
    }
};

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_frames_delta_s2d_u8(const uint8_t* new_planes, const int32_t* full_slot, const uint8_t* full_frames,
                                          const uint8_t* prev_rm, const uint8_t* prev_cm, int64_t n,
                                          uint8_t* out_rm, uint8_t* out_cm, void* stream) {
    B200RL_REQUIRE(n >= 0, "frames_delta_s2d_u8: negative n");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(new_planes && prev_rm && prev_cm && out_rm && out_cm, "frames_delta_s2d_u8: null pointer");
    B200RL_REQUIRE(!full_slot || full_frames, "frames_delta_s2d_u8: full_slot without full_frames");
    B200RL_REQUIRE(aligned(new_planes, 16) && aligned(prev_rm, 16) && aligned(prev_cm, 16) && aligned(out_rm, 16) && aligned(out_cm, 16) &&
                   aligned(full_frames, 16) && aligned(full_slot, 4), "frames_delta_s2d_u8: misaligned buffer");
    B200RL_REQUIRE(out_rm != prev_rm && out_cm != prev_cm, "frames_delta_s2d_u8: in-place update is not supported");
    B200RL_REQUIRE(n <= (int64_t)1 << 28, "frames_delta_s2d_u8: n too large");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(s, "frames_delta", 0, (double)n * (7056 + 21168 + 21504 + 28224 + 28672));
    tc_frames_delta_s2d_u8<<<(unsigned)(n * 4), 256, 0, s>>>(new_planes, full_slot, full_frames, prev_rm, prev_cm, n, out_rm, out_cm);
    return check_launch("frames_delta_s2d_u8");
}

extern "C" int b200rl_h2d_rows_async(void* dst, const void* src, int64_t src_pitch, int64_t row_bytes, int64_t rows, void* stream) {
    B200RL_REQUIRE(rows >= 0 && row_bytes >= 0 && src_pitch >= row_bytes, "h2d_rows_async: bad geometry");
    if (rows == 0 || row_bytes == 0) return B200RL_OK;
    B200RL_REQUIRE(dst && src, "h2d_rows_async: null pointer");
    cudaError_t e = src_pitch == row_bytes
        ? cudaMemcpyAsync(dst, src, (size_t)(rows * row_bytes), cudaMemcpyHostToDevice, (cudaStream_t)stream)
        : cudaMemcpy2DAsync(dst, (size_t)row_bytes, src, (size_t)src_pitch, (size_t)row_bytes, (size_t)rows, cudaMemcpyHostToDevice,
                            (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(B200RL_ERR_CUDA, "h2d_rows_async: %s", cudaGetErrorString(e));
    return B200RL_OK;
}

extern "C" void* b200rl_stackdelta_create(int64_t n_envs, int planes, int64_t plane_bytes, int threads) {
    if (n_envs < 1 || planes < 2 || planes > 64 || plane_bytes < 1 || threads < 0 || threads > 256) {
        fail(B200RL_ERR_INVALID_ARGUMENT, "stackdelta_create: bad arguments (n_envs=%lld planes=%d plane_bytes=%lld threads=%d)",
             (long long)n_envs, planes, (long long)plane_bytes, threads);
        return nullptr;
    }
    try {
        return new StackDelta(n_envs, planes, plane_bytes, threads);
    } catch (const std::exception& e) {
        fail(B200RL_ERR_INVALID_ARGUMENT, "stackdelta_create: %s", e.what());
        return nullptr;
    }
}

extern "C" void b200rl_stackdelta_destroy(void* h) {
    if (!h) return;
    StackDelta* sd = reinterpret_cast<StackDelta*>(h);
    if (sd->verify_pending) sd->join();
    delete sd;
}

extern "C" void b200rl_stackdelta_invalidate(void* h) {
    if (!h) return;
    StackDelta* sd = reinterpret_cast<StackDelta*>(h);
    sd->finish_verify();
    sd->primed = false;
}

extern "C" int64_t b200rl_stackdelta_begin(void* h, const uint8_t* obs, int64_t env_stride, const float* done,
                                           uint8_t* new_out, uint8_t* full_out, int32_t* slot_out) {
    if (!h || !obs || !full_out || !slot_out) return (int64_t)fail(B200RL_ERR_INVALID_ARGUMENT, "stackdelta_begin: null pointer");
    StackDelta* sd = reinterpret_cast<StackDelta*>(h);
    if (env_stride < sd->planes * sd->plane_bytes)
        return (int64_t)fail(B200RL_ERR_INVALID_ARGUMENT, "stackdelta_begin: env_stride %lld smaller than one observation", (long long)env_stride);
    if (sd->verify_pending)
        return (int64_t)fail(B200RL_ERR_INVALID_ARGUMENT, "stackdelta_begin: the previous pass was not joined (call b200rl_stackdelta_wait)");
    int64_t k = 0;
    for (int64_t i = 0; i < sd->n; ++i) slot_out[i] = (!sd->primed || (done && done[i] != 0.f)) ? (int32_t)k++ : -1;
    StackDelta::Job j;
    j.obs = obs; j.env_stride = env_stride; j.slot = slot_out;
    if (k > 0) {
        if (k <= 32) {                       // a handful of resets: cheaper than waking the pool
            std::vector<int32_t> none;
            j.kind = StackDelta::PACK_FULL; j.out = full_out;
            sd->run_range(j, 0, sd->n, none);
        } else {
            j.kind = StackDelta::PACK_FULL; j.out = full_out;
            sd->start(j);
            if (!sd->join()) return (int64_t)fail(B200RL_ERR_UNSUPPORTED, "stackdelta_begin: worker pool timed out");
        }
    }
    if (new_out) {
        j.kind = StackDelta::PACK_NEW; j.out = new_out;
        sd->start(j);
        if (!sd->join()) return (int64_t)fail(B200RL_ERR_UNSUPPORTED, "stackdelta_begin: worker pool timed out");
    }
    j.kind = StackDelta::VERIFY; j.out = nullptr;
    sd->mismatch.clear();
    sd->verify_pending = true;
    sd->start(j);
    return k;
}

extern "C" int64_t b200rl_stackdelta_wait(void* h, int32_t* mismatch_out) {
    if (!h) return (int64_t)fail(B200RL_ERR_INVALID_ARGUMENT, "stackdelta_wait: null handle");
    StackDelta* sd = reinterpret_cast<StackDelta*>(h);
    if (!sd->finish_verify()) return (int64_t)fail(B200RL_ERR_UNSUPPORTED, "stackdelta_wait: worker pool timed out");
    const int64_t m = (int64_t)sd->mismatch.size();
    if (m > 0 && mismatch_out) {
        std::vector<int32_t> s(sd->mismatch);
        std::sort(s.begin(), s.end());
        memcpy(mismatch_out, s.data(), (size_t)m * sizeof(int32_t));
    }
    sd->mismatch.clear();
    return m;
}

// ------------------------------------------------------------------------------------------------ one-call launcher
// Everything one env group needs enqueued for one rollout step, in ONE call from the host language (the per-step host
// overhead of the grouped loop is what bounds the end-to-end rollout once only a plane per env crosses PCIe):
// classify + stage + start the verification (tracker), H2D of the staged pieces on the copy stream, the captured step
// graphs of every chunk on the main stream behind their upload events, the actions D2H and its event.
extern "C" int64_t b200rl_stackdelta_launch(const B200rlPartLaunch* p, const uint8_t* obs, int64_t env_stride, const float* done) {
    if (!p || p->nchunks < 1 || p->nchunks > 4) return (int64_t)fail(B200RL_ERR_INVALID_ARGUMENT, "stackdelta_launch: bad plan");
    cudaStream_t cs = (cudaStream_t)p->copy_stream, ms = (cudaStream_t)p->main_stream;
    int64_t k = 0;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess && r != cudaSuccess) e = r; };
    if (p->tracker) {
        if (!obs) return (int64_t)fail(B200RL_ERR_INVALID_ARGUMENT, "stackdelta_launch: null observation batch");
        StackDelta* sd = reinterpret_cast<StackDelta*>(p->tracker);
        const int64_t pb = sd->plane_bytes, fb = sd->planes * sd->plane_bytes;
        bool pinned = false;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, obs) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
        else cudaGetLastError();
        k = b200rl_stackdelta_begin(p->tracker, obs, env_stride, done, pinned ? nullptr : p->new_h, p->full_h, p->slot_h);
        if (k < 0) return k;
        ok(cudaStreamWaitEvent(cs, (cudaEvent_t)p->consumed_event, 0));
        if (k > 0) ok(cudaMemcpyAsync(p->full_d, p->full_h, (size_t)(k * fb), cudaMemcpyHostToDevice, cs));
        ok(cudaMemcpyAsync(p->slot_d, p->slot_h, (size_t)p->n * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
        for (int c = 0; c < p->nchunks; ++c) {
            const int64_t lo = p->chunk_lo[c], rows = p->chunk_hi[c] - p->chunk_lo[c];
            if (pinned)
                ok(cudaMemcpy2DAsync(p->new_d + lo * pb, (size_t)pb, obs + lo * env_stride + (sd->planes - 1) * pb, (size_t)env_stride,
                                     (size_t)pb, (size_t)rows, cudaMemcpyHostToDevice, cs));
            else
                ok(cudaMemcpyAsync(p->new_d + lo * pb, p->new_h + lo * pb, (size_t)(rows * pb), cudaMemcpyHostToDevice, cs));
            ok(cudaEventRecord((cudaEvent_t)p->h2d_event[c], cs));
        }
    }
    for (int c = 0; c < p->nchunks; ++c) {
        if (p->tracker) ok(cudaStreamWaitEvent(ms, (cudaEvent_t)p->h2d_event[c], 0));
        ok(cudaGraphLaunch((cudaGraphExec_t)p->graph_exec[c], ms));
    }
    ok(cudaEventRecord((cudaEvent_t)p->consumed_event, ms));
    if (p->actions_bytes > 0) {
        ok(cudaMemcpyAsync(p->actions_h, p->actions_d, (size_t)p->actions_bytes, cudaMemcpyDeviceToHost, ms));
        ok(cudaEventRecord((cudaEvent_t)p->d2h_event, ms));
    }
    if (e != cudaSuccess) return (int64_t)fail(B200RL_ERR_CUDA, "stackdelta_launch: %s", cudaGetErrorString(e));
    return k;
}

extern "C" int64_t b200rl_stackdelta_join(void* tracker, void* d2h_event, int32_t* mismatch_out) {
    if (d2h_event) {
        cudaError_t e = cudaEventSynchronize((cudaEvent_t)d2h_event);
        if (e != cudaSuccess) return (int64_t)fail(B200RL_ERR_CUDA, "stackdelta_join: %s", cudaGetErrorString(e));
    }
    if (!tracker) return 0;
    StackDelta* sd = reinterpret_cast<StackDelta*>(tracker);
    if (!sd->verify_pending) return 0;
    return b200rl_stackdelta_wait(tracker, mismatch_out);
}

// ------------------------------------------------------------------------------------------------ minibatch shuffle
// numpy.random.shuffle of the reference's index vector (cleanrl/ppo.py:245, the GLOBAL legacy RandomState = MT19937),
// restated natively: the Fisher-Yates walk `for i = n-1 .. 1: j = interval(i); swap(x[i], x[j])` with numpy's masked
// rejection sampler over 32-bit draws.  Consumes the generator exactly as numpy does (the caller hands in / takes back
// the 624-word state), so the permutation AND every later numpy draw are bit-identical -- at a fraction of the host
// time, which is all the host contributes to an update once its kernels are replayed from CUDA graphs.
namespace b200rl {
struct MT19937 {
    uint32_t* key;
    int pos;
    void gen() {
        constexpr int N = 624, M = 397;
        constexpr uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
        int i = 0;
        uint32_t y;
        for (; i < N - M; ++i) { y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & A); }
        for (; i < N - 1; ++i) { y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & A); }
        y = (key[N - 1] & UP) | (key[0] & LO);
        key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        pos = 0;
    }
    uint32_t next32() {
        if (pos == 624) gen();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
};
}  // namespace b200rl

extern "C" int b200rl_mt19937_shuffle_i64(uint32_t* key624, int32_t* pos, int64_t* data, int64_t n) {
    B200RL_REQUIRE(key624 && pos && (data || n == 0), "mt19937_shuffle: null pointer");
    B200RL_REQUIRE(*pos >= 0 && *pos <= 624, "mt19937_shuffle: generator position %d outside [0, 624]", (int)*pos);
    B200RL_REQUIRE(n >= 0 && n <= 0xffffffffLL, "mt19937_shuffle: n outside the 32-bit draw path");
    b200rl::MT19937 g{key624, (int)*pos};
    for (int64_t i = n - 1; i >= 1; --i) {
        uint32_t mask = (uint32_t)i;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t j;
        while ((j = (g.next32() & mask)) > (uint32_t)i) {}
        const int64_t t = data[j];
        data[j] = data[i];
        data[i] = t;
    }
    *pos = g.pos;
    return B200RL_OK;
}
