"""Frame-stack delta upload on the GPU (csrc/frame_stack.cu, PPOEngine._launch_part_delta): uploading only the newest frame
plane and rebuilding the rollout slot on the device must give bit-identical rollout buffers to uploading every observation
whole (what cleanrl/ppo_atari_envpool.py:226,239 does), including resets, the first step, the bootstrap slot, the next
iteration's first slot, and envs that break the shifted-stack contract without being flagged done."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_delta_kernel_equals_full_conversion(lib):
    from cleanrl_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    n = 67
    prev = torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, generator=g)
    new = torch.randint(0, 256, (n, 84, 84), dtype=torch.uint8, generator=g)
    cur = torch.cat([prev[:, 1:], new[:, None]], dim=1).contiguous()
    full_idx = [3, 11, 12, 40, 66]
    fresh = torch.randint(0, 256, (len(full_idx), 4, 84, 84), dtype=torch.uint8, generator=g)
    slot = torch.full((n,), -1, dtype=torch.int32)
    for k, i in enumerate(full_idx):
        slot[i] = k
        cur[i] = fresh[k]
    prm, pcm = ops.frames_to_s2d_u8(prev.to(dev))
    want_rm, want_cm = ops.frames_to_s2d_u8(cur.to(dev))
    out_rm = ops.alloc_u8_rollout_rows((n, 441, 64), dev)
    out_cm = torch.zeros((n, 64, 448), dtype=torch.uint8, device=dev)
    ops.frames_delta_s2d_u8(cur[:, 3].reshape(n, 7056).contiguous().to(dev), prm, pcm, out_rm, out_cm,
                            full_slot=slot.to(dev), full_frames=fresh.reshape(-1, 28224).to(dev))
    assert torch.equal(out_rm, want_rm) and torch.equal(out_cm, want_cm)
    # without a slot table every env is a shifted stack
    cur2 = torch.cat([prev[:, 1:], new[:, None]], dim=1).contiguous()
    want_rm, want_cm = ops.frames_to_s2d_u8(cur2.to(dev))
    ops.frames_delta_s2d_u8(new.reshape(n, 7056).to(dev), prm, pcm, out_rm, out_cm)
    assert torch.equal(out_rm, want_rm) and torch.equal(out_cm, want_cm)


class _SilentResets:
    """A frame-stacked env that, every few steps, changes a plane of some envs WITHOUT flagging them done."""

    def __init__(self, env, every=3):
        self.env, self.every, self.t = env, every, 0
        self.num_envs = env.num_envs

    def reset(self):
        return self.env.reset()

    def step(self, a):
        obs, r, d, info = self.env.step(a)
        self.t += 1
        if self.t % self.every == 0:
            obs = np.array(obs)                         # private copy (the ring itself stays consistent)
            for i in (1, self.num_envs // 2, self.num_envs - 1):
                obs[i, (self.t // self.every) % 3, 10:20, 30] ^= 0x5A
        return obs, r, d, info


def _rollouts(delta, N, T, iters, pinned, wrap=None, mode="stack"):
    from bench import ppo_args
    from cleanrl_b200.agents import NatureCNNAgent
    from cleanrl_b200.ppo_engine import PPOEngine
    from cleanrl_b200.synthetic_envs import SyntheticAtariVec
    dev = torch.device("cuda")
    old = os.environ.get("CLEANRL_B200_DELTA_UPLOAD")
    os.environ["CLEANRL_B200_DELTA_UPLOAD"] = "1" if delta else "0"
    try:
        torch.manual_seed(3)
        spaces = SyntheticAtariVec(2, seed=1)
        spaces.single_observation_space, spaces.single_action_space = spaces.observation_space, spaces.action_space
        agent = NatureCNNAgent(spaces).to(dev); agent.precision = "bf16"
        eng = PPOEngine(agent, ppo_args(N, T, 4, "bf16"), (4, 84, 84), np.uint8, N, dev, gae_mode=1)
    finally:
        if old is None:
            os.environ.pop("CLEANRL_B200_DELTA_UPLOAD", None)
        else:
            os.environ["CLEANRL_B200_DELTA_UPLOAD"] = old
    assert eng.delta_upload == delta
    torch.manual_seed(11)
    parts = [SyntheticAtariVec(N // 2, seed=5 + p, mode=mode, pool=8, p_done=0.05, pinned=pinned) for p in range(2)]
    if wrap is not None:
        parts = [wrap(e) for e in parts]
    obs_p = [e.reset() for e in parts]
    done_p = [np.zeros(N // 2, dtype=np.float32) for _ in parts]
    outs = []
    for _ in range(iters):
        obs_p, done_p = eng.collect(parts, obs_p, done_p)
        eng.finish_rollout_parts(obs_p, done_p)
        torch.cuda.synchronize()
        outs.append({k: getattr(eng, k).clone() for k in ("obs", "obs_t", "next_obs", "next_obs_t", "actions", "logprobs", "values",
                                                          "rewards", "dones", "advantages", "returns")})
    return eng, outs


@pytest.mark.parametrize("pinned", [True, False])
def test_delta_upload_equals_whole_upload(lib, pinned):
    N, T = 256, 11                                       # 2 iterations x 12 observations: the 8-plane ring wraps twice
    e_full, full = _rollouts(False, N, T, 2, pinned)
    e_delta, delta = _rollouts(True, N, T, 2, pinned)
    for it in range(2):
        for k in full[it]:
            assert torch.equal(full[it][k], delta[it][k]), (it, k)
    assert e_delta.delta_redos == 0 and e_delta.delta_upload
    # whole upload: (T + 1) observations per iteration; delta: one whole observation once, then a plane + the resets
    assert e_full.delta_full_frames == 0
    assert N <= e_delta.delta_full_frames < N + 0.12 * N * (2 * T + 1)
    assert e_delta.h2d_bytes < 0.45 * e_full.h2d_bytes


def test_contract_violations_are_redone_from_full_frames(lib):
    N, T = 256, 9                                        # iteration 1: python launch path (captures), iteration 2: one-call plans
    _, full = _rollouts(False, N, T, 2, False, wrap=_SilentResets)
    e_delta, delta = _rollouts(True, N, T, 2, False, wrap=_SilentResets)
    for it in range(2):
        for k in full[it]:
            assert torch.equal(full[it][k], delta[it][k]), (it, k)
    assert e_delta.delta_redos >= 8 and e_delta.delta_upload       # a few bad envs: redo, keep the delta path
    assert len(e_delta._plans) > 0


def test_unstacked_env_falls_back_to_whole_uploads(lib):
    N, T = 128, 6
    _, full = _rollouts(False, N, T, 1, True, mode="pool")
    e_delta, delta = _rollouts(True, N, T, 1, True, mode="pool")
    for k in full[0]:
        assert torch.equal(full[0][k], delta[0][k]), k
    assert e_delta.delta_redos >= 1 and not e_delta.delta_upload
