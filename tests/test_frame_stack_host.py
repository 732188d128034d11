"""Host side of the frame-stack delta upload (b200rl_stackdelta_*, csrc/frame_stack.cu) and the frame-stacked synthetic
env, on CPU: the tracker's classification / staging / verification against a numpy restatement of the contract
`obs[t][:, :3] == obs[t-1][:, 1:]` (envpool stack_num=4, cleanrl/ppo_atari_envpool.py:185-196)."""
import numpy as np
import pytest

from cleanrl_b200 import ops
from cleanrl_b200.synthetic_envs import SyntheticAtariVec


@pytest.fixture(scope="module", autouse=True)
def _built():
    from cleanrl_b200 import build
    build.build()


def test_stack_env_is_frame_stacked_across_the_ring_wrap():
    env = SyntheticAtariVec(24, seed=3, mode="stack", pool=8, p_done=0.1)
    prev = env.reset().copy()
    seen_done = 0
    for t in range(40):                                   # 5 wraps of the 8-plane ring
        obs, r, done, info = env.step(np.random.default_rng(t).integers(0, 4, size=24))
        assert obs.shape == (24, 4, 84, 84) and obs.dtype == np.uint8 and obs.strides[1:] == (7056, 84, 1)
        for i in range(24):
            shifted = np.array_equal(obs[i, :3], prev[i, 1:])
            assert shifted == (not done[i]), (t, i, done[i])
        seen_done += int(done.sum())
        prev = obs.copy()
    assert seen_done > 10


@pytest.mark.parametrize("threads", [0, 3])
def test_tracker_matches_numpy_contract(threads):
    rng = np.random.default_rng(0)
    n, P, pb = 37, 4, 7056
    tr = ops.StackDeltaTracker(n, P, pb, threads=threads, pinned=False)
    big = rng.integers(0, 256, size=(n, 9, 84, 84), dtype=np.uint8)          # strided source: env stride 9 planes
    prev = None
    for t in range(6):
        obs = big[:, t % 5:t % 5 + 4]
        done = (rng.random(n) < 0.2).astype(np.float32)
        silent = np.zeros(n, dtype=bool)
        if t in (2, 4):                                   # contract violations the env does not flag
            bad = rng.choice(n, size=3, replace=False)
            for i in bad:
                if not done[i]:
                    big[i, t % 5 + int(rng.integers(0, 3)), 5, 7] ^= 0xFF
                    silent[i] = True
        if t == 3:
            obs = np.ascontiguousarray(obs)               # dense batches work too
        k = tr.begin(obs, done, pack_new=True)
        slot = tr.slot_h.numpy().copy()
        full = done != 0 if prev is not None else np.ones(n, dtype=bool)
        assert k == int(full.sum())
        assert np.array_equal(slot >= 0, full) and np.array_equal(slot[full], np.arange(k))
        assert np.array_equal(tr.new_h.numpy().reshape(n, 84, 84), obs[:, 3])
        assert np.array_equal(tr.full_h.numpy()[:k].reshape(k, 4, 84, 84), obs[full])
        mis = tr.wait()
        if prev is None:
            assert len(mis) == 0
        else:
            expect = np.array([i for i in range(n) if not full[i] and not np.array_equal(obs[i, :3], prev[i, 1:])], dtype=np.int32)
            assert np.array_equal(mis, expect)
            assert set(np.nonzero(silent)[0]) <= set(expect.tolist())
        prev = obs.copy()
    # after invalidate() every env is staged as a full frame again
    tr.invalidate()
    assert tr.begin(big[:, 0:4], None) == n
    assert len(tr.wait()) == 0
    # the mirror is the last observation: an identical "shifted" batch verifies cleanly, a begin() without wait() is an error
    assert tr.begin(big[:, 1:5], None) == 0
    with pytest.raises(RuntimeError):
        tr.begin(big[:, 2:6], None)
    assert len(tr.wait()) == 0


def test_tracker_rejects_bad_geometry():
    tr = ops.StackDeltaTracker(4, 4, 7056, threads=0, pinned=False)
    with pytest.raises(AssertionError):
        tr.begin(np.zeros((4, 4, 84, 84), dtype=np.uint8)[:, :, ::2], None)
    with pytest.raises(AssertionError):
        tr.begin(np.zeros((4, 3, 84, 84), dtype=np.uint8), None)


@pytest.mark.timeout(120)
def test_worker_pool_never_stalls_with_more_threads_than_items():
    """Regression: a pass is complete when its work items are done, not when every worker has woken up (a pool that counted
    workers dead-locked when a wake-up was consumed by a worker that had already finished its share)."""
    rng = np.random.default_rng(1)
    n = 64                                                # 8 work items, 16 workers
    big = rng.integers(0, 256, size=(n, 12, 84, 84), dtype=np.uint8)
    trs = [ops.StackDeltaTracker(n, threads=16, pinned=False) for _ in range(2)]
    for it in range(600):
        w = it % 8
        for tr in trs:
            tr.begin(big[:, w:w + 4], (rng.random(n) < 0.05).astype(np.float32), pack_new=(it % 3 == 0))
        for tr in trs:
            exp_bad = 0 if (it == 0 or w != 0) else None
            mis = tr.wait()
            if exp_bad == 0:
                assert len(mis) == 0
