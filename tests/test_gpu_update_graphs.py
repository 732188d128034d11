"""The update replayed as per-epoch CUDA graphs (PPOEngine._capture_epoch; Adam's step / lr scalars from a device table,
b200rl_clip_adam_dyn_f32) must leave exactly the parameters, optimiser state and logged statistics of the launch-by-launch
update (cleanrl/ppo.py:233-293) -- iteration after iteration, with a changing learning rate."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(graphs, iters=4):
    from bench import ppo_args
    from cleanrl_b200.agents import NatureCNNAgent
    from cleanrl_b200.ppo_engine import PPOEngine
    from cleanrl_b200.synthetic_envs import SyntheticAtariVec
    dev = torch.device("cuda")
    N, T = 64, 16
    np.random.seed(5); torch.manual_seed(5)
    spaces = SyntheticAtariVec(2, seed=1)
    spaces.single_observation_space, spaces.single_action_space = spaces.observation_space, spaces.action_space
    agent = NatureCNNAgent(spaces).to(dev); agent.precision = "bf16"
    eng = PPOEngine(agent, ppo_args(N, T, iters, "bf16"), (4, 84, 84), np.uint8, N, dev, gae_mode=1)
    eng.update_graphs = graphs
    torch.manual_seed(9)
    env = SyntheticAtariVec(N, seed=3, mode="fresh")
    obs, done = env.reset(), np.zeros(N, dtype=np.float32)
    out = []
    for it in range(iters):
        for t in range(T):
            a = eng.policy_step(t, obs, done)
            obs, r, done, _ = env.step(a.copy())
            eng.record_reward(t, r)
        eng.finish_rollout(obs, done)
        st = eng.update(2.5e-4 * (1.0 - it / iters))
        out.append((st["per_update"].copy(), eng.flat.flat.clone(), eng.flat.exp_avg.clone(), eng.flat.exp_avg_sq.clone(), eng.flat.step))
    return eng, out


def test_graph_update_equals_launch_by_launch(lib):
    e0, eager = _run(False)
    e1, graph = _run(True)
    assert len(e1._upd_graphs) == 4 and len(e0._upd_graphs) == 0          # one graph per epoch, captured in iteration 2
    for it, (a, b) in enumerate(zip(eager, graph)):
        assert a[4] == b[4], it
        assert np.array_equal(a[0], b[0]), f"statistics differ in iteration {it}"
        for x, y in zip(a[1:4], b[1:4]):
            assert torch.equal(x, y), f"parameters / Adam state differ in iteration {it}"


def test_adam_scalars_from_device_match_by_value(lib):
    from cleanrl_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    P = 10007
    base = [torch.randn(P, generator=g).to(dev) for _ in range(2)] + [torch.rand(P, generator=g).to(dev)]
    grads = torch.randn(P, generator=g).to(dev)
    for step, lr in ((1, 2.5e-4), (17, 1.3e-4), (4000, 7e-6)):
        a = [t.clone() for t in base]
        b = [t.clone() for t in base]
        ops.clip_adam(a[0], grads, a[1], a[2], step, lr, max_norm=0.5)
        sc = torch.tensor(ops.adam_step_scalars(step, lr), dtype=torch.float32, device=dev)
        ops.clip_adam_dyn(b[0], grads, b[1], b[2], sc, max_norm=0.5)
        for x, y in zip(a, b):
            assert torch.equal(x, y), (step, lr)
