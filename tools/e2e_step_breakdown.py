"""Where one end-to-end rollout step spends its wall time (host side), for the bench.py e2e configuration.

    python tools/e2e_step_breakdown.py [num_envs] [num_steps]

Splits PPOEngine.policy_step into: enqueue (H2D chunks + graph launches), wait (device finishes + actions D2H),
env.step on the host, record_reward.  Measurement tool only."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import ppo_args  # noqa: E402
from cleanrl_b200 import ppo_engine as pe  # noqa: E402
from cleanrl_b200.agents import NatureCNNAgent  # noqa: E402
from cleanrl_b200.synthetic_envs import SyntheticAtariVec  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
args = ppo_args(N, T, 4, "bf16")
envs = SyntheticAtariVec(N, seed=1, mode="pool", pinned=True)
envs.single_observation_space, envs.single_action_space = envs.observation_space, envs.action_space
agent = NatureCNNAgent(envs).to(dev)
agent.precision = "bf16"
eng = pe.PPOEngine(agent, args, (4, 84, 84), np.uint8, N, dev)
acc = {"enqueue": 0.0, "wait": 0.0, "env": 0.0, "record": 0.0}
orig_sync = pe._sync
t_mark = [0.0]


def timed_sync():
    t0 = time.perf_counter()
    acc["enqueue"] += t0 - t_mark[0]
    orig_sync()
    acc["wait"] += time.perf_counter() - t0


obs, done = envs.reset(), np.zeros(N, dtype=np.float32)
for it in range(3):
    if it == 2:
        pe._sync = timed_sync
        for k in acc:
            acc[k] = 0.0
    t_it = time.perf_counter()
    for step in range(T):
        t_mark[0] = time.perf_counter()
        a = eng.policy_step(step, obs, done)
        t1 = time.perf_counter()
        obs, rew, done, info = envs.step(a)
        t2 = time.perf_counter()
        eng.record_reward(step, rew)
        t3 = time.perf_counter()
        acc["env"] += t2 - t1
        acc["record"] += t3 - t2
    torch.cuda.synchronize()
    roll = time.perf_counter() - t_it
    pe._sync = orig_sync                      # the update's own sync must not be charged to the rollout steps
    eng.finish_rollout(obs, done)
    eng.update(2.5e-4)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t_it
pe._sync = orig_sync
print(json.dumps({"num_envs": N, "num_steps": T, "rollout_ms": round(roll * 1e3, 2), "iteration_ms": round(tot * 1e3, 2),
                  "per_step_us": {k: round(v / T * 1e6, 1) for k, v in acc.items()},
                  "h2d_bytes_per_step": N * 28224, "h2d_chunks": eng.h2d_chunks}))
