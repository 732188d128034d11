"""A/B on one box: the 16 minibatch updates of a bench iteration launched kernel by kernel vs replayed as per-epoch CUDA
graphs.  Prints, per call of PPOEngine.update, the device time (CUDA events) and the host time spent shuffling / enqueueing /
waiting.    python tools/update_graph_ab.py"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import ppo_args  # noqa: E402
from cleanrl_b200 import build  # noqa: E402
from cleanrl_b200.agents import NatureCNNAgent  # noqa: E402
from cleanrl_b200.ppo_engine import PPOEngine  # noqa: E402
from cleanrl_b200.synthetic_envs import SyntheticAtariVec  # noqa: E402

build.build()
N, T = 1024, 128
dev = torch.device("cuda:0")
args = ppo_args(N, T, 64, "bf16")
np.random.seed(1); torch.manual_seed(1)
envs = SyntheticAtariVec(N, seed=1, mode="pool", pinned=True)
envs.single_observation_space, envs.single_action_space = envs.observation_space, envs.action_space
agent = NatureCNNAgent(envs).to(dev)
agent.precision = "bf16"
eng = PPOEngine(agent, args, (4, 84, 84), np.uint8, N, dev, gae_mode=1)
pool_dev = torch.from_numpy(envs._batches).to(dev)
g = torch.Generator().manual_seed(1)
eng.rollout_resident(pool_dev)
eng.rewards.copy_(torch.randint(0, 2, (T, N), generator=g).float().to(dev))
eng.dones.copy_((torch.rand(T, N, generator=g) < 0.02).float().to(dev))
eng._to_storage(pool_dev[0], None)
eng.finish_rollout(None, None, resident=True)
out = {}
for mode in ("eager", "graphs", "eager", "graphs"):
    eng.update_graphs = mode == "graphs"
    rows = []
    for i in range(6):
        h0 = dict(eng.host_seconds)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        eng.update(2.5e-4)
        b.record()
        torch.cuda.synchronize()
        rows.append({"device_ms": round(a.elapsed_time(b), 2),
                     **{k + "_ms": round(1e3 * (eng.host_seconds[k] - h0[k]), 2) for k in h0}})
    out.setdefault(mode, []).append(rows)
print(json.dumps(out))
