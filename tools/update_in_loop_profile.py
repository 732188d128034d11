"""Per-launch kernel times of the 16 minibatch updates of ONE bench iteration (device-resident rollout), one profile window
per minibatch: shows whether a kernel's time in the real loop differs from tools/profile_update.py's isolated minibatch.

    python tools/update_in_loop_profile.py [kernel name, default fc_dgrad]"""
import ctypes
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import ppo_args  # noqa: E402
from cleanrl_b200 import _lib, build  # noqa: E402
from cleanrl_b200.agents import NatureCNNAgent  # noqa: E402
from cleanrl_b200.ppo_engine import PPOEngine  # noqa: E402
from cleanrl_b200.synthetic_envs import SyntheticAtariVec  # noqa: E402

build.build()
lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "fc_dgrad"
N, T = 1024, 128
dev = torch.device("cuda:0")
args = ppo_args(N, T, 8, "bf16")
np.random.seed(1); torch.manual_seed(1)
envs = SyntheticAtariVec(N, seed=1, mode="pool", pinned=True)
envs.single_observation_space, envs.single_action_space = envs.observation_space, envs.action_space
agent = NatureCNNAgent(envs).to(dev)
agent.precision = "bf16"
eng = PPOEngine(agent, args, (4, 84, 84), np.uint8, N, dev, gae_mode=1)
pool_dev = torch.from_numpy(envs._batches).to(dev)
g = torch.Generator().manual_seed(1)
eng_rewards = torch.randint(0, 2, (T, N), generator=g).float().to(dev)
eng_dones = (torch.rand(T, N, generator=g) < 0.02).float().to(dev)


def rollout():
    eng.rollout_resident(pool_dev)
    eng.rewards.copy_(eng_rewards); eng.dones.copy_(eng_dones)
    eng._to_storage(pool_dev[T % pool_dev.shape[0]], None)
    eng.finish_rollout(None, None, resident=True)


for _ in range(2):
    rollout(); eng.update(2.5e-4)
rollout()
rows = []
orig = eng.minibatch_update
buf = ctypes.create_string_buffer(1 << 16)


def profiled(mb_inds, lr, k=0):
    lib.b200rl_profile_reset(); lib.b200rl_profile_enable(1)
    orig(mb_inds, lr, k)
    lib.b200rl_profile_enable(0)
    lib.b200rl_profile_summary(buf, 1 << 16)
    r = {x["name"]: round(1e3 * x["ms"] / x["launches"], 1) for x in json.loads(buf.value.decode())}
    rows.append(r)


eng.update_graphs = False
eng.minibatch_update = profiled
eng.update(2.5e-4)
names = sorted(rows[0], key=lambda n: -rows[0][n])
print(json.dumps({"kernel": which, "us_per_minibatch": [r.get(which) for r in rows],
                  "all_kernels_us_minibatch0": {n: rows[0][n] for n in names},
                  "all_kernels_us_minibatch9": {n: rows[9][n] for n in names}}))
