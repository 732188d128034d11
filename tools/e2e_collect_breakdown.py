"""Where one end-to-end rollout step of the grouped loop (PPOEngine.collect, 2 env groups) spends its HOST wall time, for
the bench.py e2e configuration with the frame-stack delta upload on or off.

    python tools/e2e_collect_breakdown.py [num_envs] [num_steps] [stack|pool] [delta 0|1]

Per group step: wait (actions D2H event + verification join), env.step, launch (tracker begin + H2D enqueue + graph
launches).  Measurement tool only."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
mode = sys.argv[3] if len(sys.argv) > 3 else "stack"
os.environ["CLEANRL_B200_DELTA_UPLOAD"] = sys.argv[4] if len(sys.argv) > 4 else "1"
from bench import ppo_args  # noqa: E402
from cleanrl_b200 import ppo_engine as pe  # noqa: E402
from cleanrl_b200.agents import NatureCNNAgent  # noqa: E402
from cleanrl_b200.synthetic_envs import SyntheticAtariVec  # noqa: E402

dev = torch.device("cuda:0")
args = ppo_args(N, T, 4, "bf16")
spaces = SyntheticAtariVec(2, seed=1)
spaces.single_observation_space, spaces.single_action_space = spaces.observation_space, spaces.action_space
agent = NatureCNNAgent(spaces).to(dev)
agent.precision = "bf16"
eng = pe.PPOEngine(agent, args, (4, 84, 84), np.uint8, N, dev, gae_mode=1)
G = 2
parts = [SyntheticAtariVec(N // G, seed=1 + g, mode=mode, pinned=True) for g in range(G)]
acc = {"wait": 0.0, "env": 0.0, "launch": 0.0, "begin": 0.0}
on = [False]


def wrap(obj, name, key):
    f = getattr(obj, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        if on[0]:
            acc[key] += time.perf_counter() - t0
        return r
    setattr(obj, name, g)


wrap(eng, "wait_actions", "wait")
wrap(eng, "launch_part", "launch")
for e in parts:
    wrap(e, "step", "env")
obs_p = [e.reset() for e in parts]
done_p = [np.zeros(N // G, dtype=np.float32) for _ in parts]
res = {}
for it in range(4):
    on[0] = it == 3
    if on[0] and eng._delta is not None:
        for d in eng._delta:
            wrap(d["tr"], "begin", "begin")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    obs_p, done_p = eng.collect(parts, obs_p, done_p)
    eng.finish_rollout_parts(obs_p, done_p)
    torch.cuda.synchronize()
    roll = time.perf_counter() - t0
    eng.update(2.5e-4)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
print(json.dumps({"num_envs": N, "num_steps": T, "obs": mode, "delta_upload": bool(eng.delta_upload), "redos": eng.delta_redos,
                  "host_threads": os.environ.get("CLEANRL_B200_HOST_THREADS", "default"),
                  "rollout_ms": round(roll * 1e3, 2), "iteration_ms": round(tot * 1e3, 2),
                  "per_env_step_us": {k: round(v / T * 1e6, 1) for k, v in acc.items()},
                  "note": "per_env_step_us sums both groups; 'begin' (tracker classify + stage) is part of 'launch'"}))
