"""Evaluation helper with the surface of cleanrl_utils/evals/ppo_eval.py:7-36:
``evaluate(model_path, make_env, env_id, eval_episodes, run_name, Model, device, capture_video, gamma)``
builds one env, ``Model(envs).to(device)``, ``load_state_dict(torch.load(model_path))`` and rolls episodes with
``agent.get_action_and_value`` -- here executed by libb200rl kernels.  ``.cleanrl_model`` files interchange with
the reference because the ``state_dict`` keys are the same.
"""
from __future__ import annotations

import numpy as np
import torch


def evaluate(model_path, make_env, env_id, eval_episodes, run_name, Model, device=torch.device("cuda"),
             capture_video=True, gamma=0.99, envs=None, max_steps=100000):
    if envs is None:
        import gymnasium as gym  # type: ignore

        envs = gym.vector.SyncVectorEnv([make_env(env_id, 0, capture_video, run_name, gamma)])
    agent = Model(envs).to(device)
    agent.load_state_dict(torch.load(model_path, map_location=device))
    agent.eval()
    obs, _ = envs.reset()
    episodic_returns = []
    steps = 0
    while len(episodic_returns) < eval_episodes and steps < max_steps:
        actions, _, _, _ = agent.get_action_and_value(torch.as_tensor(np.asarray(obs)).to(device))
        next_obs, _, _, _, infos = envs.step(actions.cpu().numpy())
        if "final_info" in infos:
            for info in infos["final_info"]:
                if not info or "episode" not in info:
                    continue
                print(f"eval_episode={len(episodic_returns)}, episodic_return={info['episode']['r']}")
                episodic_returns += [info["episode"]["r"]]
        obs = next_obs
        steps += 1
    return episodic_returns
