"""Post-training evaluation with the call surfaces of cleanrl_utils/evals/ppo_eval.py:7-36 and dqn_eval.py:9-44
(SURVEY.md 8f rank 1).

``evaluate(model_path, make_env, env_id, eval_episodes, run_name, Model, device, capture_video, gamma)`` rebuilds
the agent from a ``.cleanrl_model`` file (a plain ``state_dict`` whose keys equal the reference's, so files written
by either implementation load in both) and plays until ``eval_episodes`` episode returns have been collected.
The policy runs on libb200rl kernels through ``Model.get_action_and_value``.
"""
from __future__ import annotations

import numpy as np
import torch


def _single_env(make_env, env_id, capture_video, run_name, gamma):
    import gymnasium as gym  # type: ignore

    return gym.vector.SyncVectorEnv([make_env(env_id, 0, capture_video, run_name, gamma)])


def _finished_returns(infos):
    """Episode returns reported by gymnasium's RecordEpisodeStatistics in this step (possibly none)."""
    out = []
    for info in infos.get("final_info", ()):
        if info and "episode" in info:
            out.append(info["episode"]["r"])
    return out


def evaluate(model_path, make_env, env_id, eval_episodes, run_name, Model, device=torch.device("cuda"),
             capture_video=True, gamma=0.99, envs=None, max_steps=100000):
    envs = envs if envs is not None else _single_env(make_env, env_id, capture_video, run_name, gamma)
    agent = Model(envs).to(device)
    agent.load_state_dict(torch.load(model_path, map_location=device))
    agent.eval()

    episodic_returns = []
    obs, _ = envs.reset()
    for _ in range(max_steps):
        if len(episodic_returns) >= eval_episodes:
            break
        with torch.no_grad():
            action = agent.get_action_and_value(torch.as_tensor(np.asarray(obs)).to(device))[0]
        obs, _, _, _, infos = envs.step(action.cpu().numpy())
        for ret in _finished_returns(infos):
            print(f"eval_episode={len(episodic_returns)}, episodic_return={ret}")
            episodic_returns.append(ret)
    return episodic_returns


def evaluate_q(model_path, make_env, env_id, eval_episodes, run_name, Model, device=torch.device("cuda"),
               epsilon=0.05, capture_video=True, envs=None, max_steps=1000000):
    """Epsilon-greedy rollout of a saved Q-network (cleanrl_utils/evals/dqn_eval.py): ``Model(envs)`` is rebuilt from
    the ``.cleanrl_model`` state_dict, greedy actions come from ``Model.forward`` (libb200rl kernels + argmax), with
    probability ``epsilon`` every env takes a uniformly sampled action instead (python's ``random`` as the reference)."""
    import random

    if envs is None:
        import gymnasium as gym  # type: ignore

        envs = gym.vector.SyncVectorEnv([make_env(env_id, 0, 0, capture_video, run_name)])
    model = Model(envs).to(device)
    model.load_state_dict(torch.load(model_path, map_location=device))
    model.eval()

    returns = []
    obs, _ = envs.reset()
    for _ in range(max_steps):
        if len(returns) >= eval_episodes:
            break
        if random.random() < epsilon:
            actions = np.array([envs.single_action_space.sample() for _ in range(envs.num_envs)])
        else:
            with torch.no_grad():
                q = model(torch.as_tensor(np.asarray(obs)).to(device))
            actions = q.argmax(dim=1).cpu().numpy()
        obs, _, _, _, infos = envs.step(actions)
        for ret in _finished_returns(infos):
            print(f"eval_episode={len(returns)}, episodic_return={ret}")
            returns.append(ret)
    return returns
