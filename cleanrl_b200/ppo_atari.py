"""Drop-in for cleanrl/ppo_atari.py (single-process PPO on gymnasium Atari, NatureCNN Agent) on libb200rl kernels.

Same flags (cleanrl/ppo_atari.py:20-90), ``make_env`` wrapper stack (:93-117), Agent surface
(``network`` / ``actor`` / ``critic`` => same state_dict keys, :126-152), TensorBoard tags and stdout as the
reference.  The loop is the gymnasium-API loop of ppo.py around the NatureCNN agent; frames stay uint8 from the
env to the GPU (the reference's ``x / 255.0`` is folded into the first convolution's epilogue), the rollout is
stored once as space-to-depth bf16 when ``--precision bf16``.
"""
from __future__ import annotations

import os
import sys

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from cleanrl_b200 import cli
from cleanrl_b200 import ppo as _loop
from cleanrl_b200.agents import NatureCNNAgent as Agent, layer_init  # noqa: F401

Args = cli.ppo_atari_args(os.path.basename(__file__)[: -len(".py")])
run_name = None


def make_env(env_id, idx, capture_video, run_name):
    """The reference's wrapped Atari env (ppo_atari.py:93-117); only used when gymnasium + ALE are installed."""
    def thunk():
        import gymnasium as gym  # type: ignore
        from cleanrl_utils.atari_wrappers import (ClipRewardEnv, EpisodicLifeEnv, FireResetEnv,  # type: ignore
                                                  MaxAndSkipEnv, NoopResetEnv)

        if capture_video and idx == 0:
            env = gym.wrappers.RecordVideo(gym.make(env_id, render_mode="rgb_array"), f"videos/{run_name}")
        else:
            env = gym.make(env_id)
        env = gym.wrappers.RecordEpisodeStatistics(env)
        env = MaxAndSkipEnv(NoopResetEnv(env, noop_max=30), skip=4)
        env = EpisodicLifeEnv(env)
        if "FIRE" in env.unwrapped.get_action_meanings():
            env = FireResetEnv(env)
        env = gym.wrappers.ResizeObservation(ClipRewardEnv(env), (84, 84))
        return gym.wrappers.FrameStack(gym.wrappers.GrayScaleObservation(env), 4)

    return thunk


def make_envs(args, run_name):
    if not cli.use_synthetic(args):
        try:
            import gymnasium as gym  # type: ignore  # noqa: F401
        except ImportError as e:
            raise cli.env_import_error("gymnasium (+ ale-py, cleanrl_utils.atari_wrappers)", e) from e
        return gym.vector.SyncVectorEnv([make_env(args.env_id, i, args.capture_video, run_name)
                                         for i in range(args.num_envs)])
    from cleanrl_b200.synthetic_envs import SyntheticGymnasiumVec

    return SyntheticGymnasiumVec(args.num_envs, kind="atari")


def main(argv=None, writer_factory=None, env_factory=None, on_iteration=None, agent_hook=None):
    global run_name
    args = cli.parse(Args, argv)
    engine, run_name = _loop.run(args, Agent, make_envs, np.uint8, "cleanrl_b200.ppo_atari", writer_factory,
                                 env_factory, on_iteration, agent_hook)
    return engine


if __name__ == "__main__":
    main()
