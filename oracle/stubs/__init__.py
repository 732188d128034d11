"""Fake ``gym`` / ``envpool`` / ``gymnasium`` modules (TEST INFRASTRUCTURE ONLY).

The reference scripts import env libraries that are absent from this image.
``install()`` puts ~100 lines of stand-ins into ``sys.modules`` so that the
UNMODIFIED reference scripts under /root/reference can be executed with
``runpy`` on CPU.  The vector envs they hand out are the
deterministic synthetic envs from ``cleanrl_b200.synthetic_envs``.
"""
from __future__ import annotations

import sys
import types

import numpy as np

from cleanrl_b200 import synthetic_envs as S

# knobs the harness can set before the script runs
CONFIG = {"atari_mode": "fresh", "gymnasium_kind": "discrete", "seed_override": None}
LAST_ENVS = []


class _Wrapper:
    """gym.Wrapper stand-in: forwards unknown attributes to the wrapped env."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def __class_getitem__(cls, item):
        return cls

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)

    def close(self):
        return self.env.close()


def _spaces_module(name):
    m = types.ModuleType(name)
    m.Discrete = S.Discrete
    m.Box = S.Box
    return m


def _envpool_make(env_id, env_type="gym", num_envs=1, episodic_life=True, reward_clip=True, seed=0, **kw):
    assert env_type == "gym"
    env = S.SyntheticAtariVec(num_envs, seed=seed, mode=CONFIG["atari_mode"])
    LAST_ENVS.append(env)
    return env


class _SyncVectorEnv:
    """gymnasium.vector.SyncVectorEnv stand-in: ignores the thunks' envs and
    returns one synthetic vector env with the same count."""

    def __new__(cls, env_fns, **kw):
        if CONFIG["gymnasium_kind"] == "cartpole":
            env = S.CartPoleVec(len(env_fns))
        else:
            env = S.SyntheticGymnasiumVec(len(env_fns), kind=CONFIG["gymnasium_kind"])
        LAST_ENVS.append(env)
        return env


class _DictObsProcgen:
    """procgen.ProcgenEnv stand-in: observations are dicts {"rgb": uint8 [N, 64, 64, 3]} until the script's own
    gym.wrappers.TransformObservation(envs, lambda obs: obs["rgb"]) unwraps them (cleanrl/ppo_procgen.py:176-178)."""

    def __init__(self, num_envs, env_name="starpilot", num_levels=0, start_level=0, distribution_mode="easy", **kw):
        self.env = S.SyntheticProcgenVec(num_envs, seed=CONFIG.get("procgen_seed", 0))
        self.num_envs = num_envs
        self.action_space = self.env.action_space
        self.observation_space = {"rgb": self.env.observation_space}
        LAST_ENVS.append(self.env)

    def reset(self, **kw):
        return {"rgb": self.env.reset()}

    def step(self, action):
        o, r, d, info = self.env.step(action)
        return {"rgb": o}, r, d, info

    def close(self):
        pass


class _TransformObservation(_Wrapper):
    def __init__(self, env, f):
        super().__init__(env)
        self.f = f

    def reset(self, **kwargs):
        return self.f(self.env.reset(**kwargs))

    def step(self, action):
        o, r, d, info = self.env.step(action)
        return self.f(o), r, d, info


class _PassThrough(_Wrapper):
    def __init__(self, env, *a, **k):
        super().__init__(env)


def install():
    gym = types.ModuleType("gym")
    gym.Wrapper = _Wrapper
    gym.spaces = _spaces_module("gym.spaces")
    gym.wrappers = types.ModuleType("gym.wrappers")
    gym.wrappers.TransformObservation = _TransformObservation
    for w in ("RecordEpisodeStatistics", "RecordVideo", "NormalizeReward", "TransformReward"):
        setattr(gym.wrappers, w, _PassThrough)      # reward normalisation is host-side env code, identity in the fixtures
    procgen = types.ModuleType("procgen")
    procgen.ProcgenEnv = _DictObsProcgen
    envpool = types.ModuleType("envpool")
    envpool.make = _envpool_make

    gymn = types.ModuleType("gymnasium")
    gymn.Wrapper = _Wrapper
    gymn.Env = object
    gymn.spaces = _spaces_module("gymnasium.spaces")
    gymn.vector = types.ModuleType("gymnasium.vector")
    gymn.vector.SyncVectorEnv = _SyncVectorEnv
    gymn.wrappers = types.ModuleType("gymnasium.wrappers")
    for w in ("RecordVideo", "RecordEpisodeStatistics", "ResizeObservation", "GrayScaleObservation",
              "FrameStack", "FlattenObservation", "ClipAction", "NormalizeObservation",
              "TransformObservation", "NormalizeReward", "TransformReward"):
        setattr(gymn.wrappers, w, _Wrapper)
    gymn.make = lambda *a, **k: None
    gymn.ObservationWrapper = _Wrapper
    gymn.RewardWrapper = _Wrapper
    gymn.ActionWrapper = _Wrapper

    mods = {
        "gym": gym, "gym.spaces": gym.spaces, "gym.wrappers": gym.wrappers, "procgen": procgen, "envpool": envpool,
        "gymnasium": gymn, "gymnasium.spaces": gymn.spaces,
        "gymnasium.vector": gymn.vector, "gymnasium.wrappers": gymn.wrappers,
    }
    for k, v in mods.items():
        sys.modules[k] = v
    return mods


def uninstall():
    for k in ("gym", "gym.spaces", "gym.wrappers", "procgen", "envpool", "gymnasium", "gymnasium.spaces",
              "gymnasium.vector", "gymnasium.wrappers"):
        sys.modules.pop(k, None)
