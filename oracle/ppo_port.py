"""CPU port of the reference PPO-Atari iteration (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

A compact torch-CPU restatement of cleanrl/ppo_atari_envpool.py:199-341 (same ops: fp32 NCHW
conv/linear through torch, autograd backward, clip_grad_norm_, Adam eps=1e-5) used where the
reference itself cannot be run because /root/reference does not exist (the GPU box):
``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference``.  It is validated against the
unmodified reference script by tests/test_oracle_port.py (identical losses on the same seed).
Never imported by the product.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn
from torch.distributions.categorical import Categorical

from cleanrl_b200.synthetic_envs import SyntheticAtariVec


def _init(layer, std=np.sqrt(2), bias_const=0.0):
    torch.nn.init.orthogonal_(layer.weight, std)
    torch.nn.init.constant_(layer.bias, bias_const)
    return layer


class RefAgent(nn.Module):
    """Same module tree / init order as the reference Agent (ppo_atari_envpool.py:123-139)."""

    def __init__(self, n_actions):
        super().__init__()
        self.network = nn.Sequential(
            _init(nn.Conv2d(4, 32, 8, stride=4)), nn.ReLU(), _init(nn.Conv2d(32, 64, 4, stride=2)), nn.ReLU(),
            _init(nn.Conv2d(64, 64, 3, stride=1)), nn.ReLU(), nn.Flatten(), _init(nn.Linear(64 * 7 * 7, 512)), nn.ReLU())
        self.actor = _init(nn.Linear(512, n_actions), std=0.01)
        self.critic = _init(nn.Linear(512, 1), std=1)

    def get_value(self, x):
        return self.critic(self.network(x / 255.0))

    def get_action_and_value(self, x, action=None):
        hidden = self.network(x / 255.0)
        probs = Categorical(logits=self.actor(hidden))
        if action is None:
            action = probs.sample()
        return action, probs.log_prob(action), probs.entropy(), self.critic(hidden)


def run(num_envs=8, num_steps=32, num_iterations=2, seed=1, env_mode="fresh", num_minibatches=4, update_epochs=4,
        learning_rate=2.5e-4, gamma=0.99, gae_lambda=0.95, clip_coef=0.1, ent_coef=0.01, vf_coef=0.5,
        max_grad_norm=0.5, anneal_lr=True, total_iterations=None, threads=None, log=None, device="cpu"):
    """Run ``num_iterations`` PPO iterations; returns dict(per-iteration losses, sps, seconds).

    device="cpu" is the CPU baseline; device="cuda" reproduces what the reference does with --cuda (fp32 rollout
    storage on the device, eager torch/cuDNN ops, per-step .cpu() sync) for tools/ref_eager_gpu.py."""
    device = torch.device(device)
    import random
    if threads:
        torch.set_num_threads(threads)
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    envs = SyntheticAtariVec(num_envs, seed=seed, mode=env_mode)
    agent = RefAgent(envs.action_space.n).to(device)
    opt = torch.optim.Adam(agent.parameters(), lr=learning_rate, eps=1e-5)
    T, N = num_steps, num_envs
    B = T * N
    M = B // num_minibatches
    total_iterations = total_iterations or num_iterations
    obs = torch.zeros((T, N, 4, 84, 84)).to(device); actions = torch.zeros((T, N)).to(device)
    logprobs = torch.zeros((T, N)).to(device); rewards = torch.zeros((T, N)).to(device)
    dones = torch.zeros((T, N)).to(device); values = torch.zeros((T, N)).to(device)
    next_obs = torch.Tensor(envs.reset()).to(device); next_done = torch.zeros(N).to(device)
    out = {"losses": [], "iter_seconds": []}
    t_start = time.time()
    for it in range(1, num_iterations + 1):
        t0 = time.time()
        if anneal_lr:
            opt.param_groups[0]["lr"] = (1.0 - (it - 1.0) / total_iterations) * learning_rate
        for step in range(T):
            obs[step] = next_obs; dones[step] = next_done
            with torch.no_grad():
                a, lp, _, v = agent.get_action_and_value(next_obs)
                values[step] = v.flatten()
            actions[step] = a; logprobs[step] = lp
            o, r, d, info = envs.step(a.cpu().numpy())
            rewards[step] = torch.tensor(r).to(device).view(-1)
            next_obs, next_done = torch.Tensor(o).to(device), torch.Tensor(d).to(device)
        with torch.no_grad():
            next_value = agent.get_value(next_obs).reshape(1, -1)
            adv = torch.zeros_like(rewards); last = 0
            for t in reversed(range(T)):
                nnt = 1.0 - (next_done if t == T - 1 else dones[t + 1])
                nv = next_value if t == T - 1 else values[t + 1]
                delta = rewards[t] + gamma * nv * nnt - values[t]
                adv[t] = last = delta + gamma * gae_lambda * nnt * last
            returns = adv + values
        b_obs = obs.reshape((-1, 4, 84, 84)); b_lp = logprobs.reshape(-1); b_act = actions.reshape(-1)
        b_adv = adv.reshape(-1); b_ret = returns.reshape(-1); b_val = values.reshape(-1)
        inds = np.arange(B)
        for epoch in range(update_epochs):
            np.random.shuffle(inds)
            for s in range(0, B, M):
                mb = inds[s:s + M]
                _, nlp, ent, nv = agent.get_action_and_value(b_obs[mb], b_act.long()[mb])
                logratio = nlp - b_lp[mb]; ratio = logratio.exp()
                ma = b_adv[mb]; ma = (ma - ma.mean()) / (ma.std() + 1e-8)
                pg = torch.max(-ma * ratio, -ma * torch.clamp(ratio, 1 - clip_coef, 1 + clip_coef)).mean()
                nv = nv.view(-1)
                vu = (nv - b_ret[mb]) ** 2
                vc = (b_val[mb] + torch.clamp(nv - b_val[mb], -clip_coef, clip_coef) - b_ret[mb]) ** 2
                vl = 0.5 * torch.max(vu, vc).mean()
                el = ent.mean()
                loss = pg - ent_coef * el + vl * vf_coef
                opt.zero_grad(); loss.backward()
                nn.utils.clip_grad_norm_(agent.parameters(), max_grad_norm)
                opt.step()
        out["losses"].append(dict(pg_loss=pg.item(), v_loss=vl.item(), entropy=el.item()))
        if device.type == "cuda":
            torch.cuda.synchronize()
        out["iter_seconds"].append(time.time() - t0)
        if log:
            log(f"[cpu port] iteration {it}: {B / out['iter_seconds'][-1]:.0f} SPS")
    out["seconds"] = time.time() - t_start
    out["env_steps"] = num_iterations * B
    return out
