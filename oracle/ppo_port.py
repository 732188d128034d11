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
        max_grad_norm=0.5, anneal_lr=True, total_iterations=None, threads=None, log=None, device="cpu",
        on_update=None):
    """Run ``num_iterations`` PPO iterations; returns dict(per-iteration losses, sps, seconds).

    device="cpu" is the CPU baseline; device="cuda" reproduces what the reference does with --cuda (fp32 rollout
    storage on the device, eager torch/cuDNN ops, per-step .cpu() sync) for tools/ref_eager_gpu.py.

    ``on_update(ctx)`` (tests only) is called for every minibatch update after ``loss.backward()`` and BEFORE
    ``clip_grad_norm_`` / ``optimizer.step()`` with the live objects (agent, optimizer, rollout buffers, minibatch
    indices, loss scalars), so a test can replay each update on identical inputs through the CUDA path."""
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
                if on_update is not None:
                    on_update(dict(iteration=it, epoch=epoch, start=s, mb=mb, agent=agent, opt=opt, obs=obs, actions=actions,
                                   logprobs=logprobs, advantages=adv, returns=returns, values=values,
                                   losses=dict(pg_loss=pg.item(), v_loss=vl.item(), entropy=el.item(), loss=loss.item(),
                                               approx_kl=((ratio - 1) - logratio).mean().item(),
                                               old_approx_kl=(-logratio).mean().item(),
                                               clipfrac=((ratio - 1.0).abs() > clip_coef).float().mean().item()),
                                   lr=opt.param_groups[0]["lr"]))
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


def run_sliced(num_envs=1024, num_steps=128, slices_per_iteration=16, n_slices=4, warmup_slices=1, seed=1,
               num_minibatches=4, learning_rate=2.5e-4, gamma=0.99, gae_lambda=0.95, clip_coef=0.1, ent_coef=0.01,
               vf_coef=0.5, max_grad_norm=0.5, threads=None, log=None):
    """Bounded sample of the reference iteration AT ITS OWN SHAPES (bench.py's reference arm / cpu_baseline).

    One PPO iteration at num_envs=1024, num_steps=128 takes minutes on host cores, so the timed unit is one SLICE =
    1/16 of an iteration with every tensor at full size: ``num_steps/16`` policy steps over all ``num_envs`` envs
    (forward, Categorical sample, env step, buffer stores), then ONE minibatch update of the full minibatch size
    ``num_envs*num_steps/num_minibatches`` drawn from the [num_steps, num_envs] rollout buffer (gather, forward,
    loss, autograd backward, clip_grad_norm_, Adam), and the GAE loop once every 16th slice.  Sixteen slices do
    exactly the work of one reference iteration with 4 epochs x 4 minibatches (ppo_atari_envpool.py:224-325); a
    slice advances ``num_envs * num_steps / 16`` env steps.  The buffer starts from a rollout of random frames."""
    import random
    if threads:
        torch.set_num_threads(threads)
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    envs = SyntheticAtariVec(num_envs, seed=seed, mode="pool")
    agent = RefAgent(envs.action_space.n)
    opt = torch.optim.Adam(agent.parameters(), lr=learning_rate, eps=1e-5)
    T, N = num_steps, num_envs
    B = T * N
    M = B // num_minibatches
    S = slices_per_iteration
    assert T % S == 0
    tps = T // S
    obs = torch.randint(0, 256, (T, N, 4, 84, 84), dtype=torch.uint8).float()
    actions = torch.randint(0, envs.action_space.n, (T, N)).float()
    logprobs = torch.full((T, N), -1.386); rewards = torch.zeros((T, N)); dones = torch.zeros((T, N))
    values = torch.randn((T, N)) * 0.1
    adv = torch.randn((T, N)); returns = adv + values
    next_obs = torch.Tensor(envs.reset()); next_done = torch.zeros(N)
    inds = np.arange(B); np.random.shuffle(inds)
    out = {"slice_seconds": []}
    for sl in range(warmup_slices + n_slices):
        t0 = time.time()
        base = (sl % S) * tps
        for step in range(base, base + tps):
            obs[step] = next_obs; dones[step] = next_done
            with torch.no_grad():
                a, lp, _, v = agent.get_action_and_value(next_obs)
                values[step] = v.flatten()
            actions[step] = a; logprobs[step] = lp
            o, r, d, info = envs.step(a.cpu().numpy())
            rewards[step] = torch.tensor(r).view(-1)
            next_obs, next_done = torch.Tensor(o), torch.Tensor(d)
        if sl % S == S - 1:
            with torch.no_grad():
                next_value = agent.get_value(next_obs).reshape(1, -1)
                adv = torch.zeros_like(rewards); last = 0
                for t in reversed(range(T)):
                    nnt = 1.0 - (next_done if t == T - 1 else dones[t + 1])
                    nv = next_value if t == T - 1 else values[t + 1]
                    delta = rewards[t] + gamma * nv * nnt - values[t]
                    adv[t] = last = delta + gamma * gae_lambda * nnt * last
                returns = adv + values
            np.random.shuffle(inds)
        b_obs = obs.reshape((-1, 4, 84, 84)); b_lp = logprobs.reshape(-1); b_act = actions.reshape(-1)
        b_adv = adv.reshape(-1); b_ret = returns.reshape(-1); b_val = values.reshape(-1)
        s0 = (sl % num_minibatches) * M
        mb = inds[s0:s0 + M]
        _, nlp, ent, nv = agent.get_action_and_value(b_obs[mb], b_act.long()[mb])
        logratio = nlp - b_lp[mb]; ratio = logratio.exp()
        ma = b_adv[mb]; ma = (ma - ma.mean()) / (ma.std() + 1e-8)
        pg = torch.max(-ma * ratio, -ma * torch.clamp(ratio, 1 - clip_coef, 1 + clip_coef)).mean()
        nv = nv.view(-1)
        vu = (nv - b_ret[mb]) ** 2
        vc = (b_val[mb] + torch.clamp(nv - b_val[mb], -clip_coef, clip_coef) - b_ret[mb]) ** 2
        vl = 0.5 * torch.max(vu, vc).mean()
        loss = pg - ent_coef * ent.mean() + vl * vf_coef
        opt.zero_grad(); loss.backward()
        nn.utils.clip_grad_norm_(agent.parameters(), max_grad_norm)
        opt.step()
        dt = time.time() - t0
        if sl >= warmup_slices:
            out["slice_seconds"].append(dt)
        if log:
            log(f"[cpu port] slice {sl}: {N * tps / dt:.0f} SPS")
    out["env_steps_per_slice"] = N * tps
    out["minibatch_size"] = M
    out["loss_last"] = float(loss.detach())
    return out
