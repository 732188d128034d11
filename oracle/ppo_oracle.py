"""CPU restatement of the reference PPO hot path, one pure function per kernel.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy fp32 arithmetic with
the reference's operation order; each function cites the reference lines it
restates.  Third-party arithmetic on this path lives in torch (pinned
``torch==2.4.1`` in the reference's pyproject.toml:19; 2.11.0 installed here):
Categorical / multinomial, clip_grad_norm_, Adam -- restated from the installed
sources.  Pinned by tests/test_oracle_*.py against (i) the unmodified reference
script run through oracle/ref_harness.py (fixtures in tests/golden/) and
(ii) torch autograd / torch.optim.Adam directly.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------- GAE
def gae(rewards, values, dones, next_value, next_done, gamma, gae_lambda):
    """Reverse-scan GAE, cleanrl/ppo.py:218-231 (== ppo_atari_envpool.py:250-263).

    Every ``*``/``+``/``-`` is rounded separately to fp32; ``gamma`` becomes fp32
    once, ``gamma * gae_lambda`` is a double product rounded once (ppo.py:230).
    rewards/values/dones: f32 [T, N]; next_value, next_done: f32 [N].
    Returns (advantages, returns) f32 [T, N].
    """
    rewards = np.asarray(rewards, dtype=f32)
    values = np.asarray(values, dtype=f32)
    dones = np.asarray(dones, dtype=f32)
    T, N = rewards.shape
    g = f32(gamma)
    gl = f32(float(gamma) * float(gae_lambda))
    adv = np.zeros((T, N), dtype=f32)
    last = np.zeros(N, dtype=f32)
    for t in reversed(range(T)):
        if t == T - 1:
            nnt = f32(1.0) - np.asarray(next_done, dtype=f32).reshape(N)
            nv = np.asarray(next_value, dtype=f32).reshape(N)
        else:
            nnt = f32(1.0) - dones[t + 1]
            nv = values[t + 1]
        delta = (rewards[t] + (g * nv) * nnt) - values[t]
        last = delta + ((gl * nnt) * last)
        adv[t] = last
    return adv, adv + values


# ------------------------------------------------------------------- Categorical
def _logsumexp(x):
    """ATen logsumexp: amax, exp(x - max), sum, log, + max."""
    m = x.max(axis=-1, keepdims=True)
    m = np.where(np.isinf(m), f32(0), m)
    return np.log(np.exp(x - m, dtype=f32).sum(axis=-1, keepdims=True, dtype=f32)) + m


def categorical(logits):
    """``Categorical(logits=...)`` ctor + lazy ``probs`` (torch/distributions/categorical.py):
    normalised logits = logits - logsumexp; probs = softmax(normalised logits)."""
    logits = np.asarray(logits, dtype=f32)
    nl = logits - _logsumexp(logits)
    m = nl.max(axis=-1, keepdims=True)
    e = np.exp(nl - m, dtype=f32)
    p = e / e.sum(axis=-1, keepdims=True, dtype=f32)
    return nl.astype(f32), p.astype(f32)


def categorical_sample(logits, noise):
    """``probs.sample()`` == ``torch.multinomial(probs, 1, True)`` ==
    ``argmax(probs / q)`` with ``q ~ Exp(1)`` drawn by the caller from torch's
    generator (ATen native/Distributions.cpp multinomial fast path), then
    ``log_prob`` (gather of normalised logits) and ``entropy`` (-sum p*logp,
    logits clamped at finfo.min).  Reference call site:
    cleanrl/ppo_atari_envpool.py:143-149.  Returns (action i64, logprob, entropy)."""
    nl, p = categorical(logits)
    action = np.argmax(p / np.asarray(noise, dtype=f32), axis=-1).astype(np.int64)
    return (action,) + categorical_eval(logits, action)


def categorical_eval(logits, action):
    nl, p = categorical(logits)
    logprob = np.take_along_axis(nl, np.asarray(action).reshape(-1, 1), axis=-1)[:, 0]
    cl = np.maximum(nl, np.finfo(f32).min)
    ent = -(cl * p).sum(axis=-1, dtype=f32)
    return logprob.astype(f32), ent.astype(f32)


# ---------------------------------------------------------------------- PPO loss
def ppo_loss(new_logits, new_value, mb_inds, b_actions, b_logprobs, b_advantages, b_returns, b_values,
             clip_coef, ent_coef, vf_coef, norm_adv=True, clip_vloss=True):
    """Minibatch PPO loss + analytic gradients wrt (logits, value).

    Restates cleanrl/ppo.py:251-285 (ratio, both KLs, clipfrac, per-minibatch
    advantage normalisation with the UNBIASED std, clipped surrogate, clipped
    value loss, entropy bonus).  Gradients follow torch autograd's rules for
    ``max`` (ties split evenly) and ``clamp`` (inclusive pass-through).
    Returns (stats dict, dlogits [M,A], dvalue [M]).
    """
    new_logits = np.asarray(new_logits, dtype=f32)
    M, A = new_logits.shape
    idx = np.arange(M) if mb_inds is None else np.asarray(mb_inds)
    a = np.asarray(b_actions)[idx].astype(np.int64)
    nl, p = categorical(new_logits)
    newlogprob = np.take_along_axis(nl, a.reshape(-1, 1), axis=-1)[:, 0]
    entropy = -(np.maximum(nl, np.finfo(f32).min) * p).sum(axis=-1, dtype=f32)
    logratio = newlogprob - np.asarray(b_logprobs, dtype=f32)[idx]
    ratio = np.exp(logratio, dtype=f32)
    old_approx_kl = (-logratio).mean(dtype=f32)
    approx_kl = ((ratio - f32(1)) - logratio).mean(dtype=f32)
    clipfrac = (np.abs(ratio - f32(1.0)) > f32(clip_coef)).astype(f32).mean(dtype=f32)
    adv = np.asarray(b_advantages, dtype=f32)[idx]
    adv_mean = f32(0)
    adv_std = f32(1)
    if norm_adv:
        adv_mean = adv.mean(dtype=f32)
        adv_std = f32(np.sqrt(((adv - adv_mean) ** 2).sum(dtype=f32) / f32(M - 1)))
        adv = (adv - adv_mean) / (adv_std + f32(1e-8))
    lo, hi = f32(1 - clip_coef), f32(1 + clip_coef)
    rc = np.clip(ratio, lo, hi)
    pg1 = -adv * ratio
    pg2 = -adv * rc
    pg_loss = np.maximum(pg1, pg2).mean(dtype=f32)
    nv = np.asarray(new_value, dtype=f32).reshape(M)
    R = np.asarray(b_returns, dtype=f32)[idx]
    V = np.asarray(b_values, dtype=f32)[idx]
    c = f32(clip_coef)
    if clip_vloss:
        vu = (nv - R) ** 2
        d = nv - V
        vcl = V + np.clip(d, -c, c)
        vc = (vcl - R) ** 2
        v_loss = f32(0.5) * np.maximum(vu, vc).mean(dtype=f32)
    else:
        v_loss = f32(0.5) * ((nv - R) ** 2).mean(dtype=f32)
    ent_loss = entropy.mean(dtype=f32)
    loss = pg_loss - f32(ent_coef) * ent_loss + v_loss * f32(vf_coef)

    # ---- gradients
    invM = f32(1.0) / f32(M)
    inrange = ((ratio >= lo) & (ratio <= hi)).astype(f32)
    g_ratio = np.where(pg1 > pg2, -adv, f32(0)) + np.where(pg1 == pg2, f32(0.5) * (-adv) * (f32(1) + inrange), f32(0))
    g_lp = g_ratio * ratio * invM
    g_ent = -f32(ent_coef) * invM
    onehot = np.zeros((M, A), dtype=f32)
    onehot[np.arange(M), a] = 1
    dlogits = g_lp[:, None] * (onehot - p) + g_ent * (-p * (nl + entropy[:, None]))
    if clip_vloss:
        gu = f32(2) * (nv - R)
        gc = f32(2) * (vcl - R) * ((d >= -c) & (d <= c)).astype(f32)
        gv = np.where(vu > vc, gu, f32(0)) + np.where(vc > vu, gc, f32(0)) + np.where(vu == vc, f32(0.5) * (gu + gc), f32(0))
    else:
        gv = f32(2) * (nv - R)
    dvalue = f32(vf_coef) * f32(0.5) * invM * gv
    stats = dict(pg_loss=pg_loss, v_loss=v_loss, entropy=ent_loss, old_approx_kl=old_approx_kl,
                 approx_kl=approx_kl, clipfrac=clipfrac, loss=loss, adv_mean=adv_mean, adv_std=adv_std)
    return stats, dlogits.astype(f32), dvalue.astype(f32)


# -------------------------------------------------------------------- clip + Adam
def clip_adam(params, grads, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-5,
              max_norm=0.5, world_size=1):
    """``clip_grad_norm_`` then one Adam step on a FLAT fp32 vector.

    Restates cleanrl/ppo.py:288-290 (+ the DP averaging of
    ppo_atari_multigpu.py:360-374 when ``world_size > 1``: ``grads`` is then the
    SUM over ranks), with the arithmetic of the installed torch:
    torch/nn/utils/clip_grad.py:165-174 (coef = max_norm/(norm+1e-6) clamped to 1,
    always multiplied in) and torch/optim/adam.py:_single_tensor_adam
    (lerp_, mul_/addcmul_, sqrt/bias_correction2_sqrt + eps, addcdiv_).
    ``step`` is the 1-based step count AFTER increment.  ``max_norm=None`` skips
    clipping (dqn_atari.py has none).  Returns (params, exp_avg, exp_avg_sq, total_norm).
    """
    p = np.asarray(params, dtype=f32).copy()
    g = np.asarray(grads, dtype=f32).copy()
    m = np.asarray(exp_avg, dtype=f32).copy()
    v = np.asarray(exp_avg_sq, dtype=f32).copy()
    if world_size > 1:
        g = g / f32(world_size)
    total_norm = f32(np.sqrt((g.astype(np.float64) ** 2).sum()))
    if max_norm is not None:
        coef = min(f32(max_norm) / (total_norm + f32(1e-6)), f32(1.0))
        g = g * f32(coef)
    w = f32(1 - beta1)
    m = m + w * (g - m)
    v = v * f32(beta2)
    v = v + (f32(1 - beta2) * g) * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = bc2 ** 0.5
    denom = np.sqrt(v) / f32(bc2_sqrt) + f32(eps)
    p = p + f32(-step_size) * (m / denom)
    return p, m, v, total_norm


def anneal_lr(iteration, num_iterations, learning_rate):
    """cleanrl/ppo.py:187-190 (Python double arithmetic)."""
    frac = 1.0 - (iteration - 1.0) / num_iterations
    return frac * learning_rate


def explained_variance(values, returns):
    """cleanrl/ppo.py:295-297."""
    y_pred, y_true = np.asarray(values), np.asarray(returns)
    var_y = np.var(y_true)
    return np.nan if var_y == 0 else 1 - np.var(y_true - y_pred) / var_y


# ------------------------------------------------------------- diagonal Gaussian
_LOG_SQRT_2PI = f32(0.9189385332046727)   # math.log(math.sqrt(2 * math.pi)), torch/distributions/normal.py


def gaussian_eval(mean, logstd, action):
    """``Normal(mean, exp(logstd)).log_prob(a).sum(1)`` and ``.entropy().sum(1)``
    (cleanrl/ppo_continuous_action.py:134-141; torch/distributions/normal.py log_prob / entropy)."""
    mean = np.asarray(mean, dtype=f32)
    a = np.asarray(action, dtype=f32)
    std = np.exp(np.asarray(logstd, dtype=f32).reshape(1, -1))
    var = std * std
    ls = np.log(std)
    lp = (-((a - mean) ** 2) / (f32(2) * var) - ls - _LOG_SQRT_2PI).sum(axis=1, dtype=f32)
    ent = np.broadcast_to(f32(0.5) + _LOG_SQRT_2PI + ls, mean.shape).sum(axis=1, dtype=f32)
    return lp.astype(f32), ent.astype(f32)


def gaussian_sample(mean, logstd, noise):
    """``Normal.sample()`` == ``torch.normal(mean, std)`` == ``randn.mul_(std).add_(mean)`` with the N(0,1)
    draws supplied by the caller's generator."""
    mean = np.asarray(mean, dtype=f32)
    std = np.exp(np.asarray(logstd, dtype=f32).reshape(1, -1))
    a = (np.asarray(noise, dtype=f32) * std + mean).astype(f32)
    return (a,) + gaussian_eval(mean, logstd, a)


def ppo_loss_gaussian(new_mean, logstd, new_value, mb_inds, b_actions, b_logprobs, b_advantages, b_returns, b_values,
                      clip_coef, ent_coef, vf_coef, norm_adv=True, clip_vloss=True):
    """Continuous-action twin of ``ppo_loss`` (cleanrl/ppo_continuous_action.py:262-300).
    Returns (stats, dmean [M,D], dlogstd [D], dvalue [M])."""
    new_mean = np.asarray(new_mean, dtype=f32)
    M, D = new_mean.shape
    idx = np.arange(M) if mb_inds is None else np.asarray(mb_inds)
    a = np.asarray(b_actions, dtype=f32).reshape(-1, D)[idx]
    newlogprob, entropy = gaussian_eval(new_mean, logstd, a)
    logratio = newlogprob - np.asarray(b_logprobs, dtype=f32)[idx]
    ratio = np.exp(logratio, dtype=f32)
    adv = np.asarray(b_advantages, dtype=f32)[idx]
    adv_mean, adv_std = f32(0), f32(1)
    if norm_adv:
        adv_mean = adv.mean(dtype=f32)
        adv_std = f32(np.sqrt(((adv - adv_mean) ** 2).sum(dtype=f32) / f32(M - 1)))
        adv = (adv - adv_mean) / (adv_std + f32(1e-8))
    lo, hi = f32(1 - clip_coef), f32(1 + clip_coef)
    pg1 = -adv * ratio
    pg2 = -adv * np.clip(ratio, lo, hi)
    pg_loss = np.maximum(pg1, pg2).mean(dtype=f32)
    nv = np.asarray(new_value, dtype=f32).reshape(M)
    R = np.asarray(b_returns, dtype=f32)[idx]
    V = np.asarray(b_values, dtype=f32)[idx]
    c = f32(clip_coef)
    if clip_vloss:
        vu = (nv - R) ** 2
        d = nv - V
        vcl = V + np.clip(d, -c, c)
        vc = (vcl - R) ** 2
        v_loss = f32(0.5) * np.maximum(vu, vc).mean(dtype=f32)
        gu = f32(2) * (nv - R)
        gc = f32(2) * (vcl - R) * ((d >= -c) & (d <= c)).astype(f32)
        gv = np.where(vu > vc, gu, f32(0)) + np.where(vc > vu, gc, f32(0)) + np.where(vu == vc, f32(0.5) * (gu + gc), f32(0))
    else:
        v_loss = f32(0.5) * ((nv - R) ** 2).mean(dtype=f32)
        gv = f32(2) * (nv - R)
    ent_loss = entropy.mean(dtype=f32)
    loss = pg_loss - f32(ent_coef) * ent_loss + v_loss * f32(vf_coef)
    invM = f32(1.0) / f32(M)
    inrange = ((ratio >= lo) & (ratio <= hi)).astype(f32)
    g_ratio = np.where(pg1 > pg2, -adv, f32(0)) + np.where(pg1 == pg2, f32(0.5) * (-adv) * (f32(1) + inrange), f32(0))
    g_lp = (g_ratio * ratio * invM).astype(f32)
    std = np.exp(np.asarray(logstd, dtype=f32).reshape(1, -1))
    var = std * std
    diff = a - new_mean
    dmean = g_lp[:, None] * diff / var
    dlogstd = (g_lp[:, None] * (diff * diff / var - f32(1))).sum(axis=0, dtype=f32) - f32(ent_coef)
    dvalue = f32(vf_coef) * f32(0.5) * invM * gv
    stats = dict(pg_loss=pg_loss, v_loss=v_loss, entropy=ent_loss, old_approx_kl=(-logratio).mean(dtype=f32),
                 approx_kl=((ratio - f32(1)) - logratio).mean(dtype=f32),
                 clipfrac=(np.abs(ratio - f32(1.0)) > c).astype(f32).mean(dtype=f32), loss=loss,
                 adv_mean=adv_mean, adv_std=adv_std)
    return stats, dmean.astype(f32), dlogstd.astype(f32), dvalue.astype(f32)


# ------------------------------------------------------------------------ DQN
def dqn_td_loss(q, q_target_next, actions, rewards, dones, gamma, huber=False):
    """cleanrl/dqn_atari.py:220-224: td_target = r + gamma * max_a' Qt * (1 - d); old = Q[a]; F.mse_loss(td, old)
    (``huber``: F.smooth_l1_loss).  Returns (td_loss, mean chosen Q, dL/dQ [B, A])."""
    q = np.asarray(q, dtype=f32); qt = np.asarray(q_target_next, dtype=f32)
    B, A = q.shape
    a = np.asarray(actions).reshape(B).astype(np.int64)
    td = np.asarray(rewards, dtype=f32).reshape(B) + (f32(gamma) * qt.max(axis=1)) * (f32(1) - np.asarray(dones, dtype=f32).reshape(B))
    old = q[np.arange(B), a]
    x = old - td
    if huber:
        ax = np.abs(x)
        per = np.where(ax < 1, f32(0.5) * x * x, ax - f32(0.5))
        g = np.clip(x, -1, 1)
    else:
        per = x * x
        g = f32(2) * x
    dq = np.zeros((B, A), dtype=f32)
    dq[np.arange(B), a] = g / f32(B)
    return per.mean(dtype=f32), old.mean(dtype=f32), dq
