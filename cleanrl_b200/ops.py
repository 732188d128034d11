"""Thin torch-tensor front end over the C-ABI (device pointers + current stream).

PyTorch is plumbing here: it owns device memory and the stream; every function
below only validates arguments and enqueues libb200rl kernels on
``torch.cuda.current_stream()``.  CPU tensors are rejected loudly -- there is
no fallback path.
"""
from __future__ import annotations

import torch

from . import _lib

STAT_NAMES = ("pg_loss", "v_loss", "entropy", "old_approx_kl", "approx_kl", "clipfrac", "loss",
              "adv_mean", "adv_std")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t, dtype=None, name="tensor", allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError(f"{name} is None")
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: libb200rl kernels need a CUDA tensor (got {type(t).__name__} on "
                           f"{getattr(t, 'device', '?')}); there is no CPU fallback")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def _contig(t, name):
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


class Workspace:
    """Grow-only device scratch buffer owned by the caller side (never by the kernels)."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
        return self.buf


_ws = {}


def _workspace(device, key, nbytes):
    w = _ws.setdefault((device, key), Workspace(device))
    return w.get(nbytes)


def gae(rewards, values, dones, next_value, next_done, gamma, gae_lambda, mode=0, out=None):
    """advantages, returns = GAE(...)  (reference loop: cleanrl/ppo.py:218-231)."""
    lib = _lib.load()
    T, N = rewards.shape
    for n, t in (("rewards", rewards), ("values", values), ("dones", dones)):
        _contig(t, n)
        assert t.shape == (T, N), f"{n} shape {tuple(t.shape)} != {(T, N)}"
    next_value = _contig(next_value.reshape(-1), "next_value")
    next_done = _contig(next_done.reshape(-1), "next_done")
    assert next_value.numel() == N and next_done.numel() == N
    if out is None:
        adv = torch.empty_like(rewards)
        ret = torch.empty_like(rewards)
    else:
        adv, ret = out
    f = torch.float32
    rc = lib.b200rl_gae_f32(_ptr(rewards, f, "rewards"), _ptr(values, f, "values"), _ptr(dones, f, "dones"),
                            _ptr(next_value, f, "next_value"), _ptr(next_done, f, "next_done"),
                            _ptr(adv, f, "advantages"), _ptr(ret, f, "returns"),
                            T, N, float(gamma), float(gae_lambda), int(mode), _stream())
    _lib.check(rc, "gae")
    return adv, ret


def categorical_sample(logits, noise, value_in=None, out=None):
    """action, logprob, entropy[, value] from raw logits and Exp(1) noise
    (reference: Categorical(logits).sample()/log_prob/entropy, ppo_atari_envpool.py:143-149)."""
    lib = _lib.load()
    n, A = logits.shape
    assert logits.stride(1) == 1
    _contig(noise, "noise")
    assert noise.shape == (n, A)
    dev = logits.device
    if out is None:
        action = torch.empty(n, dtype=torch.int64, device=dev)
        logprob = torch.empty(n, dtype=torch.float32, device=dev)
        entropy = torch.empty(n, dtype=torch.float32, device=dev)
        value = torch.empty(n, dtype=torch.float32, device=dev) if value_in is not None else None
    else:
        action, logprob, entropy, value = out
    f = torch.float32
    ldv = 0
    if value_in is not None:
        value_in = value_in.reshape(n, -1)
        ldv = value_in.stride(0)
    rc = lib.b200rl_categorical_sample_f32(
        _ptr(logits, f, "logits"), logits.stride(0), _ptr(noise, f, "noise"),
        _ptr(value_in, f, "value_in", True), ldv, n, A,
        _ptr(action, torch.int64, "action"), _ptr(logprob, f, "logprob"),
        _ptr(entropy, f, "entropy", True), _ptr(value, f, "value_out", True), _stream())
    _lib.check(rc, "categorical_sample")
    return action, logprob, entropy, value


def ppo_loss(new_logits, new_value, mb_inds, b_actions, b_logprobs, b_advantages, b_returns, b_values,
             clip_coef, ent_coef, vf_coef, norm_adv=True, clip_vloss=True, dlogits=None, dvalue=None, stats=None):
    """Fused PPO minibatch loss + gradients (reference: cleanrl/ppo.py:250-285).
    Returns (stats f32[16] device tensor, dlogits [M,A], dvalue [M])."""
    lib = _lib.load()
    M, A = new_logits.shape
    dev = new_logits.device
    assert new_logits.stride(1) == 1
    new_value = new_value.reshape(M, -1)
    f = torch.float32
    if dlogits is None:
        dlogits = torch.empty(M, A, dtype=f, device=dev)
    if dvalue is None:
        dvalue = torch.empty(M, dtype=f, device=dev)
    dv2 = dvalue.reshape(M, -1)
    if stats is None:
        stats = torch.zeros(16, dtype=f, device=dev)
    nbytes = lib.b200rl_ppo_loss_workspace_bytes(M)
    ws = _workspace(dev, "loss", nbytes)
    if mb_inds is not None:
        _contig(mb_inds, "mb_inds")
        assert mb_inds.numel() == M
    for n_, t in (("b_actions", b_actions), ("b_logprobs", b_logprobs), ("b_advantages", b_advantages),
                  ("b_returns", b_returns), ("b_values", b_values)):
        _contig(t, n_)
    rc = lib.b200rl_ppo_loss_f32(
        _ptr(new_logits, f, "new_logits"), new_logits.stride(0), _ptr(new_value, f, "new_value"), new_value.stride(0),
        _ptr(mb_inds, torch.int64, "mb_inds", True),
        _ptr(b_actions, torch.int64, "b_actions"), _ptr(b_logprobs, f, "b_logprobs"),
        _ptr(b_advantages, f, "b_advantages"), _ptr(b_returns, f, "b_returns"), _ptr(b_values, f, "b_values"),
        M, A, float(clip_coef), float(ent_coef), float(vf_coef), int(bool(norm_adv)), int(bool(clip_vloss)),
        _ptr(dlogits, f, "dlogits"), dlogits.stride(0), _ptr(dv2, f, "dvalue"), dv2.stride(0),
        _ptr(stats, f, "stats"), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "ppo_loss")
    return stats, dlogits, dvalue


def clip_adam(params, grads, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-5,
              max_norm=0.5, world_size=1, norm_out=None):
    """In-place fused clip_grad_norm_ + Adam step on flat f32 vectors
    (reference: cleanrl/ppo.py:289-290; DP averaging ppo_atari_multigpu.py:369-373)."""
    lib = _lib.load()
    P = params.numel()
    f = torch.float32
    for n_, t in (("params", params), ("grads", grads), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        _contig(t, n_)
        assert t.numel() == P
    nbytes = lib.b200rl_clip_adam_workspace_bytes(P)
    ws = _workspace(params.device, "adam", nbytes)
    rc = lib.b200rl_clip_adam_f32(
        _ptr(params, f, "params"), _ptr(grads, f, "grads"), _ptr(exp_avg, f, "exp_avg"),
        _ptr(exp_avg_sq, f, "exp_avg_sq"), P, int(step), float(lr), float(beta1), float(beta2), float(eps),
        -1.0 if max_norm is None else float(max_norm), int(world_size),
        _ptr(norm_out, f, "norm_out", True), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "clip_adam")
    return params
