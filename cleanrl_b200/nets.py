"""Network execution plans over libb200rl layers (forward AND hand-written backward).

An ``nn.Module`` agent keeps the reference's module tree (so ``state_dict()``
keys match the reference's, cleanrl_utils/evals/ppo_eval.py:18-26) but its
parameters are re-pointed at ONE flat fp32 device buffer (``FlatParams``); a
second flat buffer receives the gradients.  One flat gradient = one
``ncclAllReduce`` per update (ppo_atari_multigpu.py:360-374 does cat + copy
instead) and one fused clip+Adam launch.

No autograd anywhere: each layer's backward is an explicit kernel call.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class FlatParams:
    """Flat fp32 parameter / gradient / Adam-state storage for a list of nn.Parameters.

    ``order`` lets a plan choose its own memory order (e.g. actor.weight next to
    critic.weight so both heads are one [A+1, hidden] GEMM operand); every
    elementwise consumer (all-reduce SUM, Adam) and the global L2 norm are
    order-independent.
    """

    def __init__(self, params, device):
        self.params = list(params)
        self.numel = sum(p.numel() for p in self.params)
        pad = (-self.numel) % 4  # float4 kernels
        self.flat = torch.zeros(self.numel + pad, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.views, self.gviews = [], []
        off = 0
        for p in self.params:
            n = p.numel()
            v = self.flat[off:off + n].view(p.shape)
            v.copy_(p.data.to(device=device, dtype=torch.float32))
            p.data = v                       # module now aliases the flat buffer
            g = self.grad[off:off + n].view(p.shape)
            p.grad = g
            self.views.append(v)
            self.gviews.append(g)
            off += n
        self.step = 0

    def view_of(self, p):
        for q, v, g in zip(self.params, self.views, self.gviews):
            if q is p:
                return v, g
        raise KeyError("parameter not in flat buffer")


class _Layer:
    pass


class Conv(_Layer):
    def __init__(self, conv: nn.Conv2d, act, in_div=1.0):
        assert conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1] and conv.dilation == (1, 1)
        self.m, self.stride, self.act, self.in_div, self.pad = conv, conv.stride[0], act, in_div, int(conv.padding[0])

    def fwd(self, x, rows=None):
        return ops.conv2d_fwd(x, self.m.weight.data, self.m.bias.data, self.stride, self.act, rows=rows, in_div=self.in_div,
                              pad=self.pad)

    def bwd_weight(self, x, dy, rows=None):
        ops.conv2d_bwd_weight(x, dy, self.m.weight.grad, self.m.bias.grad, self.stride, rows=rows, in_div=self.in_div,
                              pad=self.pad)

    def bwd_data(self, dy, x_post, prev_act, in_hw=None):
        return ops.conv2d_bwd_data(dy, self.m.weight.data, x_post, prev_act, self.stride, pad=self.pad, in_hw=in_hw)


class Linear(_Layer):
    def __init__(self, lin, act, weight=None, bias=None, wgrad=None, bgrad=None):
        self.m, self.act = lin, act
        self._w, self._b, self._dw, self._db = weight, bias, wgrad, bgrad

    @property
    def w(self):
        return self._w if self._w is not None else self.m.weight.data

    @property
    def b(self):
        return self._b if self._b is not None else self.m.bias.data

    @property
    def dw(self):
        return self._dw if self._dw is not None else self.m.weight.grad

    @property
    def db(self):
        return self._db if self._db is not None else self.m.bias.grad

    def fwd(self, x, rows=None):
        return ops.linear_fwd(x, self.w, self.b, self.act, rows=rows)

    def bwd_weight(self, x, dy, rows=None):
        ops.linear_bwd_weight(x.reshape(x.shape[0], -1), dy, self.dw, self.db, rows=rows)

    def bwd_data(self, dy, x_post, prev_act):
        xp = None if x_post is None else x_post.reshape(x_post.shape[0], -1)
        return ops.linear_bwd_data(dy, self.w, xp, prev_act)


class Chain:
    """layer_1 -> ... -> layer_k with cached activations for the backward sweep."""

    def __init__(self, layers):
        self.layers = layers
        self.acts = None
        self.x = None
        self.rows = None

    def fwd(self, x, rows=None, keep=False):
        acts = []
        h = x
        for i, l in enumerate(self.layers):
            h = l.fwd(h, rows=rows if i == 0 else None)
            acts.append(h)
        if keep:
            self.acts, self.x, self.rows = acts, x, rows
        return h

    def bwd(self, dy, need_dx=False):
        """dy = gradient wrt the LAST layer's pre-activation output... (last act must be folded by caller)"""
        L = self.layers
        for i in reversed(range(len(L))):
            x_in = self.acts[i - 1] if i > 0 else self.x
            L[i].bwd_weight(x_in, dy, rows=self.rows if i == 0 else None)
            if i > 0:
                dy = L[i].bwd_data(dy, x_in, L[i - 1].act)
                if isinstance(L[i], Linear) and x_in.dim() == 4:
                    dy = dy.view(x_in.shape)
        self.acts = None
        return None
