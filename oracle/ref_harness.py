"""Run the UNMODIFIED reference scripts on CPU and record what they compute.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works only where
/root/reference exists (this container); its outputs are committed as small
fixtures under tests/golden/ by oracle/make_golden.py.

Mechanism: fake env modules from ``oracle.stubs`` are put
in ``sys.modules``; the script runs under ``runpy.run_path(..., "__main__")``
(the same boundary cleanrl_utils/tuner.py:90-92 uses); a capturing
``SummaryWriter`` and hooks on ``np.random.shuffle`` / ``Adam.step`` snapshot
the script's module-level variables at the end of every iteration and at
every optimizer step.
"""
from __future__ import annotations

import os
import runpy
import sys
from pathlib import Path

import numpy as np
import torch

REFERENCE_ROOT = Path(os.environ.get("CLEANRL_REFERENCE", "/root/reference"))


def _main_globals():
    f = sys._getframe(1)
    while f is not None:
        if f.f_globals.get("__name__") == "__main__" and "args" in f.f_globals and "agent" in f.f_globals:
            return f.f_globals
        f = f.f_back
    return None


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().copy()
    return np.array(x).copy()


class Recorder:
    def __init__(self, keep_params=False):
        self.scalars = []          # (tag, value, step)
        self.iterations = []       # per-iteration tensor snapshots
        self.updates = []          # per optimizer step
        self.shuffles = []
        self.texts = []
        self.keep_params = keep_params
        self.last_logits = None

    # called from the capturing writer on 'charts/learning_rate' (first scalar after the update loop)
    def end_of_iteration(self, g):
        snap = {}
        for k in ("actions", "logprobs", "rewards", "dones", "values", "advantages", "returns",
                  "next_value", "next_done"):
            if k in g:
                snap[k] = _np(g[k])
        snap["param_sums"] = np.array([p.detach().double().sum().item() for p in g["agent"].parameters()])
        snap["param_abs_sums"] = np.array([p.detach().double().abs().sum().item() for p in g["agent"].parameters()])
        snap["global_step"] = int(g["global_step"])
        self.iterations.append(snap)

    def on_adam_step(self, g, before):
        if before:
            rec = {}
            for k in ("pg_loss", "v_loss", "entropy_loss", "old_approx_kl", "approx_kl", "loss"):
                if k in g:
                    rec[k] = float(g[k].detach())
            rec["clipfrac"] = float(g["clipfracs"][-1])
            rec["mb_inds_head"] = _np(g["mb_inds"][:16])
            rec["lr"] = float(g["optimizer"].param_groups[0]["lr"])
            grads = [p.grad.detach() for p in g["agent"].parameters()]
            rec["grad_norm_postclip"] = float(torch.sqrt(sum((x.double() ** 2).sum() for x in grads)))
            if self.keep_params:
                rec["grads"] = [_np(x) for x in grads]
                rec["params_before"] = [_np(p) for p in g["agent"].parameters()]
                for k in ("mb_inds", "newlogprob", "entropy", "newvalue", "b_logprobs", "b_advantages",
                          "b_returns", "b_values", "b_actions", "mb_advantages"):
                    if k in g:
                        rec[k] = _np(g[k])
                if self.last_logits is not None:
                    rec["logits"] = _np(self.last_logits)
            self.updates.append(rec)
        else:
            rec = self.updates[-1]
            ps = list(g["agent"].parameters())
            rec["param_sums"] = np.array([p.detach().double().sum().item() for p in ps])
            if self.keep_params:
                rec["params"] = [_np(p) for p in ps]


def run_reference(script, argv, atari_mode="fresh", gymnasium_kind="discrete", keep_params=False,
                  threads=None):
    """Execute /root/reference/cleanrl/<script> with ``argv`` and return (recorder, globals)."""
    from oracle import stubs

    path = REFERENCE_ROOT / "cleanrl" / script
    if not path.exists():
        raise FileNotFoundError(f"{path} (the reference only exists in the build container)")
    stubs.CONFIG["atari_mode"] = atari_mode
    stubs.CONFIG["gymnasium_kind"] = gymnasium_kind
    stubs.install()
    rec = Recorder(keep_params=keep_params)

    import torch.utils.tensorboard as tb

    class CapturingWriter:
        def __init__(self, *a, **k):
            pass

        def add_text(self, tag, text, *a, **k):
            rec.texts.append((tag, text))

        def add_scalar(self, tag, value, step=None, *a, **k):
            v = float(value) if not isinstance(value, float) else value
            rec.scalars.append((tag, v, int(step) if step is not None else None))
            if tag == "charts/learning_rate":
                g = _main_globals()
                if g is not None:
                    rec.end_of_iteration(g)

        def close(self):
            pass

    orig_writer = tb.SummaryWriter
    orig_shuffle = np.random.shuffle
    orig_step = torch.optim.Adam.step
    orig_argv = sys.argv
    orig_threads = torch.get_num_threads()

    def shuffle(x):
        orig_shuffle(x)
        rec.shuffles.append(np.array(x[:32]).copy())

    def step(self_, *a, **k):
        g = _main_globals()
        if g is not None:
            rec.on_adam_step(g, before=True)
        out = orig_step(self_, *a, **k)
        if g is not None:
            rec.on_adam_step(g, before=False)
        return out

    from torch.distributions.categorical import Categorical
    orig_cat_init = Categorical.__init__

    def cat_init(self_, probs=None, logits=None, validate_args=None):
        if logits is not None:
            rec.last_logits = logits.detach()
        return orig_cat_init(self_, probs=probs, logits=logits, validate_args=validate_args)

    Categorical.__init__ = cat_init
    tb.SummaryWriter = CapturingWriter
    np.random.shuffle = shuffle
    torch.optim.Adam.step = step
    sys.argv = [str(path)] + list(argv)
    sys.path.insert(0, str(REFERENCE_ROOT))
    if threads:
        torch.set_num_threads(threads)
    try:
        g = runpy.run_path(str(path), run_name="__main__")
    finally:
        tb.SummaryWriter = orig_writer
        Categorical.__init__ = orig_cat_init
        np.random.shuffle = orig_shuffle
        torch.optim.Adam.step = orig_step
        sys.argv = orig_argv
        sys.path.remove(str(REFERENCE_ROOT))
        torch.set_num_threads(orig_threads)
        stubs.uninstall()
    return rec, g
