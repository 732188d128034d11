"""Diagnostic: run bench-like resident iterations and stop at the first minibatch update whose flat gradient holds a
non-finite entry; print where (decoded conv1 weight coordinates), what, and whether it reproduces."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import ppo_args  # noqa: E402
from cleanrl_b200 import build, ops  # noqa: E402
from cleanrl_b200 import ppo_engine as pe  # noqa: E402
from cleanrl_b200.agents import NatureCNNAgent  # noqa: E402
from cleanrl_b200.synthetic_envs import SyntheticAtariVec  # noqa: E402

build.build()
N, T = 1024, 128
dev = torch.device("cuda:0")
args = ppo_args(N, T, 48, "bf16")
np.random.seed(1); torch.manual_seed(1)
envs = SyntheticAtariVec(N, seed=1, mode="pool", pinned=True)
envs.single_observation_space, envs.single_action_space = envs.observation_space, envs.action_space
agent = NatureCNNAgent(envs).to(dev)
agent.precision = "bf16"
torch.manual_seed(1)
eng = pe.PPOEngine(agent, args, (4, 84, 84), np.uint8, N, dev, gae_mode=1)
eng.update_graphs = False
pool_dev = torch.from_numpy(envs._batches).to(dev)
g = torch.Generator().manual_seed(1)
rew = torch.randint(0, 2, (T, N), generator=g).float().to(dev)
don = (torch.rand(T, N, generator=g) < 0.02).float().to(dev)
found = {}
orig = ops.clip_adam
count = [0]


def checked(params, grads, *a, **k):
    count[0] += 1
    if not found:
        bad = ~torch.isfinite(grads)
        nb = int(bad.sum())
        if nb:
            idx = bad.nonzero().flatten()
            vals = grads[idx[:8]].tolist()
            i = idx.cpu().numpy()
            conv1 = i[i < 8192]
            dec = [(int(x) // 256, (int(x) % 256) // 64, (int(x) % 64) // 8, int(x) % 8) for x in conv1[:12]]
            found.update({"update": count[0], "nonfinite": nb, "first_idx": i[:12].tolist(), "values": vals,
                          "in_conv1_w": int(len(conv1)), "conv1_decode(co,c,ky,kx)": dec,
                          "co_set": sorted({int(x) // 256 for x in conv1}), "cin_set": sorted({(int(x) % 256) // 64 for x in conv1}),
                          "ky_set": sorted({(int(x) % 64) // 8 for x in conv1}), "kx_set": sorted({int(x) % 8 for x in conv1}),
                          "other_ranges": [int(x) for x in i[i >= 8192][:8]],
                          "finite_abs_max": float(grads[torch.isfinite(grads)].abs().max()),
                          "logits_finite": None})
    return orig(params, grads, *a, **k)


last_bwd = {}
_tc_holder = {}


def install_bwd_recorder():
    tc = agent._tc_plan()
    if "orig" in _tc_holder:
        return
    _tc_holder["orig"] = tc.backward

    def rec(*a, **k):
        last_bwd["a"], last_bwd["k"] = a, k
        return _tc_holder["orig"](*a, **k)
    tc.backward = rec


def inspect_failure():
    tc = agent._tc_plan()
    M = eng.M
    acts = tc._acts[(M, 2)].view(torch.float16)
    off = M * 49216
    d1 = acts[off: off + M * 14112].view(M, 441, 32)
    fin = torch.isfinite(d1)
    out = {"dact1_nonfinite": int((~fin).sum()), "dact1_absmax_finite": float(d1[fin].abs().max()),
           "dact1_saturated(65504)": int((d1.abs() == 65504).sum())}
    if out["dact1_nonfinite"]:
        w = (~fin).nonzero()[:6].tolist()
        out["dact1_nonfinite_where(img,pos,co)"] = w
        out["dact1_nonfinite_co_set"] = sorted(set((~fin).nonzero()[:, 2].tolist()))
        out["dact1_nonfinite_pos_set"] = sorted(set((~fin).nonzero()[:, 1].tolist()))[:20]
    # is the failure reproducible on identical inputs?
    reps = []
    for _ in range(4):
        eng.flat.grad.zero_()
        _tc_holder["orig"](*last_bwd["a"], **last_bwd["k"])
        torch.cuda.synchronize()
        bad = (~torch.isfinite(eng.flat.grad)).nonzero().flatten()
        reps.append((int(bad.numel()), sorted({int(x) // 256 for x in bad.tolist() if x < 8192})))
    out["rerun_nonfinite(count, co_set)"] = reps
    return out


_orig_checked = checked


def checked2(params, grads, *a, **k):
    was = bool(found)
    r = _orig_checked(params, grads, *a, **k)
    if found and not was:
        found["inspect"] = inspect_failure()
    return r


pe.ops.clip_adam = checked2
install_bwd_recorder()
it = 0
while not found and it < (int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    eng.rollout_resident(pool_dev)
    eng.rewards.copy_(rew); eng.dones.copy_(don)
    eng._to_storage(pool_dev[T % pool_dev.shape[0]], None)
    eng.finish_rollout(None, None, resident=True)
    st = eng.update(2.5e-4)
    it += 1
found["iterations_run"] = it
found["adv_absmax"] = float(eng.advantages.abs().max())
found["returns_absmax"] = float(eng.returns.abs().max())
found["values_absmax"] = float(eng.values.abs().max())
found["params_nonfinite"] = int((~torch.isfinite(eng.flat.flat)).sum())
print(json.dumps(found))
