"""One PPO minibatch update (M = 32768, the BASELINE.json configs[1] minibatch) + a few rollout steps,
for ncu: `ncu ... python tools/profile_update.py [M] [reps] [u8|s2d|u8s2d]`.  Prints per-kernel CUDA-event times when run bare.

u8  (default): the drop-in path, uint8 NCHW frames gathered through random minibatch indices.
s2d: the engine path, rollout stored as space-to-depth bf16, sorted minibatch indices (what bench.py runs)."""
import json
import sys
import ctypes

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import ppo_args  # noqa: E402
from cleanrl_b200 import _lib, build, ops  # noqa: E402
from cleanrl_b200.agents import NatureCNNAgent  # noqa: E402
from cleanrl_b200.synthetic_envs import Box, Discrete  # noqa: E402

build.build()
lib = _lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
B = 4 * M
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fmt = sys.argv[3] if len(sys.argv) > 3 else "u8"


class E:
    single_observation_space = Box(0, 255, (4, 84, 84), np.uint8)
    single_action_space = Discrete(4)


torch.manual_seed(1)
dev = torch.device("cuda:0")
agent = NatureCNNAgent(E()).to(dev)
agent.precision = "bf16"
flat = agent.flat
obs = torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, device=dev)
inds = torch.randperm(B, device=dev)[:M].contiguous()
obs_roll = obs[:1024].clone()
aux = {}
if fmt == "s2d":
    s2d = torch.empty((B, 21, 21, 64), dtype=torch.bfloat16, device=dev)
    for lo in range(0, B, 16384):
        ops.frames_to_s2d(obs[lo:lo + 16384], out=s2d[lo:lo + 16384])
    obs = s2d
    inds = inds.sort().values.contiguous()
elif fmt == "u8s2d":       # the engine's default rollout: uint8 space-to-depth rows in both orientations
    rm = ops.alloc_u8_rollout_rows((B, 441, 64), dev)
    cm = torch.empty((B, 64, 448), dtype=torch.uint8, device=dev)
    for lo in range(0, B, 16384):
        ops.frames_to_s2d_u8(obs[lo:lo + 16384], rm[lo:lo + 16384], cm[lo:lo + 16384])
    obs = rm
    aux = {"aux": cm}
    inds = inds.sort().values.contiguous()
b_act = torch.randint(0, 4, (B,), device=dev)
b_lp = torch.full((B,), -1.386, device=dev)
b_adv = torch.randn(B, device=dev)
b_ret = torch.randn(B, device=dev)
b_val = torch.randn(B, device=dev)
dhead, dl, dv = agent.alloc_head_grad(M, dev)
stats = torch.zeros(16, device=dev)


def minibatch(step):
    logits, value = agent.forward_train(obs, inds, **aux)
    ops.ppo_loss(logits, value, inds, b_act, b_lp, b_adv, b_ret, b_val, 0.1, 0.01, 0.5, dlogits=dl, dvalue=dv, stats=stats)
    agent.backward(dhead)
    ops.clip_adam(flat.flat, flat.grad, flat.exp_avg, flat.exp_avg_sq, step, 2.5e-4)
    agent.params_updated()


roll_rm = ops.alloc_u8_rollout_rows((1024, 441, 64), dev) if fmt == "u8s2d" else None
roll_cm = torch.empty((1024, 64, 448), dtype=torch.uint8, device=dev) if fmt == "u8s2d" else None


def rollout_step():
    if fmt == "u8s2d":               # what PPOEngine.policy_step launches: frame conversion into the slot, then the forward
        ops.frames_to_s2d_u8(obs_roll, roll_rm, roll_cm)
        agent.get_action_and_value(roll_rm)
    else:
        agent.get_action_and_value(obs_roll)


for i in range(2):
    minibatch(i + 1)
    rollout_step()
torch.cuda.synchronize()
lib.b200rl_profile_reset()
lib.b200rl_profile_enable(1)
torch.cuda.profiler.start()          # ncu --profile-from-start off captures exactly the measured region
for i in range(reps):
    minibatch(i + 3)
    rollout_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
lib.b200rl_profile_enable(0)
buf = ctypes.create_string_buffer(1 << 16)
lib.b200rl_profile_summary(buf, 1 << 16)
rows = sorted(json.loads(buf.value.decode()), key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in rows)
print(f"M={M}: total kernel ms per (minibatch + 1 rollout step) = {tot / reps:.3f}")
for r in rows:
    ms = r["ms"] / r["launches"]
    tf = r["flops"] / r["launches"] / (ms * 1e-3) / 1e12 if r["flops"] else 0
    gb = r["bytes"] / r["launches"] / (ms * 1e-3) / 1e9 if r["bytes"] else 0
    print(f"  {r['name']:18s} x{r['launches'] // reps:3d}  {ms * 1e3:9.1f} us/launch  {tf:7.1f} TFLOP/s  {gb:7.1f} GB/s  share {r['ms'] / tot:5.1%}")
