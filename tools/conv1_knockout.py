"""Bottleneck analysis of the uint8 conv1 kernels: time them with pipeline stages knocked out.
  forward  (B200RL_DBG_CONV1  bits): 1 = no global stores, 2 = no TMA loads, 4 = no MMAs, 8 = no tcgen05.ld
  wgrad    (B200RL_DBG_CONV1W bits): 1 = no uint8 -> fp16 conversion, 2 = no TMA loads, 4 = no MMAs, 8 = no bias sums,
                                      16 = no tcgen05.st, 32 = no shared-memory loads
Each configuration runs in a fresh process (the flag is read once).   python tools/conv1_knockout.py [M] [fwd|wgrad]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from cleanrl_b200 import build, ops
from cleanrl_b200.agents import NatureCNNAgent
from cleanrl_b200.synthetic_envs import Box, Discrete
build.build()
M = int(sys.argv[1]); gather = sys.argv[2] == "1"
class E:
    single_observation_space = Box(0, 255, (4, 84, 84), np.uint8); single_action_space = Discrete(4)
torch.manual_seed(1); dev = torch.device("cuda")
agent = NatureCNNAgent(E()).to(dev); agent.precision = "bf16"; agent.flat
B = 4 * M if gather else M
rm = ops.alloc_u8_rollout_rows((B, 441, 64), dev); rm.random_(0, 256)
rows = torch.randperm(B, device=dev)[:M].sort().values.contiguous() if gather else None
lib = ops._lib.load()
which = sys.argv[3]
cm = torch.empty((B, 64, 448), dtype=torch.uint8, device=dev).random_(0, 256)
dhead, dl, dv = agent.alloc_head_grad(M, dev); dhead.normal_(0, 1e-3)
def go():
    agent._forward_heads(rm, rows=rows, keep=(which == "wgrad"), aux=cm)
    if which == "wgrad": agent.backward(dhead)
for _ in range(3): go()
torch.cuda.synchronize()
lib.b200rl_profile_reset(); lib.b200rl_profile_enable(1)
for _ in range(10): go()
torch.cuda.synchronize(); lib.b200rl_profile_enable(0)
import ctypes
buf = ctypes.create_string_buffer(1 << 16); lib.b200rl_profile_summary(buf, 1 << 16)
r = [x for x in json.loads(buf.value.decode()) if x["name"] == ("conv1_fwd" if which == "fwd" else "conv1_wgrad")][0]
print(json.dumps({"us": 1e3 * r["ms"] / r["launches"]}))
''' % ROOT

M = sys.argv[1] if len(sys.argv) > 1 else "32768"
which = sys.argv[2] if len(sys.argv) > 2 else "fwd"
out = []
configs = (0, 1, 2, 4, 3, 6, 7) if which == "fwd" else (0, 1, 2, 4, 8, 9, 3, 6, 15)
if os.environ.get("KNOCKOUT_CONFIGS"):
    configs = tuple(int(x) for x in os.environ["KNOCKOUT_CONFIGS"].split(","))
var = "B200RL_DBG_CONV1" if which == "fwd" else "B200RL_DBG_CONV1W"
for gather in ("1",) if which == "wgrad" else ("1", "0"):
    for dbg in configs:
        env = dict(os.environ, **{var: str(dbg)})
        r = subprocess.run([sys.executable, "-c", WORKER, M, gather, which], env=env, capture_output=True, text=True)
        try:
            us = json.loads(r.stdout.strip().splitlines()[-1])["us"]
        except Exception:
            us = None
            print(r.stderr[-500:])
        labels = ((1, "no-store"), (2, "no-tma"), (4, "no-mma"), (8, "no-tmem-ld")) if which == "fwd" else \
            ((1, "no-convert"), (2, "no-tma"), (4, "no-mma"), (8, "no-bias-sums"), (16, "no-tmem-st"), (32, "no-lds"))
        names = [n for b, n in labels if dbg & b] or ["full"]
        out.append({"gather": gather == "1", "dbg": dbg, "config": "+".join(names), "us": us})
        print(out[-1], flush=True)
print(json.dumps(out))
