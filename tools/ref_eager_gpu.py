"""Same-GPU eager baseline (SURVEY.md §8d "reference timed beside it", item ii).

Runs the torch restatement of the reference loop (oracle/ppo_port.py, validated against the unmodified
script by tests/test_oracle_port.py) with device="cuda": fp32 rollout storage on the GPU, eager
torch/cuDNN/cuBLAS ops, autograd, foreach Adam — what `python cleanrl/ppo_atari_envpool.py --cuda`
executes, minus envpool (synthetic frames from a host pool).  Measurement tool, never imported by the product.

    python tools/ref_eager_gpu.py [num_envs] [num_steps] [iterations]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from oracle import ppo_port


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    assert torch.cuda.is_available()
    out = ppo_port.run(num_envs=n, num_steps=t, num_iterations=iters, seed=1, env_mode="pool", device="cuda",
                       total_iterations=max(iters, 10))
    secs = out["iter_seconds"]
    steady = secs[1:] if len(secs) > 1 else secs
    print(json.dumps({"impl": "torch eager (reference loop restated), same GPU", "num_envs": n, "num_steps": t,
                      "iter_seconds": [round(x, 4) for x in secs],
                      "sps_steady": round(n * t * len(steady) / sum(steady), 1),
                      "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
                      "losses_last": out["losses"][-1]}))


if __name__ == "__main__":
    main()
