"""-m gpu (needs >= 2 GPUs, skipped otherwise): cleanrl_b200/ppo_atari_multigpu.py under torchrun with NCCL --
one all-reduce of the flat gradient per update over NVLink; replicas must stay identical (reference debug print,
cleanrl/ppo_atari_multigpu.py:284-286)."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("script", ["ppo_atari_multigpu.py", "ppo_atari_multigpu_envpool.py"])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_two_rank_nccl_replicas_identical(lib, precision, script):
    """bf16: the gradient exchange is split (fc + heads on a side stream under the conv backward, conv layers after it);
    fp32: one all-reduce.  Either way replicas must stay bit-identical."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "cleanrl_b200" / script), "--backend", "nccl",
           "--local-num-envs", "16", "--num-steps", "8", "--total-timesteps", "768", "--synthetic-env",
           "--precision", precision, "--seed", "3"]
    env = dict(os.environ, CLEANRL_B200_TB_OFF="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    sums = {}
    for m in re.finditer(r"local_rank: (\d), action.sum\(\): (-?\d+), iteration: (\d+), agent.actor.weight.sum\(\): (\S+)", r.stdout):
        sums.setdefault(int(m.group(3)), {})[int(m.group(1))] = (int(m.group(2)), float(m.group(4)))
    assert len(sums) == 3
    for it, d in sums.items():
        assert set(d) == {0, 1}
        assert d[0][1] == d[1][1], (it, d)          # identical replicas after every all-reduced update
    assert any(d[0][0] != d[1][0] for d in sums.values())   # different env/sampling streams per rank
    assert "SPS:" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_update_graphs_with_captured_exchange_match_eager(lib):
    """CLEANRL_B200_UPDATE_GRAPHS_DP=1 replays the update -- including the overlapped NCCL gradient exchange -- from per-epoch
    CUDA graphs: replicas stay identical and every printed sum equals the launch-by-launch run's."""
    outs = []
    for dp_graphs in ("0", "1"):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
               "--master-port", "29541", str(ROOT / "cleanrl_b200" / "ppo_atari_multigpu_envpool.py"), "--backend", "nccl",
               "--local-num-envs", "16", "--num-steps", "8", "--total-timesteps", str(2 * 16 * 8 * 5), "--synthetic-env",
               "--precision", "bf16", "--seed", "3"]
        env = dict(os.environ, CLEANRL_B200_TB_OFF="1", CLEANRL_B200_UPDATE_GRAPHS_DP=dp_graphs)
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        sums = {}
        for m in re.finditer(r"local_rank: (\d), action.sum\(\): (-?\d+), iteration: (\d+), agent.actor.weight.sum\(\): (\S+)", r.stdout):
            sums[(int(m.group(3)), int(m.group(1)))] = (int(m.group(2)), m.group(4))
        assert len(sums) == 10
        for it in range(1, 6):
            assert sums[(it, 0)][1] == sums[(it, 1)][1], (dp_graphs, it)
        outs.append(sums)
    assert outs[0] == outs[1]
