"""-m gpu: the IMPALA-CNN agent (cleanrl/ppo_procgen.py) on libb200rl: padded convolutions, max-pool, residual blocks and
their explicit backward vs torch autograd; the drop-in script vs a run of the UNMODIFIED reference script
(tests/golden/ppo_procgen_n8_t16_seed2.npz)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


class _Envs:
    def __init__(self, A):
        from cleanrl_b200.synthetic_envs import Box, Discrete
        self.single_observation_space = Box(0, 255, (64, 64, 3), np.uint8)
        self.single_action_space = Discrete(A)


def _cpu_noise(n, A, device):
    return torch.empty(n, A, dtype=torch.float32).exponential_(1).to(device)


def _ref_forward(sd, x_nhwc):
    """cleanrl/ppo_procgen.py:89-150 evaluated with torch functional ops in fp64."""
    h = x_nhwc.permute(0, 3, 1, 2).double() / 255.0
    for i in range(3):
        h = F.conv2d(h, sd[f"network.{i}.conv.weight"], sd[f"network.{i}.conv.bias"], padding=1)
        h = F.max_pool2d(h, kernel_size=3, stride=2, padding=1)
        for b in (0, 1):
            inp = h
            h = F.conv2d(F.relu(h), sd[f"network.{i}.res_block{b}.conv0.weight"], sd[f"network.{i}.res_block{b}.conv0.bias"], padding=1)
            h = F.conv2d(F.relu(h), sd[f"network.{i}.res_block{b}.conv1.weight"], sd[f"network.{i}.res_block{b}.conv1.bias"], padding=1)
            h = h + inp
    hid = F.relu(F.linear(F.relu(h.flatten(1)), sd["network.5.weight"], sd["network.5.bias"]))
    return F.linear(hid, sd["actor.weight"], sd["actor.bias"]), F.linear(hid, sd["critic.weight"], sd["critic.bias"])[:, 0]


@pytest.mark.parametrize("n,B,A", [(6, 20, 15), (1, 3, 4), (33, 33, 15)])
def test_impala_agent_forward_backward_vs_autograd(lib, n, B, A):
    from cleanrl_b200.agents import ImpalaAgent
    torch.manual_seed(2)
    agent = ImpalaAgent(_Envs(A)).cuda()
    agent.flat
    g = torch.Generator().manual_seed(4)
    obs = torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.uint8, generator=g)
    rows = torch.randperm(B, generator=g)[:n]
    ref = {k: v.detach().cpu().double().requires_grad_(True) for k, v in agent.state_dict().items()}
    logits, value = _ref_forward(ref, obs[rows])
    lg, val = agent.forward_train(obs.cuda(), rows.cuda())
    torch.cuda.synchronize()
    assert (lg.cpu().double() - logits).abs().max() <= 1e-5 * max(1.0, logits.abs().max().item())
    assert (val.cpu().double() - value).abs().max() <= 1e-5 * max(1.0, value.abs().max().item())
    gl = torch.randn(n, A, generator=g)
    gv = torch.randn(n, generator=g)
    dhead, dl, dv = agent.alloc_head_grad(n, torch.device("cuda"))
    dl.copy_(gl); dv.copy_(gv)
    agent.backward(dhead)
    torch.cuda.synchronize()
    ((logits * gl.double()).sum() + (value * gv.double()).sum()).backward()
    for k, p in agent.named_parameters():
        gr = ref[k].grad
        err = (p.grad.cpu().double() - gr).abs().max().item() / max(gr.abs().max().item(), 1e-30)
        assert err <= 1e-4, (k, err)


def test_maxpool_matches_torch(lib):
    from cleanrl_b200 import ops
    g = torch.Generator().manual_seed(1)
    for shape in ((3, 5, 64, 64), (2, 4, 7, 9), (1, 1, 1, 1)):
        x = torch.randn(shape, generator=g)
        x[0, 0, 0, :] = 1.0                                  # ties: torch takes the first maximum of the window
        xr = x.clone().requires_grad_(True)
        y = F.max_pool2d(xr, 3, 2, 1)
        dy = torch.randn(y.shape, generator=g)
        y.backward(dy)
        yk, arg = ops.maxpool3s2_fwd(x.cuda())
        dxk = ops.maxpool3s2_bwd(dy.cuda(), arg, shape[-2:])
        assert torch.equal(yk.cpu(), y.detach())
        assert torch.allclose(dxk.cpu(), xr.grad, atol=1e-6)


class _Writer:
    def __init__(self, *a, **k):
        self.scalars = []
    def add_text(self, *a, **k): pass
    def add_scalar(self, tag, v, step): self.scalars.append((tag, float(np.asarray(v).reshape(-1)[0]), int(step)))
    def close(self): pass


def test_procgen_script_reproduces_reference_run(lib):
    """cleanrl_b200/ppo_procgen.py vs the unmodified cleanrl/ppo_procgen.py (3 iterations, N = 8, T = 16, 2 epochs x 4
    minibatches): iteration 1 actions bit-exact, logprobs / values / advantages / returns <= 1e-5, first update's losses
    <= 1e-5, the iteration's other updates <= 1e-4; update counts, TensorBoard tags / steps identical."""
    from cleanrl_b200 import ppo_procgen as S
    z = np.load(GOLDEN / "ppo_procgen_n8_t16_seed2.npz")
    argv = [a for a in z["argv"].tolist() if a != "--no-cuda"] + ["--synthetic-env"]
    snaps, writers = [], []

    def on_it(it, eng, st):
        snaps.append({k: getattr(eng, k).cpu().numpy().copy() for k in
                      ("actions", "logprobs", "values", "rewards", "dones", "advantages", "returns")} | {"st": st})

    def hook(agent):
        agent.noise_fn = _cpu_noise

    def wf(path):
        w = _Writer(); writers.append(w); return w

    S.main(argv, writer_factory=wf, on_iteration=on_it, agent_hook=hook)
    n_it = z["actions"].shape[0]
    assert len(snaps) == n_it
    s = snaps[0]
    rel = lambda a, b: np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(1.0, np.abs(b).max())
    assert np.array_equal(s["actions"], z["actions"][0].astype(np.int64)), "iteration 1: actions differ"
    assert np.array_equal(s["rewards"], z["rewards"][0]) and np.array_equal(s["dones"], z["dones"][0])
    for k in ("logprobs", "values", "advantages", "returns"):
        assert rel(s[k], z[k][0]) <= 1e-5, (k, rel(s[k], z[k][0]))
    per = s["st"]["per_update"]
    assert per.shape[0] == 8
    for u in range(8):
        for col, key in ((0, "upd_pg_loss"), (1, "upd_v_loss"), (2, "upd_entropy_loss"), (4, "upd_approx_kl"), (6, "upd_loss")):
            ref = float(z[key][u])
            assert abs(per[u, col] - ref) <= (1e-5 if u == 0 else 1e-4) * max(1.0, abs(ref)), (u, key, per[u, col], ref)
    for it in range(1, n_it):
        assert (snaps[it]["actions"] == z["actions"][it].astype(np.int64)).mean() >= 0.3
    ours = {}
    for tag, v, step in writers[0].scalars:
        ours.setdefault(tag, []).append((step, v))
    for key in z.files:
        if key.startswith("tb/") and key != "tb/charts/SPS":
            tag, ref = key[3:], z[key]
            assert tag in ours, tag
            got = np.array(ours[tag])
            if not tag.startswith("charts/episodic"):
                assert np.array_equal(got[:, 0], ref[:, 0]), tag
