"""-m gpu: the recurrent agent (cleanrl/ppo_atari_lstm.py) on libb200rl: LSTM cell kernels + explicit BPTT vs torch autograd,
and the whole drop-in script vs a run of the UNMODIFIED reference script (tests/golden/ppo_atari_lstm_n8_t16_seed4.npz)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


class _Envs:
    def __init__(self, A, c=1):
        from cleanrl_b200.synthetic_envs import Box, Discrete
        self.single_observation_space = Box(0, 255, (c, 84, 84), np.uint8)
        self.single_action_space = Discrete(A)


def _cpu_noise(n, A, device):
    return torch.empty(n, A, dtype=torch.float32).exponential_(1).to(device)


@pytest.mark.parametrize("S,n,A", [(5, 3, 4), (16, 2, 6), (1, 7, 4)])
def test_lstm_agent_forward_backward_vs_autograd(lib, S, n, A):
    """get_states / heads / BPTT of LSTMAgent against the same modules evaluated by torch (nn.LSTM stepped with the
    done-masked state exactly as cleanrl/ppo_atari_lstm.py:128-150) and autograd: outputs 1e-5, every gradient 1e-4."""
    from cleanrl_b200.agents import LSTMAgent
    torch.manual_seed(3)
    agent = LSTMAgent(_Envs(A)).cuda()
    agent.flat
    g = torch.Generator().manual_seed(5)
    B = S * n + 11
    obs = torch.randint(0, 256, (B, 1, 84, 84), dtype=torch.uint8, generator=g)
    rows = torch.randperm(B, generator=g)[:S * n]
    done = (torch.rand(S * n, generator=g) < 0.25).float()
    h0 = torch.randn(1, n, 128, generator=g) * 0.3
    c0 = torch.randn(1, n, 128, generator=g) * 0.3
    # ---- torch reference on CPU (fp64 for a clean target)
    import copy
    ref = {k: v.detach().cpu().double().requires_grad_(True) for k, v in agent.state_dict().items()}
    x = obs[rows].double() / 255.0
    a = torch.relu(F.conv2d(x, ref["network.0.weight"], ref["network.0.bias"], stride=4))
    a = torch.relu(F.conv2d(a, ref["network.2.weight"], ref["network.2.bias"], stride=2))
    a = torch.relu(F.conv2d(a, ref["network.4.weight"], ref["network.4.bias"], stride=1))
    feats = torch.relu(F.linear(a.flatten(1), ref["network.7.weight"], ref["network.7.bias"])).reshape(S, n, 512)
    h, c = h0[0].double(), c0[0].double()
    outs = []
    dd = done.double().reshape(S, n)
    for t in range(S):
        h = (1 - dd[t]).view(-1, 1) * h
        c = (1 - dd[t]).view(-1, 1) * c
        gates = F.linear(feats[t], ref["lstm.weight_ih_l0"], ref["lstm.bias_ih_l0"]) + F.linear(h, ref["lstm.weight_hh_l0"], ref["lstm.bias_hh_l0"])
        i, f, gg, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    hidden = torch.cat(outs)
    logits = F.linear(hidden, ref["actor.weight"], ref["actor.bias"])
    value = F.linear(hidden, ref["critic.weight"], ref["critic.bias"])[:, 0]
    # ---- kernels
    lg, val = agent.forward_train(obs.cuda(), rows.cuda(), (h0.cuda(), c0.cuda()), torch.zeros(B).index_copy_(0, rows, done).cuda())
    torch.cuda.synchronize()
    assert (lg.cpu().double() - logits).abs().max() <= 1e-5 * max(1.0, logits.abs().max().item())
    assert (val.cpu().double() - value).abs().max() <= 1e-5 * max(1.0, value.abs().max().item())
    hid_k, (hS, cS) = agent.get_states(obs[rows].cuda(), (h0.cuda(), c0.cuda()), done.cuda())
    assert (hid_k.cpu().double() - hidden).abs().max() <= 1e-5
    assert (hS[0].cpu().double() - h).abs().max() <= 1e-5 and (cS[0].cpu().double() - c).abs().max() <= 1e-5
    gl = torch.randn(S * n, A, generator=g)
    gv = torch.randn(S * n, generator=g)
    agent.forward_train(obs.cuda(), rows.cuda(), (h0.cuda(), c0.cuda()), torch.zeros(B).index_copy_(0, rows, done).cuda())
    dhead, dl, dv = agent.alloc_head_grad(S * n, torch.device("cuda"))
    dl.copy_(gl); dv.copy_(gv)
    agent.backward(dhead)
    torch.cuda.synchronize()
    ((logits * gl.double()).sum() + (value * gv.double()).sum()).backward()
    for k, p in agent.named_parameters():
        gr = ref[k].grad
        err = (p.grad.cpu().double() - gr).abs().max().item() / max(gr.abs().max().item(), 1e-30)
        assert err <= 1e-4, (k, err)


class _Writer:
    def __init__(self, *a, **k):
        self.scalars = []
    def add_text(self, *a, **k): pass
    def add_scalar(self, tag, v, step): self.scalars.append((tag, float(np.asarray(v).reshape(-1)[0]), int(step)))
    def close(self): pass


def test_lstm_script_reproduces_reference_run(lib):
    """cleanrl_b200/ppo_atari_lstm.py vs the unmodified cleanrl/ppo_atari_lstm.py (3 iterations, N = 8, T = 16): iteration 1
    actions bit-exact (same torch CPU noise stream), logprobs / values / advantages / returns <= 1e-5, the first update's
    losses <= 1e-5, the iteration's other updates <= 1e-4; env-wise minibatch order (numpy shuffle of env indices), update
    counts, TensorBoard tags / steps / learning rates identical."""
    from cleanrl_b200 import ppo_atari_lstm as S
    z = np.load(GOLDEN / "ppo_atari_lstm_n8_t16_seed4.npz")
    argv = [a for a in z["argv"].tolist() if a != "--no-cuda"] + ["--synthetic-env"]
    snaps, writers = [], []

    def on_it(it, state, st):
        snaps.append({k: state[k].cpu().numpy().copy() for k in
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
    assert per.shape[0] == 16
    for u in range(16):
        for col, key in ((0, "upd_pg_loss"), (1, "upd_v_loss"), (2, "upd_entropy_loss"), (4, "upd_approx_kl"), (6, "upd_loss")):
            ref = float(z[key][u])
            assert abs(per[u, col] - ref) <= (1e-5 if u == 0 else 1e-4) * max(1.0, abs(ref)), (u, key, per[u, col], ref)
    for it in range(1, n_it):
        assert (snaps[it]["actions"] == z["actions"][it].astype(np.int64)).mean() >= 0.5
        assert snaps[it]["st"]["per_update"].shape[0] == 16
    ours = {}
    for tag, v, step in writers[0].scalars:
        ours.setdefault(tag, []).append((step, v))
    for key in z.files:
        if not key.startswith("tb/") or key == "tb/charts/SPS":
            continue
        tag, ref = key[3:], z[key]
        assert tag in ours, tag
        got = np.array(ours[tag])
        if tag == "charts/learning_rate":
            assert np.array_equal(got, ref), tag
        elif tag.startswith("charts/episodic"):
            k1 = int((ref[:, 0] <= z["tb/charts/learning_rate"][0, 0]).sum())
            assert np.allclose(got[:k1, 1], ref[:k1, 1], rtol=1e-6), tag
        else:
            assert np.array_equal(got[:, 0], ref[:, 0]), tag
            tol = 1.01 / 32 if tag == "losses/clipfrac" else 2e-4
            assert abs(got[0, 1] - ref[0, 1]) <= tol * max(1.0, abs(ref[0, 1])), (tag, got[0], ref[0])
