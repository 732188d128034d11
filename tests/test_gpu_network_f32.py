"""-m gpu: fp32 CUDA-core layers (conv2d / linear, fwd + hand-written bwd) vs torch CPU fp32 autograd
of the same op (floating-point kernels => torch fp32 reference, tolerance written per test)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
RTOL = 2e-5   # fp32 accumulation-order differences over K <= 3136


def _close(a, b, rtol=RTOL):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    scale = max(b.abs().max().item(), 1e-12)
    return (a - b).abs().max().item() <= rtol * scale


CONVS = [  # n, Cin, H, W, Cout, K, stride, dtype, in_div
    (5, 4, 84, 84, 32, 8, 4, torch.uint8, 255.0),
    (3, 32, 20, 20, 64, 4, 2, torch.float32, 1.0),
    (3, 64, 9, 9, 64, 3, 1, torch.float32, 1.0),
    (2, 3, 11, 13, 5, 3, 2, torch.float32, 1.0),
    (1, 1, 5, 5, 1, 5, 1, torch.float32, 1.0),
]


@pytest.mark.parametrize("n,Cin,H,W,Cout,K,s,dt,div", CONVS)
def test_conv2d_fwd_bwd(lib, n, Cin, H, W, Cout, K, s, dt, div):
    from cleanrl_b200 import ops
    g = torch.Generator().manual_seed(n * 100 + Cin)
    if dt == torch.uint8:
        x = torch.randint(0, 256, (n + 3, Cin, H, W), generator=g, dtype=torch.uint8)
    else:
        x = torch.relu(torch.randn(n + 3, Cin, H, W, generator=g))
    rows = torch.randperm(n + 3, generator=g)[:n]
    w = torch.randn(Cout, Cin, K, K, generator=g) * 0.1
    b = torch.randn(Cout, generator=g) * 0.1
    xr = (x[rows].float() / div).requires_grad_(True)
    wt = w.clone().requires_grad_(True); bt = b.clone().requires_grad_(True)
    y_ref = torch.relu(F.conv2d(xr, wt, bt, stride=s))
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    xc, wc, bc, rc = x.cuda(), w.cuda(), b.cuda(), rows.cuda()
    y = ops.conv2d_fwd(xc, wc, bc, s, "relu", rows=rc, in_div=div)
    assert _close(y, y_ref)
    dy_pre = (dy * (y_ref > 0)).cuda().contiguous()      # gradient wrt the pre-activation output
    dw = torch.zeros_like(wc); db = torch.zeros_like(bc)
    ops.conv2d_bwd_weight(xc, dy_pre, dw, db, s, rows=rc, in_div=div)
    assert _close(dw, wt.grad) and _close(db, bt.grad)
    if dt != torch.uint8:
        x_post = x[rows].cuda().contiguous()
        dx = ops.conv2d_bwd_data(dy_pre, wc, x_post, "relu", s)
        assert _close(dx, xr.grad * (x[rows] > 0))


@pytest.mark.parametrize("n,inf,outf,act", [(1024, 3136, 512, "relu"), (37, 512, 5, None), (128, 4, 64, "tanh"),
                                            (300, 64, 64, "tanh"), (5, 17, 6, None), (1, 1, 1, None)])
def test_linear_fwd_bwd(lib, n, inf, outf, act):
    from cleanrl_b200 import ops
    g = torch.Generator().manual_seed(n + inf)
    x = torch.tanh(torch.randn(n + 2, inf, generator=g))
    rows = torch.randperm(n + 2, generator=g)[:n]
    w = torch.randn(outf, inf, generator=g) / np.sqrt(inf)
    b = torch.randn(outf, generator=g) * 0.1
    xr = x[rows].clone().requires_grad_(True)
    wt = w.clone().requires_grad_(True); bt = b.clone().requires_grad_(True)
    pre = F.linear(xr, wt, bt)
    y_ref = {"relu": torch.relu, "tanh": torch.tanh, None: lambda t: t}[act](pre)
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    y = ops.linear_fwd(x.cuda(), w.cuda(), b.cuda(), act, rows=rows.cuda())
    assert _close(y, y_ref)
    if act == "relu":
        dpre = dy * (y_ref > 0)
    elif act == "tanh":
        dpre = dy * (1 - y_ref.detach() ** 2)
    else:
        dpre = dy
    dpre = dpre.cuda().contiguous()
    dw = torch.zeros(outf, inf, device="cuda"); db = torch.zeros(outf, device="cuda")
    ops.linear_bwd_weight(x.cuda(), dpre, dw, db, rows=rows.cuda())
    assert _close(dw, wt.grad) and _close(db, bt.grad)
    # treat x as a tanh output of a previous layer: dx_pre = dx * (1 - x^2)
    dx = ops.linear_bwd_data(dpre, w.cuda(), x[rows].cuda().contiguous(), "tanh")
    assert _close(dx, xr.grad * (1 - x[rows] ** 2))


def _ref_agent_cnn(A):
    import torch.nn as nn

    class Ref(nn.Module):
        def __init__(self):
            super().__init__()
            self.network = nn.Sequential(nn.Conv2d(4, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(),
                                         nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(3136, 512), nn.ReLU())
            self.actor = nn.Linear(512, A)
            self.critic = nn.Linear(512, 1)
    return Ref()


class _Envs:
    def __init__(self, shape, n, dtype=np.uint8):
        from cleanrl_b200.synthetic_envs import Box, Discrete
        self.single_observation_space = Box(0, 255, shape, dtype)
        self.single_action_space = Discrete(n)


def test_naturecnn_agent_forward_backward_and_state_dict_interchange(lib):
    """Whole-agent fp32 path vs a stock torch module with the SAME state_dict keys
    (reference Agent: ppo_atari_envpool.py:123-149): values/logprobs and every parameter gradient."""
    from torch.distributions import Categorical
    from cleanrl_b200.agents import NatureCNNAgent
    from cleanrl_b200 import ops
    torch.manual_seed(3)
    A, n, B = 4, 24, 64
    agent = NatureCNNAgent(_Envs((4, 84, 84), A)).cuda()
    ref = _ref_agent_cnn(A)
    agent.flat  # bind
    sd = {k: v.detach().cpu().clone() for k, v in agent.state_dict().items()}
    assert list(sd) == list(ref.state_dict())          # identical keys, identical order
    ref.load_state_dict(sd)
    obs = torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8)
    rows = torch.randperm(B)[:n]
    act = torch.randint(0, A, (n,))
    x = obs[rows].float() / 255.0
    hid = ref.network(x)
    c = Categorical(logits=ref.actor(hid))
    val = ref.critic(hid)
    # eval path through the public method
    a2, lp, ent, v = agent.get_action_and_value(obs[rows].cuda(), act.cuda())
    assert torch.equal(a2.cpu(), act)
    assert (lp.cpu() - c.log_prob(act)).abs().max() < 1e-5
    assert (ent.cpu() - c.entropy()).abs().max() < 1e-5
    assert (v.cpu() - val).abs().max() < 1e-5 and v.shape == (n, 1)
    assert (agent.get_value(obs[rows].cuda()).cpu() - val).abs().max() < 1e-5
    # training path: gathered forward + backward of an arbitrary head gradient
    logits, value = agent.forward_train(obs.cuda(), rows.cuda())
    dhead, dl, dv = agent.alloc_head_grad(n, torch.device("cuda"))
    g = torch.Generator().manual_seed(1)
    gl = torch.randn(n, A, generator=g); gv = torch.randn(n, generator=g)
    dl.copy_(gl); dv.copy_(gv)
    agent.backward(dhead)
    (ref.actor(hid) * gl).sum().backward(retain_graph=True)
    (val[:, 0] * gv).sum().backward()
    for (k, p), (k2, q) in zip(agent.named_parameters(), ref.named_parameters()):
        assert k == k2
        assert _close(p.grad, q.grad, 5e-5), k
    # train with kernels <-> evaluate with stock torch (and back)
    flat = agent.flat
    ops.clip_adam(flat.flat, flat.grad, flat.exp_avg, flat.exp_avg_sq, 1, 1e-3)
    ref.load_state_dict({k: v.cpu() for k, v in agent.state_dict().items()})
    assert (agent.get_value(obs[rows].cuda()).cpu() - ref.critic(ref.network(x))).abs().max() < 1e-4
    agent.load_state_dict(sd)
    assert (agent.get_value(obs[rows].cuda()).cpu() - val).abs().max() < 1e-5


def test_mlp_agent_forward_backward(lib):
    from torch.distributions import Categorical
    from cleanrl_b200.agents import MLPAgent
    torch.manual_seed(0)
    envs = _Envs((4,), 2, np.float32)
    agent = MLPAgent(envs)
    import copy
    ref = copy.deepcopy(agent)
    agent = agent.cuda()
    x = torch.randn(50, 4)
    act = torch.randint(0, 2, (50,))
    c = Categorical(logits=ref.actor(x)); val = ref.critic(x)
    a2, lp, ent, v = agent.get_action_and_value(x.cuda(), act.cuda())
    assert (lp.cpu() - c.log_prob(act)).abs().max() < 1e-5 and (v.cpu() - val).abs().max() < 1e-5
    rows = torch.arange(50)
    agent.forward_train(x.cuda(), rows.cuda())
    dhead, dl, dv = agent.alloc_head_grad(50, torch.device("cuda"))
    gl = torch.randn(50, 2); gv = torch.randn(50)
    dl.copy_(gl); dv.copy_(gv)
    agent.backward(dhead)
    (ref.actor(x) * gl).sum().backward(); (ref.critic(x)[:, 0] * gv).sum().backward()
    for (k, p), (_, q) in zip(agent.named_parameters(), ref.named_parameters()):
        assert _close(p.grad, q.grad, 5e-5), k
