"""-m gpu: Gaussian-policy kernels vs the oracle, and cleanrl_b200/ppo_continuous_action.py vs the unmodified
reference run (tests/golden/ppo_continuous_n4_t64_seed2.npz, HalfCheetah-shaped synthetic env: obs 17, act 6)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5   # fp32 tolerance of BASELINE.json north_star (returns/advantages/losses)


@pytest.mark.parametrize("n,D", [(512, 6), (1, 1), (1000, 17), (33, 32)])
def test_gaussian_sample_and_eval_vs_oracle(lib, n, D):
    from cleanrl_b200 import ops
    g = torch.Generator().manual_seed(n + D)
    mean = torch.randn(n, D, generator=g)
    logstd = torch.randn(D, generator=g) * 0.3
    eps = torch.randn(n, D, generator=g)
    val = torch.randn(n, generator=g)
    a_o, lp_o, ent_o = O.gaussian_sample(mean.numpy(), logstd.numpy(), eps.numpy())
    a, lp, ent, v = ops.gaussian_sample(mean.cuda(), logstd.cuda(), eps.cuda(), val.cuda())
    assert np.abs(a.cpu().numpy() - a_o).max() <= TOL * max(1.0, np.abs(a_o).max())
    lp_o2, ent_o2 = O.gaussian_eval(mean.numpy(), logstd.numpy(), a.cpu().numpy())
    assert np.abs(lp.cpu().numpy() - lp_o2).max() <= TOL * max(1.0, np.abs(lp_o2).max())
    assert np.abs(ent.cpu().numpy() - ent_o2).max() <= TOL * max(1.0, np.abs(ent_o2).max())
    assert torch.equal(v.cpu(), val)
    lp2, ent2 = ops.gaussian_eval(mean.cuda(), logstd.cuda(), a)
    assert torch.equal(lp2, lp) and torch.equal(ent2, ent)


def test_gaussian_sample_matches_torch_normal_on_device(lib):
    """Same CUDA generator state => the actions Normal(mean, std).sample() draws on the GPU."""
    from torch.distributions.normal import Normal
    from cleanrl_b200 import ops
    n, D = 4096, 6
    mean = torch.randn(n, D, device="cuda")
    logstd = (torch.randn(1, D, device="cuda") * 0.2)
    torch.manual_seed(77)
    dist = Normal(mean, logstd.expand_as(mean).exp())
    a_t = dist.sample()
    torch.manual_seed(77)
    eps = torch.randn(n, D, device="cuda")
    a, lp, ent, _ = ops.gaussian_sample(mean, logstd.view(-1).contiguous(), eps)
    assert (a - a_t).abs().max().item() <= 1e-6 * a_t.abs().max().item()
    assert (lp - dist.log_prob(a).sum(1)).abs().max().item() <= TOL * 10
    assert (ent - dist.entropy().sum(1)).abs().max().item() <= TOL


@pytest.mark.parametrize("M,D,B", [(32768, 6, 1048576 // 8), (128, 6, 512), (2, 1, 4), (1000, 17, 3000)])
@pytest.mark.parametrize("norm_adv,clip_vloss", [(True, True), (False, False)])
def test_ppo_loss_gaussian_vs_oracle(lib, M, D, B, norm_adv, clip_vloss):
    from cleanrl_b200 import ops
    g = torch.Generator().manual_seed(M + D)
    mean = torch.randn(M, D, generator=g)
    logstd = torch.randn(D, generator=g) * 0.2
    nv = torch.randn(M, generator=g)
    b_act = torch.randn(B, D, generator=g)
    b_lp = torch.randn(B, generator=g) * 0.2 - 1.4 * D
    b_adv = torch.randn(B, generator=g); b_ret = torch.randn(B, generator=g); b_val = b_ret + 0.3 * torch.randn(B, generator=g)
    inds = torch.randperm(B, generator=g)[:M]
    st_o, dm_o, dls_o, dv_o = O.ppo_loss_gaussian(mean.numpy(), logstd.numpy(), nv.numpy(), inds.numpy(), b_act.numpy(),
                                                  b_lp.numpy(), b_adv.numpy(), b_ret.numpy(), b_val.numpy(), 0.2, 0.01, 0.5,
                                                  norm_adv, clip_vloss)
    st, dm, dls, dv = ops.ppo_loss_gaussian(mean.cuda(), logstd.cuda(), nv.cuda(), inds.cuda(), b_act.cuda(), b_lp.cuda(),
                                            b_adv.cuda(), b_ret.cuda(), b_val.cuda(), 0.2, 0.01, 0.5, norm_adv, clip_vloss)
    st = st.cpu().numpy()
    for i, k in enumerate(ops.STAT_NAMES):
        assert abs(st[i] - float(st_o[k])) <= 3 * TOL * max(1.0, abs(float(st_o[k]))), (k, st[i], st_o[k])
    assert np.abs(dm.cpu().numpy() - dm_o).max() <= 1e-4 * np.abs(dm_o).max() + 1e-12
    assert np.abs(dls.cpu().numpy() - dls_o).max() <= 1e-4 * max(1.0, np.abs(dls_o).max())
    assert np.abs(dv.cpu().numpy() - dv_o).max() <= TOL * np.abs(dv_o).max() + 1e-12


class _Writer:
    def __init__(self, *a, **k): self.scalars = []
    def add_text(self, *a, **k): pass
    def add_scalar(self, tag, v, step): self.scalars.append((tag, float(v), int(step)))
    def close(self): pass


def test_continuous_script_reproduces_reference_run(lib, tmp_path, monkeypatch):
    from cleanrl_b200 import ppo_continuous_action as S
    z = np.load(GOLDEN / "ppo_continuous_n4_t64_seed2.npz")
    argv = [a for a in z["argv"].tolist() if a != "--no-cuda"] + ["--synthetic-env", "--save-model"]
    snaps, writers, agents = [], [], []

    def on_it(it, eng, st):
        snaps.append({k: getattr(eng, k).cpu().numpy().copy() for k in
                      ("actions", "logprobs", "values", "rewards", "dones", "advantages", "returns")} | {"st": st})

    def hook(agent):
        agent.noise_fn = lambda n, D, dev: torch.randn(n, D).to(dev)    # CPU generator, as the CPU reference run
        agents.append(agent)

    def wf(p):
        w = _Writer(); writers.append(w); return w

    monkeypatch.chdir(tmp_path)
    S.main(argv, writer_factory=wf, on_iteration=on_it, agent_hook=hook)
    s0 = snaps[0]
    assert list(agents[0].state_dict().keys()) == z["state_dict_keys"].tolist()
    assert np.abs(s0["actions"] - z["actions"][0]).max() <= 2e-6 * np.abs(z["actions"][0]).max()
    assert np.array_equal(s0["dones"], z["dones"][0])
    for k in ("rewards", "logprobs", "values", "advantages", "returns"):
        d = np.abs(s0[k].astype(np.float64) - z[k][0]).max() / max(1.0, np.abs(z[k][0]).max())
        assert d <= 2 * TOL, (k, d)
    per = s0["st"]["per_update"]
    for col, key in ((0, "upd_pg_loss"), (1, "upd_v_loss"), (2, "upd_entropy_loss"), (4, "upd_approx_kl"), (6, "upd_loss")):
        ref = float(z[key][0])
        assert abs(per[0, col] - ref) <= 2 * TOL * max(1.0, abs(ref)), (key, per[0, col], ref)
    # --save-model wrote a .cleanrl_model with the reference's keys and the eval helper ran on it
    files = list(tmp_path.glob("runs/*/ppo_continuous_action.cleanrl_model"))
    assert len(files) == 1
    sd = torch.load(files[0])
    assert list(sd.keys()) == z["state_dict_keys"].tolist()
    assert any(t == "eval/episodic_return" for t, _, _ in writers[0].scalars)
