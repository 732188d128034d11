"""-m gpu parity tests: GAE / sampler / loss / clip+Adam kernels through the
C-ABI against the CPU oracle (oracle/ppo_oracle.py) on the same seeded inputs.

Bars (north_star): bit-exact for integer outputs (actions) and for the
sequential GAE; <= 1e-5 (fp32) for returns/advantages/losses/Adam.
"""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5  # tolerance stated by BASELINE.json north_star for fp32 outputs


def _gae_inputs(T, N, seed=0, p_done=0.02):
    g = torch.Generator().manual_seed(seed)
    r = torch.randint(-1, 2, (T, N), generator=g).float()
    v = torch.randn(T, N, generator=g)
    d = (torch.rand(T, N, generator=g) < p_done).float()
    nv = torch.randn(N, generator=g)
    nd = (torch.rand(N, generator=g) < p_done).float()
    return r, v, d, nv, nd


# shapes: reference jax test (123,7); C1 (128,4); C2 (128,1024); C4 (2048,512); ragged / tiny edges
GAE_SHAPES = [(123, 7), (128, 4), (128, 1024), (2048, 512), (1, 1), (1, 33), (17, 100), (5, 1025)]


@pytest.mark.parametrize("T,N", GAE_SHAPES)
@pytest.mark.parametrize("gamma,lam", [(0.99, 0.95), (0.9, 0.8), (1.0, 1.0)])
def test_gae_sequential_bit_exact(lib, T, N, gamma, lam):
    from cleanrl_b200 import ops
    r, v, d, nv, nd = _gae_inputs(T, N, seed=T * 131 + N)
    adv_o, ret_o = O.gae(r.numpy(), v.numpy(), d.numpy(), nv.numpy(), nd.numpy(), gamma, lam)
    adv, ret = ops.gae(r.cuda(), v.cuda(), d.cuda(), nv.cuda(), nd.cuda(), gamma, lam, mode=0)
    assert np.array_equal(adv.cpu().numpy(), adv_o)
    assert np.array_equal(ret.cpu().numpy(), ret_o)


@pytest.mark.parametrize("T,N", GAE_SHAPES)
def test_gae_chunked_scan_tolerance(lib, T, N):
    from cleanrl_b200 import ops
    r, v, d, nv, nd = _gae_inputs(T, N, seed=T * 7 + N)
    adv_o, ret_o = O.gae(r.numpy(), v.numpy(), d.numpy(), nv.numpy(), nd.numpy(), 0.99, 0.95)
    adv, ret = ops.gae(r.cuda(), v.cuda(), d.cuda(), nv.cuda(), nd.cuda(), 0.99, 0.95, mode=1)
    scale = max(1.0, float(np.abs(adv_o).max()))
    assert np.abs(adv.cpu().numpy() - adv_o).max() <= TOL * scale
    assert np.abs(ret.cpu().numpy() - ret_o).max() <= TOL * scale


def test_gae_all_done_and_no_done(lib):
    from cleanrl_b200 import ops
    T, N = 64, 96
    r, v, d, nv, nd = _gae_inputs(T, N, seed=3)
    for fill in (0.0, 1.0):
        d2 = torch.full_like(d, fill)
        nd2 = torch.full_like(nd, fill)
        adv_o, ret_o = O.gae(r.numpy(), v.numpy(), d2.numpy(), nv.numpy(), nd2.numpy(), 0.99, 0.95)
        for mode in (0, 1):
            adv, ret = ops.gae(r.cuda(), v.cuda(), d2.cuda(), nv.cuda(), nd2.cuda(), 0.99, 0.95, mode=mode)
            if mode == 0:
                assert np.array_equal(adv.cpu().numpy(), adv_o)
            else:
                assert np.abs(adv.cpu().numpy() - adv_o).max() <= TOL * max(1.0, np.abs(adv_o).max())


def test_gae_matches_torch_reference_loop_on_device(lib):
    """The reference loop (ppo.py:218-231) executed with torch ops ON THE GPU vs our kernel."""
    from cleanrl_b200 import ops
    T, N = 128, 1024
    r, v, d, nv, nd = [x.cuda() for x in _gae_inputs(T, N, seed=11)]
    adv_t = torch.zeros_like(r)
    last = 0
    for t in reversed(range(T)):
        if t == T - 1:
            nnt, nvs = 1.0 - nd, nv
        else:
            nnt, nvs = 1.0 - d[t + 1], v[t + 1]
        delta = r[t] + 0.99 * nvs * nnt - v[t]
        adv_t[t] = last = delta + 0.99 * 0.95 * nnt * last
    adv, ret = ops.gae(r, v, d, nv, nd, 0.99, 0.95, mode=0)
    assert torch.equal(adv, adv_t)
    assert torch.equal(ret, adv_t + v)


def test_gae_rejects_cpu_tensors(lib):
    from cleanrl_b200 import ops
    r, v, d, nv, nd = _gae_inputs(4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gae(r, v, d, nv, nd, 0.99, 0.95)


@pytest.mark.parametrize("n,A", [(1024, 4), (32768, 4), (7, 2), (1000, 18), (1, 6), (257, 64)])
def test_categorical_sample_vs_oracle(lib, n, A):
    from cleanrl_b200 import ops
    g = torch.Generator().manual_seed(n + A)
    logits = torch.randn(n, A, generator=g) * 2.0
    q = torch.empty(n, A).exponential_(1, generator=g)
    val = torch.randn(n, generator=g)
    a_o, lp_o, ent_o = O.categorical_sample(logits.numpy(), q.numpy())
    a, lp, ent, v = ops.categorical_sample(logits.cuda(), q.cuda(), val.cuda())
    assert np.array_equal(a.cpu().numpy(), a_o), "sampled actions must be bit-exact"
    assert np.abs(lp.cpu().numpy() - lp_o).max() <= TOL
    assert np.abs(ent.cpu().numpy() - ent_o).max() <= TOL
    assert torch.equal(v.cpu(), val)


def test_categorical_sample_matches_torch_multinomial_on_device(lib):
    """Same CUDA generator state => same actions as Categorical(logits).sample() on the GPU."""
    from torch.distributions import Categorical
    from cleanrl_b200 import ops
    n, A = 65536, 4
    logits = (torch.randn(n, A) * 1.5).cuda()
    torch.manual_seed(123)
    c = Categorical(logits=logits)
    a_t = c.sample()
    torch.manual_seed(123)
    q = torch.empty(n, A, device="cuda").exponential_(1)
    a, lp, ent, _ = ops.categorical_sample(logits, q)
    assert torch.equal(a, a_t)
    assert (lp - c.log_prob(a_t)).abs().max().item() <= TOL
    assert (ent - c.entropy()).abs().max().item() <= TOL


def test_categorical_extreme_logits(lib):
    from cleanrl_b200 import ops
    logits = torch.tensor([[1000.0, -1000.0, 0.0, 0.0], [-50.0, -50.0, -50.0, -50.0], [0.0, 88.0, -88.0, 3.0]])
    q = torch.ones(3, 4)
    a_o, lp_o, ent_o = O.categorical_sample(logits.numpy(), q.numpy())
    a, lp, ent, _ = ops.categorical_sample(logits.cuda(), q.cuda())
    assert np.array_equal(a.cpu().numpy(), a_o)
    assert np.allclose(lp.cpu().numpy(), lp_o, atol=TOL)
    assert np.allclose(ent.cpu().numpy(), ent_o, atol=TOL)
    assert np.isfinite(ent.cpu().numpy()).all()


def _loss_inputs(M, A, B, seed):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(M, A, generator=g)
    nv = torch.randn(M, generator=g)
    b_act = torch.randint(0, A, (B,), generator=g)
    b_lp = torch.randn(B, generator=g) * 0.1 - 1.3
    b_adv = torch.randn(B, generator=g) * 2
    b_ret = torch.randn(B, generator=g)
    b_val = b_ret + 0.3 * torch.randn(B, generator=g)
    inds = torch.randperm(B, generator=g)[:M]
    return logits, nv, inds, b_act, b_lp, b_adv, b_ret, b_val


@pytest.mark.parametrize("M,A,B", [(32768, 4, 131072), (128, 2, 512), (64, 4, 256), (1000, 18, 1000), (2, 3, 5)])
@pytest.mark.parametrize("norm_adv,clip_vloss", [(True, True), (False, True), (True, False)])
def test_ppo_loss_vs_oracle(lib, M, A, B, norm_adv, clip_vloss):
    from cleanrl_b200 import ops
    t = _loss_inputs(M, A, B, seed=M + A)
    clip, entc, vfc = 0.1, 0.01, 0.5
    st_o, dl_o, dv_o = O.ppo_loss(*[x.numpy() for x in t], clip, entc, vfc, norm_adv, clip_vloss)
    c = [x.cuda() for x in t]
    st, dl, dv = ops.ppo_loss(*c, clip, entc, vfc, norm_adv, clip_vloss)
    st = st.cpu().numpy()
    for i, k in enumerate(ops.STAT_NAMES):
        assert abs(st[i] - float(st_o[k])) <= TOL * max(1.0, abs(float(st_o[k]))), (k, st[i], st_o[k])
    # gradients are O(1/M); compare relative to their own scale
    assert np.abs(dl.cpu().numpy() - dl_o).max() <= TOL * max(np.abs(dl_o).max(), 1e-30) + 1e-12
    assert np.abs(dv.cpu().numpy() - dv_o).max() <= TOL * max(np.abs(dv_o).max(), 1e-30) + 1e-12


def test_ppo_loss_identity_indices_and_clip_edges(lib):
    """ratio exactly 1 everywhere (first epoch): pg ties split, approx_kl == 0, clipfrac == 0."""
    from cleanrl_b200 import ops
    M, A = 512, 4
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(M, A, generator=g)
    act = torch.randint(0, A, (M,), generator=g)
    lp_o, _ = O.categorical_eval(logits.numpy(), act.numpy())
    nv = torch.randn(M, generator=g)
    b_adv = torch.randn(M, generator=g)
    b_ret = torch.randn(M, generator=g)
    # evaluate logprobs on device first so that ratio is exactly 1 on device
    _, lp_dev, _, _ = None, None, None, None
    st_o, dl_o, dv_o = O.ppo_loss(logits.numpy(), nv.numpy(), None, act.numpy(), lp_o, b_adv.numpy(), b_ret.numpy(),
                                  nv.numpy(), 0.2, 0.01, 0.5)
    st, dl, dv = ops.ppo_loss(logits.cuda(), nv.cuda(), None, act.cuda(), torch.from_numpy(lp_o).cuda(),
                              b_adv.cuda(), b_ret.cuda(), nv.cuda(), 0.2, 0.01, 0.5)
    st = st.cpu().numpy()
    assert abs(st[5]) == 0.0  # clipfrac
    assert abs(st[4]) <= 1e-6  # approx_kl
    assert np.abs(dl.cpu().numpy() - dl_o).max() <= 1e-5 * np.abs(dl_o).max() + 1e-9


@pytest.mark.parametrize("P", [1686693, 9155, 11085, 1, 3, 4, 1023])
@pytest.mark.parametrize("world", [1, 2, 8])
def test_clip_adam_vs_oracle(lib, P, world):
    from cleanrl_b200 import ops
    g = torch.Generator().manual_seed(P)
    p = torch.randn(P, generator=g)
    m = torch.zeros(P)
    v = torch.zeros(P)
    pc, mc, vc = p.cuda(), m.cuda(), v.cuda()
    pn, mn, vn = p.numpy(), m.numpy(), v.numpy()
    norm_dev = torch.zeros(1, device="cuda")
    for step in range(1, 5):
        gr = torch.randn(P, generator=g) * (0.02 * step) * world
        lr = O.anneal_lr(step, 10, 2.5e-4)
        pn, mn, vn, tn = O.clip_adam(pn, gr.numpy(), mn, vn, step, lr, eps=1e-5, max_norm=0.5, world_size=world)
        ops.clip_adam(pc, gr.cuda(), mc, vc, step, lr, eps=1e-5, max_norm=0.5, world_size=world, norm_out=norm_dev)
        assert abs(norm_dev.item() - tn) <= 1e-5 * max(1.0, tn)
        assert np.abs(pc.cpu().numpy() - pn).max() <= TOL * max(1.0, np.abs(pn).max())
        assert np.abs(mc.cpu().numpy() - mn).max() <= TOL * max(np.abs(mn).max(), 1e-12)
        assert np.abs(vc.cpu().numpy() - vn).max() <= TOL * max(np.abs(vn).max(), 1e-12)


def test_clip_adam_matches_torch_optimizer_on_device(lib):
    """clip_grad_norm_ + torch.optim.Adam (the reference's calls) on the GPU vs our fused kernel."""
    from cleanrl_b200 import ops
    P = 200003
    p0 = torch.randn(P, device="cuda")
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pt], lr=2.5e-4, eps=1e-5)
    pc, mc, vc = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 6):
        gr = torch.randn(P, device="cuda") * 0.01 * step
        pt.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([pt], 0.5)
        opt.step()
        ops.clip_adam(pc, gr, mc, vc, step, 2.5e-4, eps=1e-5, max_norm=0.5)
        assert (pc - pt.detach()).abs().max().item() <= TOL
