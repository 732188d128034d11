"""CPU: the loop port used as cpu_baseline (oracle/ppo_port.py) reproduces the unmodified reference
script's logged losses on the same seed (fixture from oracle/make_golden.py)."""
import numpy as np

from conftest import GOLDEN
from oracle import ppo_port


def test_port_matches_reference_losses():
    z = np.load(GOLDEN / "ppo_atari_envpool_n8_t32_seed1.npz")
    st = None
    import numpy.random as npr
    state = npr.get_state()
    try:
        out = ppo_port.run(num_envs=8, num_steps=32, num_iterations=2, total_iterations=3, seed=1, threads=4)
    finally:
        npr.set_state(state)
    vl = z["tb/losses/value_loss"][:, 1]
    pl = z["tb/losses/policy_loss"][:, 1]
    en = z["tb/losses/entropy"][:, 1]
    for i in range(2):
        assert abs(out["losses"][i]["v_loss"] - vl[i]) <= 1e-4 * max(1.0, abs(vl[i]))
        assert abs(out["losses"][i]["pg_loss"] - pl[i]) <= 1e-4
        assert abs(out["losses"][i]["entropy"] - en[i]) <= 1e-4
