"""-m gpu: the device path against outputs of the UNMODIFIED reference (tests/golden/*.npz, generated in the build container by
make_golden.py / make_golden_r2.py) at every horizon of BASELINE.json's sweep: H = 17, 25, 50, 100."""
import os

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,H", [("plan_h17.npz", 17), ("plan_h25.npz", 25), ("plan_h50.npz", 50), ("plan_h50b.npz", 50),
                                    ("plan_h100.npz", 100)])
def test_plans_equal_the_reference(oracle, name, H):
    """Exact mode: positions, crash verdict and start_s bit for bit, cost to rounding (the golden cost is summed by the Python
    twin st.cost).  Fast mode (32-bit-key kernel first, 64-bit kernel for what it hands on): the same positions on every state
    of these files, cost within 1e-6; at H=100 the 32-bit-key kernel runs with one block per SM and 2^-16 labels."""
    import torch
    from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
    G = dict(np.load(os.path.join(GOLD, name)))
    B = G["ego"].shape[0]
    op = oracle.horizon_params(H)
    eng = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=B)
    try:
        D = states_to_device({k: G[k] for k in ("ego", "cars_x", "cars_v", "cars_a", "n_cars")}, "cuda:0")
        a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
        ex = {k: v.cpu().numpy() for k, v in eng.plan(*a, mode="exact").items()}
        torch.cuda.synchronize()
        assert np.array_equal(ex["s_seq"], G["s_seq"]) and np.array_equal(ex["crash"].astype(bool), G["crash"])
        assert np.array_equal(ex["start_s"], G["start_s"])
        ok = G["cost"] > 0
        assert np.all(helpers.rel(ex["cost"][ok], G["cost"][ok]) < 1e-12)
        fa = {k: v.cpu().numpy() for k, v in eng.plan(*a, mode="fast").items()}
        info = eng.fast32_info()
        assert info["in_use"] and info["frac_bits"] == {17: 18, 25: 18, 50: 17, 100: 16}[H]
        assert np.array_equal(fa["s_seq"], G["s_seq"]) and np.array_equal(fa["crash"].astype(bool), G["crash"])
        assert np.all(helpers.rel(fa["cost"][ok], G["cost"][ok]) < 1e-6)
    finally:
        eng.close()
