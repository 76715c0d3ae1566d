"""mpc_predict_step_without_ego (reference prediction.py:22-44) against the CPU oracle: bit-exact, like the other K4 pieces.
  (Passed on a B200 at the end of round 1.)"""
import numpy as np
import pytest

from rl_mpc_lanemerging_b200 import synthetic
from tests import helpers

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("traffic,kind,dt,mcd", [("moderate", "mixed", 0.3, 5.0), ("default", "mixed", 0.2, 5.1),
                                                  ("fast", "onramp", 0.3, 7.5)])
def test_predict_step_without_ego_matches_oracle(oracle, traffic, kind, dt, mcd):
    import torch
    from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
    B = 256
    op = oracle.default_params()
    eng = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=B)
    S = synthetic.make_states(B, traffic, seed=21, kind=kind)
    S["n_cars"][:4] = 0                                       # the "no cars" branch (prediction.py:26)
    S["ego"][5] = (60.0, -1.6, 12.0, 0.0)                     # the ego leads every car (prediction.py:28-31)
    S["cars_x"][5, :3] = (40.0, 20.0, 0.0); S["cars_v"][5, :3] = 11.0; S["cars_a"][5, :3] = 0.0; S["n_cars"][5] = 3
    D = states_to_device(S, "cuda:0")
    eo, xo, vo, ao, cr = eng.predict_step_without_ego(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], dt, mcd)
    torch.cuda.synchronize()
    eo, xo, vo, ao, cr = (t.cpu().numpy() for t in (eo, xo, vo, ao, cr))
    branches = set()
    for b in range(B):
        st = helpers.oracle_state(oracle, S, b)
        out, crashed = oracle.predict_step_without_ego(op, st, dt, mcd)
        n = st.n
        assert (eo[b, 0], eo[b, 1], eo[b, 2], eo[b, 3]) == (out.ego_x, out.ego_y, out.ego_v, out.ego_a), b
        assert np.array_equal(xo[b, :n], np.array(out.x[:n])) and np.array_equal(vo[b, :n], np.array(out.v[:n])), b
        assert np.array_equal(ao[b, :n], np.array(out.a[:n])), b
        assert bool(cr[b]) == crashed, b
        es = oracle.get_ego_s(st.ego_x, st.ego_y)
        branches.add("stay" if (es < 8 or n == 0) else ("lead" if st.x[0] < st.ego_x else "follow"))
    if kind == "mixed":
        assert branches == {"stay", "lead", "follow"}
    eng.close()


def test_highway_state_method(oracle):
    from rl_mpc_lanemerging_b200.prediction import HighwayState
    op = oracle.default_params()
    s = HighwayState((30.0, -1.6), 12.0, 0.5, [60.0, 41.0, 18.0, -5.0], [11.0, 10.0, 11.0, 12.0], [0.0, 0.0, 0.0, 0.0])
    out, crashed = s.predict_step_without_ego(0.3)
    ref, rc = oracle.predict_step_without_ego(op, oracle.make_state((30.0, -1.6), 12.0, 0.5, s.other_xs, s.other_speeds,
                                                                    s.other_accelerations), 0.3, 5.0)
    assert tuple(out.ego_position) == (ref.ego_x, ref.ego_y) and out.ego_speed == ref.ego_v and crashed == rc
    assert list(out.other_xs) == list(ref.x[:4]) and list(out.other_speeds) == list(ref.v[:4])
