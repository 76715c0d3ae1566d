"""-m gpu: the batched finer_fit kernel (reference st.py:584-723).  cvxopt is not available, so the check is the QP
itself: feasibility, agreement with an independent CPU solve (scipy SLSQP) of the same problem, objective optimality."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
DT, DTC = 0.2, 0.3
LIM = dict(vmax=30.0, amin=-6.0, amax=4.5, jmin=-5.0, jmax=5.0)


def _rows(x, v0, a0):
    v = np.diff(x) / DT
    a = np.diff(np.concatenate([[v0], v])) / DT
    j = np.diff(np.concatenate([[a0], a])) / DT
    return v, a, j


def _cpu_qp(sc, v0, a0):
    from scipy.optimize import minimize
    L = len(sc)
    t = np.arange(L) * DTC
    n = int(np.round(t[-1] / DT + 1))
    if (n - 1) * DT > t[-1]:
        n -= 1
    b = np.interp(np.arange(n) * DT, t, sc)

    def cons(u):
        v, a, j = _rows(np.concatenate([[sc[0]], u]), v0, a0)
        return np.concatenate([v, LIM["vmax"] - v, a - LIM["amin"], LIM["amax"] - a, j - LIM["jmin"], LIM["jmax"] - j])
    r = minimize(lambda u: ((u - b[1:]) ** 2).sum(), b[1:], jac=lambda u: 2 * (u - b[1:]), constraints=[{"type": "ineq", "fun": cons}],
                 method="SLSQP", options=dict(maxiter=2000, ftol=1e-16))
    return np.concatenate([[sc[0]], r.x]), b


def test_finer_fit_solves_the_reference_qp(oracle):
    import torch
    from rl_mpc_lanemerging_b200 import synthetic
    from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
    op = oracle.default_params()
    eng = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=256)
    S = synthetic.make_states(256, "moderate", seed=23, kind="mixed")
    D = states_to_device(S, "cuda:0")
    plan = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode="exact")
    fine, n_fine, speed, iters = eng.finer_fit(plan["s_seq"], plan["reached_t"], D["ego"])
    torch.cuda.synchronize()
    fine, n_fine, speed, iters = fine.cpu().numpy(), n_fine.cpu().numpy(), speed.cpu().numpy(), iters.cpu().numpy()
    seq, reached = plan["s_seq"].cpu().numpy(), plan["reached_t"].cpu().numpy()
    assert iters.max() <= 30
    checked = 0
    for b in range(256):
        L = reached[b] + 1
        v0, a0 = S["ego"][b, 2], S["ego"][b, 3]
        if L == 1:
            assert n_fine[b] == 1 and speed[b] == v0                      # st.py:775-777
            continue
        t_last = (L - 1) * DTC
        n = int(np.round(t_last / DT + 1)); n -= ((n - 1) * DT > t_last)
        assert n_fine[b] == n and (L < 18 or n == 26)
        x = fine[b, :n]
        assert x[0] == seq[b, 0]                                           # equality constraint (st.py:707-711)
        v, a, j = _rows(x, v0, a0)
        eps = 1e-7
        feasible_start = LIM["amin"] - 1 <= a0 <= LIM["amax"] + 1
        if feasible_start:
            assert v.min() >= -eps and v.max() <= LIM["vmax"] + eps
            assert a.min() >= LIM["amin"] - eps and a.max() <= LIM["amax"] + eps
            assert j.min() >= LIM["jmin"] - 1e-5 and j.max() <= LIM["jmax"] + 1e-5
        assert speed[b] == (x[1] - x[0]) / DT                              # st.py:780-781
        if b % 16 == 0:
            xr, bb = _cpu_qp(seq[b, :L], v0, a0)
            assert np.abs(xr - x).max() < 2e-4, (b, np.abs(xr - x).max())        # SLSQP itself is only accurate to ~1e-5 on flat directions
            assert ((x - bb) ** 2).sum() <= ((xr - bb) ** 2).sum() + 1e-7  # at least as good as the CPU solve
            checked += 1
    assert checked >= 10
    eng.close()


def test_do_st_control_and_b_branch_run():
    """st.do_st_control (single state + batch) and the TEST_ST_STRICTLY_BETTER branch of do_combined_control."""
    import torch
    from rl_mpc_lanemerging_b200 import ddpg, st, synthetic
    from rl_mpc_lanemerging_b200.config import Settings
    from rl_mpc_lanemerging_b200.prediction import BatchedState
    Settings.reset(); Settings.CRASH_MIN_S = 20; Settings.OTHER_CAR_SPEED = 11.0
    S = synthetic.make_states(64, "moderate", seed=29, kind="mixed")
    batch = BatchedState.from_numpy(S, "cuda:0")
    v = st.do_st_control(batch)
    assert v.shape == (64,) and torch.isfinite(v).all() and (v >= -1e-9).all() and (v <= 30 + 1e-6).all()
    one = batch.to_states()[3]
    assert abs(st.do_st_control(one) - float(v[3])) < 1e-6                # exact vs fast plan may differ in the last digits only
    path = st.finer_fit(np.array([10.0, 13.0, 16.2, 19.5]), 0.2, 0.3, 10.0, 0.0)
    assert path.shape == (5,) and path[0] == 10.0
    Settings.TEST_ST_STRICTLY_BETTER = True
    agent = ddpg.DDPGAgent(device="cuda:0", seed=1)
    speed, takeover = agent.do_combined_control(batch)
    assert speed.shape == (64,) and torch.isfinite(speed).all() and takeover.dtype == torch.bool
    st.refresh_engine(); Settings.reset()
