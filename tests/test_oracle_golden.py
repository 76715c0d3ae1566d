"""-m "not gpu": the CPU oracle against the committed golden vectors (generated from the unmodified
reference by tests/golden/make_golden.py) and against the anchors recorded in SURVEY.md §8(c)."""
import os

import numpy as np
import pytest

from tests import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# (state, start_s, index sequence, DP cost) recorded from the reference's st_cy
ANCHORS = [
    (((-200.0, 21.7), 15.0, 0.0, [], [], []), -150.43274377608088,
     [0, 91, 184, 279, 376, 475, 576, 679, 784, 891, 1000, 1111, 1224, 1339, 1456, 1575, 1696, 1819], 1550.107167368958),
    (((-100.0, 8.3), 12.0, 0.5, [-20, -45, -70], [11, 11, 11], [0, 0, 0]), -49.538938220353494,
     [0, 74, 150, 228, 308, 390, 474, 560, 648, 738, 830, 924, 1020, 1118, 1218, 1320, 1424, 1530], 2196.073112911568),
    (((10.0, -1.6), 11.0, 0.0, [30, -8, -30], [11, 11, 11], [0, 0, 0]), 61.0,
     [0, 67, 135, 204, 274, 346, 420, 496, 574, 653, 733, 814, 896, 979, 1063, 1148, 1234, 1321], 2689.9431438858037),
]


@pytest.mark.parametrize("layered", [False, True])
@pytest.mark.parametrize("anchor", range(3))
def test_survey_anchors(oracle, anchor, layered):
    (pos, v, a, xs, vs, acs), s0, idx, cost = ANCHORS[anchor]
    p = oracle.default_params()
    r = oracle.plan(p, oracle.make_state(pos, v, a, xs, vs, acs), layered=layered)
    assert r["start_s"] == s0
    assert r["idx"].tolist() == idx
    assert abs(r["cost"] - cost) <= 1e-9 * cost


def _load(name):
    return dict(np.load(os.path.join(GOLD, name)))


@pytest.mark.parametrize("name,H", [("plan_h17.npz", 17), ("plan_h50.npz", 50), ("plan_h25.npz", 25), ("plan_h50b.npz", 50),
                                    ("plan_h100.npz", 100)])
def test_plan_matches_reference(oracle, name, H):
    G = _load(name)
    p = oracle.horizon_params(H)
    B = G["ego"].shape[0]
    stride = int(G["sample_stride"])
    for layered in ((False, True) if H <= 50 and B <= 40 else (True,)):     # (the Dijkstra port takes ~1 s per state at H=50, 4 s at H=100)
        r = helpers.oracle_plan_batch(oracle, p, G, H + 1, layered=layered)
        assert np.array_equal(r["s_seq"], G["s_seq"])                      # bit-identical positions, incl. zero tails
        assert np.array_equal(r["crash"], G["crash"])
        # the golden cost is summed with the Python twin st.cost (st.py:140-144) whose penalty folds D_WEIGHT in a
        # different order than st_cy (st.py:127-131 vs st_cy.pyx:34-38,50): equal to rounding only
        assert np.all(helpers.rel(r["cost"], G["cost"])[G["cost"] > 0] < 1e-12)
    for b in range(B):
        st = helpers.oracle_state(oracle, G, b)
        ob, di, sv = oracle.build_grid(p, st)
        assert sv[0] == G["start_s"][b] and sv.size == G["num_s"][b] and sv[1] - sv[0] == G["delta_s"][b]
        assert np.array_equal(ob.sum(1), G["obs_count"][b])
        assert np.array_equal(di[:, ::stride][:, :49], G["dist_samples"][b])
        if b < G["masks"].shape[0]:
            assert np.array_equal(np.packbits(ob.astype(bool), axis=1), G["masks"][b])


def test_dijkstra_equals_layered_dp(oracle):
    """The equivalence the CUDA kernels rely on: same path, same cost, every state."""
    from rl_mpc_lanemerging_b200 import synthetic
    p = oracle.default_params()
    for traffic, kind in (("moderate", "mixed"), ("default", "onramp"), ("fast", "mixed")):
        S = synthetic.make_states(48, traffic, seed=21, kind=kind)
        a = helpers.oracle_plan_batch(oracle, p, S, 18, layered=False)
        b = helpers.oracle_plan_batch(oracle, p, S, 18, layered=True)
        assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["cost"], b["cost"])
        assert np.array_equal(a["reached_t"], b["reached_t"])


def test_rollout_pieces_match_reference(oracle):
    G = _load("rollout.npz")
    p = oracle.default_params()
    B = G["ego"].shape[0]
    for b in range(B):
        st = helpers.oracle_state(oracle, G, b)
        n = st.n
        nxt, crashed = oracle.predict_step_with_ego(p, st, G["sel"][b], 0.2, 5.1)
        # the reference normalises the ramp direction with numpy.linalg.norm (BLAS dot): allow 1 ulp on the ego position
        assert np.allclose([nxt.ego_x, nxt.ego_y], G["with_ego"][b, :2], rtol=4e-16, atol=0)
        assert nxt.ego_v == G["with_ego"][b, 2] and nxt.ego_a == G["with_ego"][b, 3]
        assert np.array_equal(nxt.x[:n], G["with_x"][b, :n]) and np.array_equal(nxt.v[:n], G["with_v"][b, :n])
        assert np.array_equal(nxt.a[:n], G["with_a"][b, :n])
        assert crashed == bool(G["with_crash"][b])
        cur = st
        for _ in range(17):
            cur, _c = oracle.predict_step_without_ego(p, cur, 0.3)
        assert np.allclose([cur.ego_x, cur.ego_y], G["chain_ego"][b, :2], rtol=1e-14, atol=0)
        assert np.array_equal(cur.x[:n], G["chain_x"][b, :n]) and np.array_equal(cur.v[:n], G["chain_v"][b, :n])
        assert np.array_equal(oracle.state_vector(p, st), G["state_vec"][b])
        assert oracle.speed_from_jerk(p, st.ego_v, st.ego_a, G["jerk"][b]) == G["speed"][b]
        assert oracle.path_mean_abs_jerk(G["jerk_paths"][b], st.ego_v, st.ego_a, 0.2) == G["mean_abs_jerk"][b]


def test_cost_bound_is_exact_when_the_horizon_is_reached(oracle):
    """The fast kernel's first attempt drops every node whose label exceeds a bound (mpc_fast.cu).  In the CPU model of
    that kernel: whenever the bounded pass reaches the horizon its path and cost ARE the unbounded ones -- for the
    kernel's own bound (one step inside a penalty zone), for a bound just above the optimum, and it never reaches the
    horizon with a bound below the optimum."""
    G = dict(np.load(os.path.join(GOLD, "plan_h17.npz")))
    p = oracle.default_params()
    zone = p.d_weight * 1e6 / max(p.min_allowed_distance, 1.0)
    reached_bounded = pruned = 0
    for b in range(0, G["ego"].shape[0], 3):
        stt = helpers.oracle_state(oracle, G, b)
        ob, di, sv = oracle.build_grid(p, stt)
        full = oracle.solve_fast_model_ex(p, ob, di, sv, p.t_disc, stt.ego_v, stt.ego_a)
        feasible = full["reached_t"] == ob.shape[0] - 1
        for bound in (zone, full["cost"] + 1e-3, full["cost"] * 0.99):
            r = oracle.solve_fast_model_ex(p, ob, di, sv, p.t_disc, stt.ego_v, stt.ego_a, prune_cost=bound)
            if r["reached_t"] == ob.shape[0] - 1:
                assert feasible and bound >= full["cost"]
                assert np.array_equal(r["idx"], full["idx"]) and r["cost"] == full["cost"]
                reached_bounded += 1
                pruned += r["nodes"] < full["nodes"]
            else:
                assert (not feasible) or bound < full["cost"]
    assert reached_bounded > 40 and pruned > 20


def test_blocked_interval_is_band_plus_penalty_zones():
    """The lean bounded pass of the fast kernel blocks, per car, ONE interval of cells: from the first cell closer than
    MIN_ALLOWED_DISTANCE to the car's rear edge to the last cell closer than that to its front edge, joined with the band
    (mpc_predict.cu).  Claim behind `zone_ok` (mpc_api.cu): for m >= 4 cells that interval is exactly 'in the band, or
    min(|s - ef|, |s - eb|) < m' -- checked here in numpy on random geometry, with and without prediction uncertainty."""
    rng = np.random.default_rng(0)
    for ds, L, m, unc_max in ((0.05, 5.0, 5.0, 0.0), (0.05, 5.0, 5.0, 2.0), (0.05, 5.0, 0.2, 2.0), (0.05, 4.0, 3.0, 1.0),
                              (0.05, 5.03, 0.2, 1.0), (0.1, 5.0, 6.0, 1.0)):
        for _ in range(150):
            s0, num_s = rng.uniform(-200, 80), 3001
            s = s0 + np.arange(num_s) * ((s0 + ds) - s0)
            obs, unc = rng.uniform(s0 - 20, s[-1] + 12), rng.uniform(0, unc_max)
            du, ef, eb = int(unc / ds), obs - L - unc, obs + L + unc
            si, dl = int((obs - s0) / ds), int(L / ds)
            imin, imax = max(si - dl - du, 0), min(si + dl + du, num_s)
            if not (imin < num_s and imax > 0):
                imin = imax = 0
            ref = np.minimum(np.abs(s - ef), np.abs(s - eb)) < m
            ref[imin:imax] = True
            out_lo, out_hi = (s < ef) & (np.abs(s - ef) >= m), (s > eb) & (np.abs(s - eb) >= m)
            zlo = int(np.argmin(out_lo)) if not out_lo.all() else num_s
            zhi = int(num_s - np.argmin(out_hi[::-1])) if not out_hi.all() else 0
            if imin < imax:
                zlo, zhi = min(zlo, imin), max(zhi, imax)
            hull = np.zeros(num_s, bool)
            hull[zlo:max(zhi, zlo)] = True
            assert np.array_equal(hull, ref), (ds, L, m, unc, s0, obs)
