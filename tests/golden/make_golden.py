"""Generates the committed golden vectors by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py          (needs /root/reference; not runnable on the GPU box)

What runs is the reference's own code: st_cy.pyx compiled by oracle/build_ref.py, and st.py /
prediction.py / control.py / dqn.py imported by oracle/ref_harness.py with stub traci / cvxopt /
matplotlib modules.  Inputs are the seeded synthetic states of rl_mpc_lanemerging_b200.synthetic plus
the three anchor states recorded in SURVEY.md §8(c).  Outputs (small .npz files next to this script):

  plan_h17.npz / plan_h50.npz   states, s_sequence, DP cost (sum of st.cost along the path), crash
                                verdict (st.test_guaranteed_crash_from_state), grid pins (start_s, num_s,
                                obstacle count per layer, distance field sampled every 61st cell,
                                packed obstacle masks of the first few states)
  rollout.npz                   predict_step_with_ego / predict_step_without_ego chains, observation
                                vectors, get_ego_speed_from_jerk, get_path_mean_abs_jerk
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from rl_mpc_lanemerging_b200 import synthetic  # noqa: E402

ANCHORS = [
    ((-200.0, 21.7), 15.0, 0.0, [], [], []),
    ((-100.0, 8.3), 12.0, 0.5, [-20.0, -45.0, -70.0], [11.0, 11.0, 11.0], [0.0, 0.0, 0.0]),
    ((10.0, -1.6), 11.0, 0.0, [30.0, -8.0, -30.0], [11.0, 11.0, 11.0], [0.0, 0.0, 0.0]),
]
SAMPLE_STRIDE = 61


def batch_states(per_case, seed, cases):
    parts = [synthetic.make_states(per_case, traffic, seed=seed + i, kind=kind) for i, (traffic, kind) in enumerate(cases)]
    S = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    # prepend the anchors
    A = len(ANCHORS)
    ego = np.zeros((A, 4)); cx = np.zeros((A, synthetic.NMAX)); cv = np.zeros_like(cx); ca = np.zeros_like(cx)
    n = np.zeros(A, np.int32)
    for i, (pos, v, a, xs, vs, acs) in enumerate(ANCHORS):
        ego[i] = [pos[0], pos[1], v, a]; n[i] = len(xs)
        cx[i, :len(xs)] = xs; cv[i, :len(xs)] = vs; ca[i, :len(xs)] = acs
    return dict(ego=np.concatenate([ego, S["ego"]]), cars_x=np.concatenate([cx, S["cars_x"]]),
                cars_v=np.concatenate([cv, S["cars_v"]]), cars_a=np.concatenate([ca, S["cars_a"]]),
                n_cars=np.concatenate([n, S["n_cars"]]))


def ref_state(R, S, b):
    n = int(S["n_cars"][b]); e = S["ego"][b]
    return R.prediction.HighwayState((float(e[0]), float(e[1])), float(e[2]), float(e[3]), S["cars_x"][b, :n].tolist(),
                                     S["cars_v"][b, :n].tolist(), S["cars_a"][b, :n].tolist())


def dp_cost(R, seq, sv, dist, v0, a0):
    """Sum of the reference's st.cost over the path with the solver's history convention (st_cy.pyx:329-330,387)."""
    dt = R.Settings.T_DISCRETIZATION
    s1 = sv[0] - dt * v0
    s2 = s1 - dt * (v0 - a0 * dt)
    s = seq[0]; total = 0
    ds = sv[1] - sv[0]
    for t in range(1, len(seq)):
        if seq[t] == 0:
            break
        k = int(round((seq[t] - sv[0]) / ds))
        assert sv[k] == seq[t]
        total = total + R.st.cost(seq[t], s, s1, s2, dt, dist[t, k])
        s2, s1, s = s1, s, seq[t]
    return total


def make_plan(R, H, S, n_masks, path):
    ft, fs = synthetic.horizon_settings(H)
    R.Settings.FUTURE_T, R.Settings.FUTURE_S = ft, fs
    B = S["ego"].shape[0]
    out = dict(seq=[], cost=[], crash=[], start_s=[], num_s=[], delta_s=[], obs_count=[], dist_samples=[], masks=[])
    for b in range(B):
        st = ref_state(R, S, b)
        seq, obs, sv, tv, dist = R.st.get_appropriate_base_st_path_and_obstacles(st)
        assert tv.size == H + 1
        out["seq"].append(seq); out["cost"].append(dp_cost(R, seq, sv, dist, st.ego_speed, st.ego_acceleration))
        out["crash"].append(bool(R.st.test_guaranteed_crash_from_state(st)))
        out["start_s"].append(sv[0]); out["num_s"].append(sv.size); out["delta_s"].append(sv[1] - sv[0])
        out["obs_count"].append(obs.sum(1)); out["dist_samples"].append(dist[:, ::SAMPLE_STRIDE][:, :49])
        if b < n_masks:
            out["masks"].append(np.packbits(obs[:, :3001] if H == 17 else obs[:, :9001], axis=1))
    np.savez_compressed(path, H=H, sample_stride=SAMPLE_STRIDE, **S,
                        s_seq=np.array(out["seq"]), cost=np.array(out["cost"]), crash=np.array(out["crash"]),
                        start_s=np.array(out["start_s"]), num_s=np.array(out["num_s"], np.int32),
                        delta_s=np.array(out["delta_s"]), obs_count=np.array(out["obs_count"], np.int32),
                        dist_samples=np.array(out["dist_samples"]), masks=np.array(out["masks"]))
    R.Settings.FUTURE_T, R.Settings.FUTURE_S = 5.0, 150.0
    print(path, B, "states; incomplete plans:", int(sum(s[-1] == 0 for s in out["seq"])), "crash:", int(sum(out["crash"])))


def make_rollout(R, S, path):
    B = S["ego"].shape[0]
    rng = np.random.default_rng(7)
    sel = rng.uniform(0, 30, B); jerk = rng.uniform(-6, 6, B)
    N = synthetic.NMAX
    w_ego = np.zeros((B, 4)); w_x = np.zeros((B, N)); w_v = np.zeros((B, N)); w_a = np.zeros((B, N)); w_crash = np.zeros(B, bool)
    c_ego = np.zeros((B, 4)); c_x = np.zeros((B, N)); c_v = np.zeros((B, N))
    vec = np.zeros((B, 20)); spd = np.zeros(B); jrk = np.zeros(B)
    for b in range(B):
        st = ref_state(R, S, b)
        nxt, crashed = st.predict_step_with_ego(sel[b], 0.2, 5.1)
        n = len(nxt.other_xs)
        w_ego[b] = [nxt.ego_position[0], nxt.ego_position[1], nxt.ego_speed, nxt.ego_acceleration]
        w_x[b, :n] = nxt.other_xs; w_v[b, :n] = nxt.other_speeds; w_a[b, :n] = nxt.other_accelerations; w_crash[b] = crashed
        cur = st
        for _ in range(17):
            cur, _c = cur.predict_step_without_ego(0.3)
        c_ego[b] = [cur.ego_position[0], cur.ego_position[1], cur.ego_speed, cur.ego_acceleration]
        c_x[b, :n] = cur.other_xs; c_v[b, :n] = cur.other_speeds
        vec[b] = R.dqn.get_state_vector_from_base_state(st)
        spd[b] = R.control.get_ego_speed_from_jerk(st.ego_speed, st.ego_acceleration, jerk[b])
        path_s = np.cumsum(rng.uniform(0, 6, 12)) + 3.0
        jrk[b] = R.st.get_path_mean_abs_jerk(path_s, st.ego_speed, st.ego_acceleration, 0.2)
        if b == 0:
            paths = np.zeros((B, 12))
        paths[b] = path_s
    np.savez_compressed(path, **S, sel=sel, jerk=jerk, with_ego=w_ego, with_x=w_x, with_v=w_v, with_a=w_a, with_crash=w_crash,
                        chain_ego=c_ego, chain_x=c_x, chain_v=c_v, state_vec=vec, speed=spd, jerk_paths=paths, mean_abs_jerk=jrk)
    print(path, B, "states; crashes predicted:", int(w_crash.sum()))


def main():
    R = ref_harness.load("configs/st_moderate.json")
    cases = [(t, k) for t in ("low", "medium", "default", "moderate", "fast") for k in ("onramp", "mixed")]
    make_plan(R, 17, batch_states(16, 100, cases), 8, os.path.join(HERE, "plan_h17.npz"))
    make_plan(R, 50, batch_states(3, 200, cases), 2, os.path.join(HERE, "plan_h50.npz"))
    make_rollout(R, batch_states(12, 300, cases), os.path.join(HERE, "rollout.npz"))


if __name__ == "__main__":
    main()
