"""Geometry and actuation helpers with the reference's names (reference control.py:160-171, 366-389).

The SUMO/traci episode loop of the reference (control.run_episode) has no counterpart here: episodes are
stepped in batches by merge_gym.MergeEnv.  Scalars and tensors are both accepted; tensor inputs stay on the
device (K4 kernels), scalars are evaluated through the same kernels as a batch of one.
"""
from __future__ import annotations

import math

import torch

from .config import Settings
from .prediction import tdiv

merge_point = (-50.9, 1.72)
merge_point2 = (1.5, -1.5)
merge_point3 = (-51, -1.5)
common_s = merge_point2[0] - merge_point3[0]


def distance(point1, point2):
    return math.sqrt((point1[0] - point2[0]) ** 2 + (point1[1] - point2[1]) ** 2)


merge_distance = distance(merge_point, merge_point2)


def get_ego_s(ego_position):
    """(x, y) -> arclength along ramp + highway (host scalar; the kernels carry their own fp64 copy,
    csrc/mpc_common.cuh get_ego_s, which is what every plan / grid uses)."""
    x, _y = ego_position
    if x < merge_point[0]:
        return -distance(ego_position, merge_point)
    if x < merge_point2[0]:
        return distance(ego_position, merge_point)
    return x - merge_point2[0] + common_s


def get_obstacle_s_from_x(vehicle_x):
    return vehicle_x - merge_point3[0]


def get_obstacle_s(vehicle_position):
    return get_obstacle_s_from_x(vehicle_position[0])


def get_ego_speed_from_jerk(current_speed, current_acceleration, jerk):
    """Speed after one tick of constant jerk, with the acceleration and speed clamps (reference control.py:160-171).
    Tensors [B] -> tensor [B] on the same device; scalars -> float.  Runs in the K4 kernel either way."""
    from . import st
    scalar = not torch.is_tensor(current_speed)
    eng = st.get_engine(1 if scalar else int(current_speed.shape[0]))
    f = lambda x: torch.as_tensor(x, dtype=torch.float64, device=eng.device).reshape(-1)  # noqa: E731
    v, a, j = f(current_speed), f(current_acceleration), f(jerk)
    ego = torch.zeros((v.shape[0], 4), dtype=torch.float64, device=eng.device)
    ego[:, 2], ego[:, 3] = v, a
    out = eng.speed_from_jerk(ego, j.contiguous())
    return float(out.item()) if scalar else out


def get_mpc_action(state):
    """north_star alias: the MPC action for a state = st.do_st_control (reference st.py:757-783)."""
    from . import st
    return st.do_st_control(state)


def get_ego_start_speed(n=None, generator=None):
    """Reference control.py:198-204: N(START_SPEED, START_SPEED_VARIANCE) clipped to [MIN, MAX]."""
    import numpy as np
    rng = generator or np.random
    if not Settings.RANDOMIZE_START_SPEED:
        return Settings.START_SPEED if n is None else np.full(n, float(Settings.START_SPEED))
    v = rng.normal(Settings.START_SPEED, Settings.START_SPEED_VARIANCE, size=n)
    return np.clip(v, Settings.MIN_START_SPEED, Settings.MAX_START_SPEED)


# ---- batched episode loop: reference control.py:229-363 (run_episode / evaluate_control) --------------------------
class EpisodeTracker:
    """Per-episode running reductions of the metrics reference run_episode records tick by tick (control.py:269-318):
    speed, |jerk| from the acceleration differences, distance to the closest vehicle once on the highway, deceleration
    forced on the car behind ("disruption"), per-segment histograms, planner take-overs.  All on the device."""

    def __init__(self, num_envs: int, device):
        f = dict(dtype=torch.float64, device=device)
        self.B, self.device = num_envs, device
        names = ("steps", "sum_speed", "max_speed", "sum_abs_jerk", "prev_acc", "min_closest", "sum_closest", "n_closest",
                 "sum_disr", "max_disr", "n_disr", "nz_disr", "takeovers")
        self.v = {n: torch.zeros(num_envs, **f) for n in names}
        self.v["min_closest"].fill_(float("inf"))
        self.bins = torch.arange(-220, 61, 20, **f)
        nb = len(self.bins) - 1
        self.seg = {n: torch.zeros((num_envs, nb), **f) for n in ("counts", "jerks", "speeds")}     # per running episode
        self.seg_counts = torch.zeros(nb, **f)                                                       # finished episodes only
        self.seg_jerks, self.seg_speeds = torch.zeros_like(self.seg_counts), torch.zeros_like(self.seg_counts)

    def reset_where(self, mask):
        for n, t in self.v.items():
            t.masked_fill_(mask, float("inf") if n == "min_closest" else 0.0)
        for t in self.seg.values():
            t.masked_fill_(mask.unsqueeze(1), 0.0)

    def record(self, state, takeover=None):
        """Called with the state the controller sees (before the tick), like the reference loop."""
        from . import dqn
        S, V, tick = Settings, self.v, float(Settings.TICK_LENGTH)
        ex, speed, acc = state.ego[:, 0], state.ego[:, 2], state.ego[:, 3]
        jerk = torch.where(V["steps"] > 0, (acc - V["prev_acc"]) / tick, torch.zeros_like(acc))       # control.py:286-289
        V["steps"] += 1; V["sum_speed"] += speed; V["max_speed"] = torch.maximum(V["max_speed"], speed)
        V["sum_abs_jerk"] += jerk.abs(); V["prev_acc"] = acc.clone()
        ego_s = dqn._ego_s(state.ego)
        valid = torch.arange(state.cars_x.shape[1], device=self.device).unsqueeze(0) < state.n_cars.unsqueeze(1)
        dx = state.cars_x - ex.unsqueeze(1)
        inf = torch.full_like(dx, float("inf"))
        ahead = torch.where(valid & (dx >= 0), dx, inf).min(1).values                                   # prediction.py:162-182
        behind_d, behind_i = torch.where(valid & (dx < 0), -dx, inf).min(1)
        behind_acc = state.cars_a.gather(1, behind_i.unsqueeze(1)).squeeze(1)
        on_highway = ego_s > float(S.MERGE_POINT_X)                                                    # control.py:292
        min_d = torch.minimum(torch.minimum(ahead, behind_d), torch.full_like(ahead, 100.0))           # 305
        m_close = on_highway & (ego_s > float(S.CRASH_MIN_S))                                          # 306-307
        V["min_closest"] = torch.where(m_close, torch.minimum(V["min_closest"], min_d), V["min_closest"])
        V["sum_closest"] += torch.where(m_close, min_d, torch.zeros_like(min_d)); V["n_closest"] += m_close.double()
        disr = torch.where(on_highway & torch.isfinite(behind_d), (-behind_acc).clamp(min=0), torch.zeros_like(behind_acc))   # 300-304, 308
        V["sum_disr"] += disr; V["max_disr"] = torch.maximum(V["max_disr"], disr)
        V["n_disr"] += on_highway.double(); V["nz_disr"] += (disr != 0).double()
        if takeover is not None:
            V["takeovers"] += takeover.double()
        seg = torch.bucketize(ex.contiguous(), self.bins, right=False) - 1                                          # stats.py:44-52
        ok = (seg >= 0) & (seg < self.seg_counts.numel())
        seg = seg.clamp(0, self.seg_counts.numel() - 1).unsqueeze(1)
        self.seg["counts"].scatter_add_(1, seg, ok.double().unsqueeze(1))
        self.seg["jerks"].scatter_add_(1, seg, (jerk.abs() * ok).unsqueeze(1))
        self.seg["speeds"].scatter_add_(1, seg, (speed.abs() * ok).unsqueeze(1))

    def finished(self, done, crashed, merged, wall_per_env_tick: float):
        """Episode dicts (host floats) for the rows where `done`."""
        idx = done.nonzero().squeeze(1)
        if idx.numel() == 0:
            return []
        self.seg_counts += self.seg["counts"][idx].sum(0); self.seg_jerks += self.seg["jerks"][idx].sum(0)
        self.seg_speeds += self.seg["speeds"][idx].sum(0)                 # the histograms cover finished episodes (stats.py:43-52)
        tick = float(Settings.TICK_LENGTH)
        rows = {n: t[idx].cpu().numpy() for n, t in self.v.items()}
        cr, mg, slots = crashed[idx].cpu().numpy(), merged[idx].cpu().numpy(), idx.cpu().numpy()
        out = []
        for i in range(idx.numel()):
            n = max(rows["steps"][i], 1.0)
            out.append(dict(
                slot=int(slots[i]), crashed=bool(cr[i]), merged=bool(mg[i]), mean_speed=rows["sum_speed"][i] / n, max_speed=rows["max_speed"][i],
                mean_abs_jerk=rows["sum_abs_jerk"][i] / n, time_taken=rows["steps"][i] * tick,
                clock_time_per_episode=wall_per_env_tick * n, clock_time_per_step=wall_per_env_tick,
                n_closest=int(rows["n_closest"][i]), closest_distance=rows["min_closest"][i],
                mean_closest_distance=rows["sum_closest"][i] / max(rows["n_closest"][i], 1.0),
                n_disruption=int(rows["n_disr"][i]), mean_disruption=rows["sum_disr"][i] / max(rows["n_disr"][i], 1.0),
                max_disruption=rows["max_disr"][i], total_disruption=rows["sum_disr"][i] * tick,
                disruption_time=rows["nz_disr"][i] * tick, steps=int(rows["steps"][i]), takeovers=int(rows["takeovers"][i])))
        return out


def evaluate_control(control_function, num_episodes=1000, state_function=None, custom_stats_function=None,
                     end_episode_callback=None, max_episode_length=100, start_velocity=None, wait_before_start=50,
                     save_state_on_crash=False, verbose=False, crash_callback=None, num_envs=None, seed=None, env=None):
    """Run `num_episodes` lane-merging episodes under `control_function` and aggregate the reference's statistics
    (reference control.py:343-363).  Episodes run `num_envs` at a time in merge_gym.MergeEnv.

    control_function(BatchedState) -> commanded speed [B], or (speed [B], takeover mask [B]) for the combined controller.
    end_episode_callback(done_mask) is called after every tick that finished episodes.  state_function,
    wait_before_start, start_velocity, crash_callback exist for signature parity (the batched world has no SUMO warm-up).
    clock_time_per_step is the wall time of a batched tick divided by the number of environments.  `env` may be any object
    with MergeEnv's interface (state, device, B, reset(), step(jerk)).

    Which episodes are counted does not depend on how long they last: environment slot b contributes its j-th episode iff
    j * B + b < num_episodes, i.e. every slot a fixed quota (the reference runs its episodes one after another and counts all of
    them; counting "the first num_episodes that finish" would over-represent short episodes -- early crashes, quick merges --
    and drop the slow tail of jams and time-outs).  seed=None takes Settings.seed_value()."""
    import time
    from . import merge_gym, stats
    B = int(num_envs or min(max(num_episodes, 1), 4096))
    Settings.MAX_EPISODE_LENGTH = max_episode_length
    if env is None:
        env = merge_gym.MergeEnv(B, seed=Settings.seed_value() if seed is None else seed)
    B = env.B
    agg = stats.StatsAggregator(save_state_on_crash)
    if custom_stats_function is not None:
        agg.add_custom_stat_callback(custom_stats_function)
    tracker = EpisodeTracker(B, env.device)
    slot_episodes = [0] * B                                    # episodes slot b has finished so far
    env.reset()
    tick = float(Settings.TICK_LENGTH)
    t_prev, wall = time.perf_counter(), 0.0
    while agg.episodes < num_episodes:
        out = control_function(env.state)
        speed, takeover = out if isinstance(out, tuple) else (out, None)
        tracker.record(env.state, takeover)
        jerk = tdiv(tdiv(speed - env.state.ego[:, 2], tick) - env.state.ego[:, 3], tick)
        _obs, _reward, done, info = env.step(jerk)
        if bool(done.any()):                                   # (host sync: evaluation bookkeeping, not the hot path)
            now = time.perf_counter(); wall = 0.9 * wall + 0.1 * (now - t_prev) if wall else now - t_prev
            for ep in tracker.finished(done, info["crashed"], info["merged"], wall / B):
                b = ep.pop("slot")
                j, slot_episodes[b] = slot_episodes[b], slot_episodes[b] + 1
                if j * B + b < num_episodes:
                    agg.add_episode_stats(ep)
                    if verbose and ep["crashed"]:
                        print("crashed")
            tracker.reset_where(done)
            if end_episode_callback is not None:
                end_episode_callback(done)
        t_prev = time.perf_counter()
    agg.add_segment_histograms(tracker.seg_counts.cpu().numpy(), tracker.seg_jerks.cpu().numpy(), tracker.seg_speeds.cpu().numpy())
    return agg
