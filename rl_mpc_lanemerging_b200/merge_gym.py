"""`MergeEnv`: B lane-merging episodes stepped together (gym-style reset/step), the batched stand-in for
the reference's SUMO-backed `ContinuousJerkEnv` (reference merge_gym.py:15-225).

Observation, action handling, reward and termination follow the reference (merge_gym.py:83-162,
dqn.py:389-446, 557-563).  SUMO is not available, so the *world* is the reference's own traffic
predictor applied as dynamics: cars follow prediction.py:46-105 (K4 kernel), new cars enter every
BASE_TRAFFIC_INTERVAL + U[0,1) s at OTHER_CAR_SPEED (control.py:215-226), the ego moves along the ramp
geometry of control.py:366-380.  Parity of this world against SUMO's Krauss model is UNPINNED
(a "next" row); the per-tick kernels it calls are parity-tested on their own.
"""
from __future__ import annotations

import torch

from . import dqn, st, synthetic
from .config import Settings
from .prediction import BatchedState, tdiv

SPAWN_X = -250.0         # highway entry: lane highwayrear_0 starts at x = -250 (merge.net.xml:49)
ARRIVAL_X = 51.5         # the ego's route ends at arrivalPos = 50 on highwayahead (control.py:42), whose lane starts at x = 1.5 (merge.net.xml:46)
EGO_START_X = -211.40638598379087    # departPos = 40 (control.py:42) along the polyline of lane ramp_0 (merge.net.xml:52) ...
EGO_START_Y = 19.98667760899962      # ... measured from its first point (-250.47, 28.47)


KRAUSS = dict(accel=4.5, decel=6.0, tau=0.5, minGap=1.0)      # vType "normal" of merge_impossible.rou.xml:3 (sigma = 0: no dawdling)


class MergeEnv:
    """Vectorised environment.  reset() -> obs [B,20] f32;  step(jerk [B]) -> (obs, reward, done, info).

    The world keeps, per episode, the nmax (32) cars nearest to the action: a car that runs more than SENSOR_RADIUS
    ahead of the ego is recycled, new ones enter at the highway start.  Everything stays on the device and is
    branch-free (masks instead of host-side `if any()`), so a tick never synchronises with the host."""

    def __init__(self, num_envs: int, seed: int = 0, auto_reset: bool = True):
        self.B = int(num_envs)
        self.eng = st.get_engine(self.B)
        self.device = self.eng.device
        self.N = self.eng.nmax
        self.seed, self.auto_reset = seed, auto_reset
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))
        f64 = dict(dtype=torch.float64, device=self.device)
        self.state = BatchedState(torch.zeros((self.B, 4), **f64), torch.zeros((self.B, self.N), **f64),
                                  torch.zeros((self.B, self.N), **f64), torch.zeros((self.B, self.N), **f64),
                                  torch.zeros(self.B, dtype=torch.int32, device=self.device))
        self.delay = torch.zeros(self.B, **f64)
        self.ticks = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        self.episodes_started = 0
        self.max_ticks = int(Settings.MAX_EPISODE_LENGTH / Settings.TICK_LENGTH)
        self.prev_acc = torch.zeros(self.B, **f64)
        self._col = torch.arange(self.N, device=self.device).unsqueeze(0)
        self._ego_y0 = EGO_START_Y

    def _rand(self, *shape):
        return torch.rand(shape, generator=self.gen, dtype=torch.float64, device=self.device)

    # ---- initial conditions: spawner-spaced traffic (control.py:215-226), ego at the ramp start (control.py:41-44, 198-204) ----
    def _fresh(self):
        """A complete set of initial conditions for all B slots (cheap; rows are selected with a mask afterwards)."""
        S = Settings
        interval, speed = float(S.BASE_TRAFFIC_INTERVAL), float(S.OTHER_CAR_SPEED)
        gaps = speed * (interval + self._rand(self.B, self.N))
        gaps[:, 0] = self._rand(self.B) * speed * (interval + 0.5)
        xs = (EGO_START_X + float(S.SENSOR_RADIUS)) - torch.cumsum(gaps, 1)
        keep = xs >= SPAWN_X
        n = keep.sum(1).to(torch.int32)
        xs = torch.where(keep, xs, torch.zeros_like(xs))
        vs = torch.where(keep, torch.full_like(xs, speed), torch.zeros_like(xs))
        if S.RANDOMIZE_START_SPEED:
            v0 = (S.START_SPEED + S.START_SPEED_VARIANCE * torch.randn(self.B, generator=self.gen, dtype=torch.float64, device=self.device)) \
                .clamp(S.MIN_START_SPEED, S.MAX_START_SPEED)
        else:
            v0 = torch.full((self.B,), float(S.START_SPEED), dtype=torch.float64, device=self.device)
        ego = torch.stack([torch.full_like(v0, EGO_START_X), torch.full_like(v0, self._ego_y0), v0, torch.zeros_like(v0)], 1)
        return ego, xs, vs, n, interval + self._rand(self.B)

    def _reset_where(self, mask: torch.Tensor):
        ego, xs, vs, n, delay = self._fresh()
        m1, S = mask.unsqueeze(1), self.state
        S.ego = torch.where(m1, ego, S.ego); S.cars_x = torch.where(m1, xs, S.cars_x); S.cars_v = torch.where(m1, vs, S.cars_v)
        S.cars_a = torch.where(m1, torch.zeros_like(S.cars_a), S.cars_a); S.n_cars = torch.where(mask, n, S.n_cars)
        self.delay = torch.where(mask, delay, self.delay)
        self.ticks = torch.where(mask, torch.zeros_like(self.ticks), self.ticks)
        self.prev_acc = torch.where(mask, torch.zeros_like(self.prev_acc), self.prev_acc)

    def reset(self):
        self._reset_where(torch.ones(self.B, dtype=torch.bool, device=self.device))
        self.episodes_started = self.B
        return self._obs()

    def _obs(self):
        return dqn.get_state_vector_from_base_state(self.state)[:, :20]

    # ---- one tick ------------------------------------------------------------------------------------
    def _step_fused(self, action: torch.Tensor):
        """step() as one kernel (mpc_env_step).  The random numbers are drawn here, in the order step() draws them, so both
        paths walk through the same trajectories."""
        from ._lib import MpcEnvParams
        S, B, dev = Settings, self.B, self.device
        jerk = action.to(dev, torch.float64).reshape(B).contiguous()
        ep = MpcEnvParams()
        ep.tick, ep.a_min, ep.a_max, ep.max_speed = float(S.TICK_LENGTH), float(S.MAX_NEGATIVE_ACCELERATION), float(S.MAX_POSITIVE_ACCELERATION), float(S.MAX_SPEED)
        ep.min_crash_distance, ep.sensor_radius, ep.spawn_x = float(S.CAR_LENGTH), float(S.SENSOR_RADIUS), SPAWN_X
        ep.other_speed, ep.interval, ep.arrival_x = float(S.OTHER_CAR_SPEED), float(S.BASE_TRAFFIC_INTERVAL), ARRIVAL_X
        ep.ego_start_x, ep.ego_start_y = EGO_START_X, float(self._ego_y0)
        ep.start_speed, ep.start_speed_var = float(S.START_SPEED), float(S.START_SPEED_VARIANCE)
        ep.min_start_speed, ep.max_start_speed = float(S.MIN_START_SPEED), float(S.MAX_START_SPEED)
        ep.time_reward_step = S.TIME_REWARD * S.TICK_LENGTH
        ep.jerk_weight, ep.crash_reward, ep.success_reward = float(S.ALT_J_WEIGHT), float(S.CRASH_REWARD), float(S.SUCCESS_REWARD)
        ep.invalid_action_step = float(S.INVALID_ACTION_PENALTY) * float(S.TICK_LENGTH)
        ep.world = 1 if getattr(S, "WORLD_MODEL", "krauss") == "krauss" else 0
        ep.krauss_accel, ep.krauss_decel, ep.krauss_tau, ep.krauss_min_gap = KRAUSS["accel"], KRAUSS["decel"], KRAUSS["tau"], KRAUSS["minGap"]
        ep.max_ticks, ep.auto_reset = int(self.max_ticks), int(bool(self.auto_reset))
        u = self._rand(B) if S.VARY_TRAFFIC_START_TIMES else None
        fresh = None
        if self.auto_reset:                                   # the draws of _fresh(), in its order
            gap_u, first_u = self._rand(B, self.N), self._rand(B)
            z = torch.randn(B, generator=self.gen, dtype=torch.float64, device=dev) if S.RANDOMIZE_START_SPEED else None
            fresh = (gap_u, first_u, z, self._rand(B))
        if getattr(self, "_fused_out", None) is None:
            self._fused_out = (torch.empty(B, dtype=torch.float64, device=dev), torch.empty((4, B), dtype=torch.uint8, device=dev),
                               torch.empty(B, dtype=torch.float64, device=dev))
        reward, flags, pj = self._fused_out
        self.eng.env_step(ep, self.state.args(), self.prev_acc, self.delay, self.ticks, jerk, u, fresh, reward, flags, pj)
        fb = flags.view(torch.bool)
        done, crashed, arrived, timeout = fb[0], fb[1], fb[2], fb[3]
        info = {"crashed": crashed.clone(), "merged": arrived.clone(), "timeout": timeout.clone(), "projected_jerk": pj.clone()}
        obs = self._obs()
        if not self.auto_reset:
            obs = torch.where(done.unsqueeze(1), torch.zeros_like(obs), obs)
        return obs, reward.clone(), done.clone(), info

    def step(self, action: torch.Tensor):
        """action: jerk [B] (continuous, reference ContinuousJerkEnv).  Finished episodes restart when auto_reset."""
        if getattr(Settings, "FUSED_ENV_STEP", False):
            return self._step_fused(action)
        S, tick, st8 = Settings, float(Settings.TICK_LENGTH), self.state
        jerk = action.to(self.device, torch.float64).reshape(self.B)
        # _handle_jerk (merge_gym.py:83-96): the projected acceleration is clipped, ELSE the projected speed (computed with the
        # unclipped acceleration) is clipped and the acceleration follows from it; either way the action counts as invalid
        v_now, prev_acc = st8.ego[:, 2].clone(), self.prev_acc
        acc_p = prev_acc + jerk * tick
        spd_p = v_now + acc_p * tick
        acc_c = acc_p.clamp(S.MAX_NEGATIVE_ACCELERATION, S.MAX_POSITIVE_ACCELERATION)
        acc_bad = (acc_p > S.MAX_POSITIVE_ACCELERATION) | (acc_p < S.MAX_NEGATIVE_ACCELERATION)
        spd_bad = ~acc_bad & ((spd_p > S.MAX_SPEED) | (spd_p < 0))
        acc_proj = torch.where(acc_bad, acc_c, torch.where(spd_bad, tdiv(spd_p.clamp(0, S.MAX_SPEED) - v_now, tick), acc_p))
        projected_jerk = tdiv(acc_proj - prev_acc, tick)
        invalid = (acc_bad | spd_bad).to(torch.float64) * (float(S.INVALID_ACTION_PENALTY) * tick)
        # the command itself: control.set_ego_jerk -> get_ego_speed_from_jerk (control.py:160-176), both clamps in sequence
        spd = (v_now + acc_c * tick).clamp(0, S.MAX_SPEED)
        # world step: the reference predictor as dynamics (K4 kernel, in place)
        for t in st8.args():
            assert t.is_contiguous()
        if getattr(S, "WORLD_MODEL", "krauss") == "krauss":
            # the traffic follows SUMO's Krauss model with the vType of merge_impossible.rou.xml:3 (mpc_krauss_step)
            crashed = self.eng.krauss_step(*st8.args(), spd.contiguous(), tick, S.CAR_LENGTH, KRAUSS["accel"], KRAUSS["decel"],
                                           KRAUSS["tau"], KRAUSS["minGap"], float(S.OTHER_CAR_SPEED)).bool()
        else:
            _, _, _, _, crashed = self.eng.predict_step_with_ego(*st8.args(), spd.contiguous(), tick, S.CAR_LENGTH, inplace=True)
            crashed = crashed.bool()
        self.prev_acc = st8.ego[:, 3].clone()
        # recycle the front car once it is out of sensor range ahead; enter a new car at the back (control.py:215-226)
        n = st8.n_cars
        gone = (n > 0) & (st8.cars_x[:, 0] - st8.ego[:, 0] > float(S.SENSOR_RADIUS))
        g1 = gone.unsqueeze(1)
        shift = lambda a: torch.where(g1, torch.cat([a[:, 1:], torch.zeros_like(a[:, :1])], 1), a)   # noqa: E731
        st8.cars_x, st8.cars_v, st8.cars_a = shift(st8.cars_x), shift(st8.cars_v), shift(st8.cars_a)
        n = n - gone.to(torch.int32)
        self.delay = self.delay - tick
        spawn = (self.delay <= 0) & (n < self.N)
        slot = spawn.unsqueeze(1) & (self._col == n.unsqueeze(1))
        st8.cars_x = torch.where(slot, torch.full_like(st8.cars_x, SPAWN_X), st8.cars_x)
        st8.cars_v = torch.where(slot, torch.full_like(st8.cars_v, float(S.OTHER_CAR_SPEED)), st8.cars_v)
        st8.cars_a = torch.where(slot, torch.zeros_like(st8.cars_a), st8.cars_a)
        st8.n_cars = n + spawn.to(torch.int32)
        u = self._rand(self.B) if S.VARY_TRAFFIC_START_TIMES else torch.zeros(self.B, dtype=torch.float64, device=self.device)
        self.delay = torch.where(spawn, u + float(S.BASE_TRAFFIC_INTERVAL), self.delay)
        self.ticks = self.ticks + 1
        arrived = (st8.ego[:, 0] > ARRIVAL_X) & ~crashed
        timeout = (self.ticks >= self.max_ticks) & ~crashed & ~arrived
        done = crashed | arrived | timeout
        # merge_gym.py:102-140: terminal reward on a crash / an arrival, else the time + jerk reward of the MEASURED jerk (new
        # acceleration against the previous one); a clipped action adds its penalty (INVALID_ACTION_PENALTY, 0 in every published config)
        reward = dqn.slotted_reward_with_jerk(None, tdiv(st8.ego[:, 3] - prev_acc, tick), crashed, arrived) + invalid
        info = {"crashed": crashed, "merged": arrived, "timeout": timeout, "projected_jerk": projected_jerk}
        if self.auto_reset:
            self._reset_where(done)
        obs = self._obs()
        if not self.auto_reset:
            obs = torch.where(done.unsqueeze(1), torch.zeros_like(obs), obs)
        return obs, reward, done, info


class JerkEnv(MergeEnv):
    """Alias with the reference's class name (merge_gym.py:15)."""


ContinuousJerkEnv = MergeEnv
