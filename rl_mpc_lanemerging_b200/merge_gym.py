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

from typing import Optional

import numpy as np
import torch

from . import dqn, st, synthetic
from .config import Settings
from .prediction import BatchedState

SPAWN_X = -250.0         # highway entry
ARRIVAL_X = 75.0         # ego route end on the highway edge (approximate; unpinned)
EGO_START_X = -215.0     # rampRoute departPos 40 (control.py:41-44)


class MergeEnv:
    """Vectorised environment.  reset() -> obs [B,20] f32;  step(jerk [B]) -> (obs, reward, done, info).

    The world keeps, per episode, the nmax (32) cars nearest to the action: cars that fall more than
    SENSOR_RADIUS behind the ego or run far ahead are recycled, new ones enter at the highway start."""

    def __init__(self, num_envs: int, seed: int = 0, auto_reset: bool = True):
        self.B = int(num_envs)
        self.eng = st.get_engine(self.B)
        self.device = self.eng.device
        self.N = self.eng.nmax
        self.seed, self.auto_reset = seed, auto_reset
        self.episode_id = np.arange(self.B, dtype=np.int64)
        self.next_episode = self.B
        f64 = dict(dtype=torch.float64, device=self.device)
        self.state = BatchedState(torch.zeros((self.B, 4), **f64), torch.zeros((self.B, self.N), **f64),
                                  torch.zeros((self.B, self.N), **f64), torch.zeros((self.B, self.N), **f64),
                                  torch.zeros(self.B, dtype=torch.int32, device=self.device))
        self.delay = torch.zeros(self.B, **f64)
        self.ticks = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        self.spawn_draws = np.zeros(self.B, dtype=np.int64)
        self.max_ticks = int(Settings.MAX_EPISODE_LENGTH / Settings.TICK_LENGTH)
        self.prev_acc = torch.zeros(self.B, **f64)

    # ---- initial conditions: spawner-spaced traffic on the whole road, ego at the ramp start ------------
    def _reset_rows(self, rows: np.ndarray):
        ids = self.episode_id[rows]
        n = len(ids)
        interval, speed = float(Settings.BASE_TRAFFIC_INTERVAL), float(Settings.OTHER_CAR_SPEED)
        u = lambda d: synthetic.uniform(self.seed, ids, d)   # noqa: E731
        d = np.arange(self.N)[None, :]
        gaps = speed * (interval + synthetic.uniform(self.seed, ids[:, None], 16 + d))
        gaps[:, 0] = u(16) * speed * (interval + 0.5)
        # nearest nmax cars around the ramp end: start the platoon a little ahead of the merge area
        xs = (EGO_START_X + float(Settings.SENSOR_RADIUS)) - np.cumsum(gaps, axis=1)
        keep = xs >= SPAWN_X
        cnt = keep.sum(1).astype(np.int32)
        xs = np.where(keep, xs, 0.0)
        z = np.sqrt(-2.0 * np.log(1.0 - u(1))) * np.cos(2.0 * np.pi * u(2))
        if Settings.RANDOMIZE_START_SPEED:
            v0 = np.clip(Settings.START_SPEED + Settings.START_SPEED_VARIANCE * z, Settings.MIN_START_SPEED, Settings.MAX_START_SPEED)
        else:
            v0 = np.full(n, float(Settings.START_SPEED))
        frac = (EGO_START_X - synthetic.RAMP_A[0]) / (synthetic.RAMP_B[0] - synthetic.RAMP_A[0])
        ey = synthetic.RAMP_A[1] + frac * (synthetic.RAMP_B[1] - synthetic.RAMP_A[1])
        ego = np.stack([np.full(n, EGO_START_X), np.full(n, ey), v0, np.zeros(n)], 1)
        r = torch.from_numpy(rows).to(self.device)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)   # noqa: E731
        S = self.state
        S.ego[r] = t(ego); S.cars_x[r] = t(xs); S.cars_v[r] = t(np.where(keep, speed, 0.0)); S.cars_a[r] = 0.0
        S.n_cars[r] = t(cnt)
        self.delay[r] = t(interval + u(3))
        self.ticks[r] = 0; self.prev_acc[r] = 0.0
        self.spawn_draws[rows] = 0

    def reset(self):
        self._reset_rows(np.arange(self.B))
        return self._obs()

    def _obs(self):
        return dqn.get_state_vector_from_base_state(self.state)[:, :20]

    # ---- one tick ------------------------------------------------------------------------------------
    def step(self, action: torch.Tensor):
        """action: jerk [B] (continuous, reference ContinuousJerkEnv).  Finished episodes restart when auto_reset."""
        S, tick, st8 = Settings, float(Settings.TICK_LENGTH), self.state
        jerk = action.to(self.device, torch.float64).reshape(self.B)
        # merge_gym.py:83-96: clip the projected acceleration / speed, remember the realised jerk
        acc = (self.prev_acc + jerk * tick).clamp(S.MAX_NEGATIVE_ACCELERATION, S.MAX_POSITIVE_ACCELERATION)
        spd = st8.ego[:, 2] + acc * tick
        clipped = (spd > S.MAX_SPEED) | (spd < 0)
        spd = spd.clamp(0, S.MAX_SPEED)
        acc = torch.where(clipped, (spd - st8.ego[:, 2]) / tick, acc)
        projected_jerk = (acc - self.prev_acc) / tick
        # world step: the reference predictor as dynamics (K4 kernel, in place)
        _, _, _, _, crashed = self.eng.predict_step_with_ego(*st8.args(), spd.contiguous(), tick, S.CAR_LENGTH, inplace=True)
        crashed = crashed.bool()
        self.prev_acc = st8.ego[:, 3].clone()
        # recycle the front car once it is out of sensor range ahead; enter a new car at the back (control.py:215-226)
        n = st8.n_cars.long()
        gone = (n > 0) & (st8.cars_x[:, 0] - st8.ego[:, 0] > float(S.SENSOR_RADIUS))
        if bool(gone.any()):
            r = gone.nonzero().squeeze(1)
            for arr in (st8.cars_x, st8.cars_v, st8.cars_a):
                arr[r] = torch.roll(arr[r], -1, dims=1)
                arr[r, -1] = 0.0
            st8.n_cars[r] -= 1
        self.delay -= tick
        spawn = (self.delay <= 0) & (st8.n_cars < self.N)
        if bool(spawn.any()):
            r = spawn.nonzero().squeeze(1)
            slot = st8.n_cars[r].long()
            st8.cars_x[r, slot] = SPAWN_X; st8.cars_v[r, slot] = float(S.OTHER_CAR_SPEED); st8.cars_a[r, slot] = 0.0
            st8.n_cars[r] += 1
            rows = r.cpu().numpy()
            u = synthetic.uniform(self.seed + 7919, self.episode_id[rows], self.spawn_draws[rows]) if S.VARY_TRAFFIC_START_TIMES \
                else np.zeros(len(rows))
            self.delay[r] = torch.from_numpy(u + float(S.BASE_TRAFFIC_INTERVAL)).to(self.device)
            self.spawn_draws[rows] += 1
        self.ticks += 1
        arrived = (st8.ego[:, 0] > ARRIVAL_X) & ~crashed
        timeout = (self.ticks >= self.max_ticks) & ~crashed & ~arrived
        done = crashed | arrived | timeout
        reward = dqn.slotted_reward_with_jerk(None, projected_jerk, crashed, arrived)
        info = {"crashed": crashed, "merged": arrived, "timeout": timeout, "projected_jerk": projected_jerk}
        if self.auto_reset and bool(done.any()):
            rows = done.nonzero().squeeze(1).cpu().numpy()
            self.episode_id[rows] = np.arange(self.next_episode, self.next_episode + len(rows))
            self.next_episode += len(rows)
            self._reset_rows(rows)
        obs = self._obs()
        if not self.auto_reset:
            obs = torch.where(done.unsqueeze(1), torch.zeros_like(obs), obs)
        return obs, reward, done, info


class JerkEnv(MergeEnv):
    """Alias with the reference's class name (merge_gym.py:15)."""


ContinuousJerkEnv = MergeEnv
