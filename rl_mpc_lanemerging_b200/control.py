"""Geometry and actuation helpers with the reference's names (reference control.py:160-171, 366-389).

The SUMO/traci episode loop of the reference (control.run_episode) has no counterpart here: episodes are
stepped in batches by merge_gym.MergeEnv.  Scalars and tensors are both accepted; tensor inputs stay on the
device (K4 kernels), scalars are evaluated through the same kernels as a batch of one.
"""
from __future__ import annotations

import math

import torch

from .config import Settings

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
