"""CPU restatement of the reference's environment tick and combined-controller decision -- TEST INFRASTRUCTURE ONLY.

Plain Python / numpy, one episode at a time, every function citing the reference lines it follows.  Only tests/ may import
this module (same rule as oracle/cpu_oracle.py).  The SUMO side of the tick (where the cars are after `control.step()`,
whether the ego collided or arrived) is an INPUT here: these functions restate what the reference computes around it --
action clipping, the speed command, reward, termination -- which is what `merge_gym.MergeEnv.step` and `mpc_env_step` must
reproduce for whatever world model sits underneath.
"""
from __future__ import annotations

import numpy as np


def speed_from_jerk(S, v, a, jerk):
    """control.get_ego_speed_from_jerk (reference control.py:160-171)."""
    na = a + jerk * S.TICK_LENGTH
    if na > S.MAX_POSITIVE_ACCELERATION:
        na = S.MAX_POSITIVE_ACCELERATION
    if na < S.MAX_NEGATIVE_ACCELERATION:
        na = S.MAX_NEGATIVE_ACCELERATION
    ns = v + na * S.TICK_LENGTH
    if ns > S.MAX_SPEED:
        ns = S.MAX_SPEED
    if ns < 0:
        ns = 0
    return ns


def handle_jerk(S, ego_speed, previous_acceleration, selected_jerk, penalty_for_invalid_action=None):
    """JerkEnv._handle_jerk (reference merge_gym.py:83-96).  Returns (projected_jerk, invalid_action_reward, commanded speed);
    the command is control.set_ego_jerk(selected_jerk) (control.py:173-178) with the simulator's current speed / acceleration,
    which are the previous state's."""
    pen = S.INVALID_ACTION_PENALTY if penalty_for_invalid_action is None else penalty_for_invalid_action
    projected_acceleration = previous_acceleration + selected_jerk * S.TICK_LENGTH
    projected_speed = ego_speed + projected_acceleration * S.TICK_LENGTH
    if projected_acceleration > S.MAX_POSITIVE_ACCELERATION or projected_acceleration < S.MAX_NEGATIVE_ACCELERATION:
        invalid = pen * S.TICK_LENGTH
        projected_acceleration = float(np.clip(projected_acceleration, S.MAX_NEGATIVE_ACCELERATION, S.MAX_POSITIVE_ACCELERATION))
    elif projected_speed > S.MAX_SPEED or projected_speed < 0:
        invalid = pen * S.TICK_LENGTH
        projected_speed = float(np.clip(projected_speed, 0, S.MAX_SPEED))
        projected_acceleration = (projected_speed - ego_speed) / S.TICK_LENGTH
    else:
        invalid = 0
    projected_jerk = (projected_acceleration - previous_acceleration) / S.TICK_LENGTH
    return projected_jerk, invalid, speed_from_jerk(S, ego_speed, previous_acceleration, selected_jerk)


def slotted_reward_with_jerk(S, jerk, crashed, arrived):
    """dqn.slotted_reward_with_jerk (reference dqn.py:557-563)."""
    if crashed:
        return S.CRASH_REWARD
    elif arrived:
        return S.SUCCESS_REWARD
    return S.TIME_REWARD * S.TICK_LENGTH - S.ALT_J_WEIGHT * jerk ** 2 * S.TICK_LENGTH


def env_step(S, ego_speed, previous_acceleration, selected_jerk, ticks_before, collided, arrived, new_acceleration):
    """JerkEnv.step (reference merge_gym.py:102-140) around the simulator: `collided` / `arrived` / `new_acceleration` are what
    control.just_had_collision(), control.ego_just_arrived() and the new HighwayState report after control.step().
    Returns dict(projected_jerk, speed_command, reward, done, crashed, merged, timeout, zero_observation, next_previous_acceleration)."""
    ticks = ticks_before + 1                                                       # :104
    projected_jerk, invalid, command = handle_jerk(S, ego_speed, previous_acceleration, selected_jerk)      # :105
    max_ticks = S.MAX_EPISODE_LENGTH / S.TICK_LENGTH                                 # :26
    out = dict(projected_jerk=projected_jerk, speed_command=command, crashed=False, merged=False, timeout=False,
               zero_observation=False, next_previous_acceleration=previous_acceleration)
    if collided:                                                                     # :108-112
        out.update(reward=slotted_reward_with_jerk(S, projected_jerk, True, False) + invalid, done=True, crashed=True, zero_observation=True)
    elif arrived:                                                                    # :113-117
        out.update(reward=slotted_reward_with_jerk(S, projected_jerk, False, True) + invalid, done=True, merged=True, zero_observation=True)
    elif ticks >= max_ticks:                                                         # :118-126
        jerk = (new_acceleration - previous_acceleration) / S.TICK_LENGTH
        out.update(reward=slotted_reward_with_jerk(S, jerk, False, False) + invalid, done=True, timeout=True)
    else:                                                                            # :127-140
        jerk = (new_acceleration - previous_acceleration) / S.TICK_LENGTH
        out.update(reward=slotted_reward_with_jerk(S, jerk, False, False) + invalid, done=False,
                   next_previous_acceleration=new_acceleration)
    return out
