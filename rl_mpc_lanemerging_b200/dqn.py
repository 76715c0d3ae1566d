"""Agent-side pieces of the hot path with the reference's names (reference dqn.py:70-200, 389-446, 557-563):
observation vector, reward, and the RL-proposes / MPC-vetoes decision loop -- all batched.

The reference evaluates one episode at a time with Python branches; here the same decision chain runs
for B episodes with masks, the predictor / observation / clamp steps in K4 kernels and the veto in the
fused K3 planner.  The policy forward is plain PyTorch on device tensors (north_star: "PyTorch only for
the existing Q-network forward").
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import control, st
from .config import Settings
from .prediction import BatchedState, HighwayState, tdiv


def get_state_vector_from_base_state(state, out: Optional[torch.Tensor] = None):
    """20-dim observation (reference dqn.py:389-446 with CARS_AHEAD = CARS_BEHIND = 2, accelerations, speed
    differences, normalisation -- the only variant the published configs use).
    HighwayState -> float32 numpy[20];  BatchedState -> float32 tensor [B,21] whose last column is left for the
    time feature the DDPG policy appends."""
    if isinstance(state, BatchedState):
        return st.get_engine(state.batch).state_vector(*state.args(), out=out)
    eng = st.get_engine()
    bs = BatchedState.from_states([state], eng.device, eng.nmax)
    return eng.state_vector(*bs.args())[0, :20].cpu().numpy()


def slotted_reward_with_jerk(state, jerk, crashed, arrived):
    """Reference dqn.py:557-563.  Scalars or tensors."""
    step = Settings.TIME_REWARD * Settings.TICK_LENGTH - Settings.ALT_J_WEIGHT * jerk ** 2 * Settings.TICK_LENGTH
    if torch.is_tensor(jerk):
        r = torch.where(arrived, torch.full_like(step, float(Settings.SUCCESS_REWARD)), step)
        return torch.where(crashed, torch.full_like(step, float(Settings.CRASH_REWARD)), r)
    if crashed:
        return Settings.CRASH_REWARD
    if arrived:
        return Settings.SUCCESS_REWARD
    return step


def get_reward_function():
    if Settings.REWARD_FUNCTION != "Slotted Jerk":
        raise NotImplementedError("only the 'Slotted Jerk' reward of the published configs is provided")
    return slotted_reward_with_jerk


class RLAgent:
    """Base class: subclasses provide get_control(state) -> jerk (reference dqn.py:70-77)."""

    def __init__(self):
        self.takeover_history: List = []

    def get_control(self, state):
        raise NotImplementedError

    def select_action(self, state):                 # north_star alias
        return self.get_control(state)

    def _setup(self):
        self.takeover_history = []

    def _cleanup(self):
        pass

    def do_control(self, state):
        """Apply the agent's jerk: the commanded speed after one tick (reference dqn.py:99-100 -> control.set_ego_jerk)."""
        jerk = self.get_control(state)
        if isinstance(state, HighwayState):
            return control.get_ego_speed_from_jerk(state.ego_speed, state.ego_acceleration, jerk)
        return control.get_ego_speed_from_jerk(state.ego[:, 2].contiguous(), state.ego[:, 3].contiguous(), jerk.double())

    def end_episode_callback(self, last_state=None):
        pass

    def combined_stats_callback(self, episode_stats):
        """Share of the ticks on which the planner took over (reference dqn.py:102-115)."""
        return {"percent st solver": episode_stats.get("takeovers", 0) / max(episode_stats.get("steps", 0), 1)}

    # ---- reference dqn.py:202-241 ----
    def evaluate(self, num_episodes, num_envs=None, csv_path="run_data.csv"):
        self._setup()
        output = control.evaluate_control(control_function=self.do_control, num_episodes=num_episodes,
                                          end_episode_callback=self.end_episode_callback, num_envs=num_envs or Settings.EVAL_NUM_ENVS)
        output.print_stats(csv_path)
        self._cleanup()
        return output

    def evaluate_combined(self, num_episodes, num_envs=None, csv_path="run_data.csv"):
        self._setup()
        output = control.evaluate_control(control_function=self.do_combined_control, num_episodes=num_episodes,
                                          end_episode_callback=self.end_episode_callback,
                                          custom_stats_function=self.combined_stats_callback, num_envs=num_envs or Settings.EVAL_NUM_ENVS)
        self._cleanup()
        output.print_stats(csv_path)
        return output

    # ---- reference dqn.py:117-200 ----------------------------------------------------------------
    def do_combined_control(self, state):
        """RL proposes a jerk, a short policy rollout and the MPC planner decide whether the planner takes over.
        HighwayState -> commanded speed (float); BatchedState -> (speed tensor [B], takeover mask [B])."""
        single = isinstance(state, HighwayState)
        eng = st.get_engine(1 if single else state.batch)
        start = BatchedState.from_states([state], eng.device, eng.nmax) if single else state
        speed, takeover = self._combined_batched(start, eng)
        self.takeover_history.append(bool(takeover.item()) if single else takeover)
        return float(speed.item()) if single else (speed, takeover)

    def _combined_batched(self, start: BatchedState, eng):
        B, dev = start.batch, eng.device
        first_action = self.get_control(start)
        cur = start.clone()
        steps = max(int(Settings.ROLLOUT_LENGTH), 1)
        alive = torch.ones(B, dtype=torch.uint8, device=dev)
        crash_u8 = torch.zeros(B, dtype=torch.uint8, device=dev)
        selected_speed = torch.zeros(B, dtype=torch.float64, device=dev)
        # rollout_s_history (dqn.py:121,139): ego arclength after every rollout step the episode actually took
        roll_s = torch.zeros((B, steps + 1), dtype=torch.float64, device=dev)
        roll_s[:, 0] = _ego_s(start.ego)
        roll_len32 = torch.ones(B, dtype=torch.int32, device=dev)
        test = cur
        for i in range(1, steps + 1):                                               # dqn.py:129-141, one fused kernel per step
            action = first_action if i == 1 else self.get_control(cur)
            eng.rollout_step(cur.args(), action.double().reshape(B).contiguous(), Settings.TICK_LENGTH,
                             Settings.COMBINATION_MIN_DISTANCE, Settings.STOP_X, i, alive, selected_speed, roll_s, roll_len32, crash_u8)
            if i == 1:
                rl_speed = selected_speed.clone()                                    # every episode takes step 1: the RL command
            if i == int(Settings.ST_TEST_ROLLOUTS) and i < steps:                   # dqn.py:142-143 (finished episodes are frozen in
                test = cur.clone()                                                   # place, so the snapshot is the test state for all)
        crash_predicted, roll_len = crash_u8.bool(), roll_len32.long()
        takeover = torch.zeros(B, dtype=torch.bool, device=dev)
        if Settings.CHECK_ROLLOUT_CRASH:                                            # 144-147
            takeover |= crash_predicted
        if Settings.LIMIT_DQN_SPEED:                                                # 148-151
            takeover |= selected_speed > Settings.DESIRED_SPEED
        if Settings.TEST_ROLLOUT_STATE:                                             # 152-155: full gap-evaluation from the rollout state
            if getattr(self, "_plan_hint", None) is None:
                self._plan_hint = st.PlanHint()                                      # episode b's rollout state is re-planned every tick
            takeover |= st.test_guaranteed_crash_from_state(test, hint=self._plan_hint)
        speed = rl_speed
        if Settings.TEST_ST_STRICTLY_BETTER:                                        # 156-197 ("b" configs)
            if Settings.REMEMBER_LAST_CHOICE_FOR_SWITCHING_COMBINED:
                raise NotImplementedError("REMEMBER_LAST_CHOICE_FOR_SWITCHING_COMBINED is False in every published config")
            plan = st.plan_batch(start)
            st_speed, fine, n_fine = st.smoothed_speed(start, plan)
            tick = float(Settings.TICK_LENGTH)
            mlen = torch.minimum(n_fine.long(), roll_len)                           # min_planning_length (171)
            st_jerk = _mean_abs_jerk(fine, mlen, start.ego[:, 2], start.ego[:, 3], tick)
            rl_jerk = _mean_abs_jerk(roll_s, mlen, start.ego[:, 2], start.ego[:, 3], tick)
            last = (mlen - 1).clamp(min=0).unsqueeze(1)
            st_dist = fine.gather(1, last).squeeze(1) - fine[:, 0]
            rl_dist = roll_s.gather(1, last).squeeze(1) - roll_s[:, 0]
            better = ((st_jerk < rl_jerk) & (st_dist > rl_dist)) | (rl_dist == 0)   # 177
            better &= (n_fine > 1) & ~takeover                                       # 166-169: single-point plan -> keep the RL action
            speed = torch.where(better, st_speed, rl_speed)
            takeover_b = better
        else:
            takeover_b = None
        if getattr(Settings, "SYNC_FREE_TAKEOVER", False):                         # planner takes over, episode list kept on the device
            speed = speed.clone()
            self._takeover_scratch = st.do_st_control_masked(start, takeover, speed, getattr(self, "_takeover_scratch", None))
            idx = takeover[:0]
        else:
            idx = takeover.nonzero().squeeze(1)                                    # (the tick's one host sync: the list's length)
        if idx.numel():                                                             # planner takes over: st.do_st_control(start_state)
            sub = BatchedState(*(t[idx].contiguous() for t in start.args()))
            speed = speed.clone()
            speed[idx] = st.do_st_control(sub)
        if takeover_b is not None:
            takeover = takeover | takeover_b
        return speed, takeover


def _ego_s(ego: torch.Tensor) -> torch.Tensor:
    """control.get_ego_s (control.py:373-380) for a batch (device tensor ops on 3 columns; rollout bookkeeping only)."""
    x, y = ego[:, 0], ego[:, 1]
    d = torch.sqrt((x - control.merge_point[0]) ** 2 + (y - control.merge_point[1]) ** 2)
    return torch.where(x < control.merge_point[0], -d, torch.where(x < control.merge_point2[0], d, x - control.merge_point2[0] + control.common_s))


def _mean_abs_jerk(seq: torch.Tensor, length: torch.Tensor, v0: torch.Tensor, a0: torch.Tensor, dt: float) -> torch.Tensor:
    """st.get_path_mean_abs_jerk (st.py:274-288) over the first `length[b]` points of every row."""
    B, N = seq.shape
    v = tdiv(seq[:, 1:] - seq[:, :-1], dt)
    a = tdiv(v - torch.cat([v0.unsqueeze(1), v[:, :-1]], 1), dt)
    j = tdiv(a - torch.cat([a0.unsqueeze(1), a[:, :-1]], 1), dt)
    mask = torch.arange(N - 1, device=seq.device).unsqueeze(0) < (length - 1).unsqueeze(1)
    return (j.abs() * mask).sum(1) / (length - 1).clamp(min=1)
