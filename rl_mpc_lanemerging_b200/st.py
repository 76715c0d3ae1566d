"""Spatio-temporal planner front end with the reference's function names (reference st.py:726-802).

Every function accepts either one `HighwayState` (drop-in: numpy results shaped like the reference's)
or a `BatchedState` (device tensors in, device tensors out).  All computation happens in libmpcb200
(K1 grid build, K2 dense solve, K3 fused plan); there is no CPU implementation here.

Arithmetic: single-state calls default to Settings.ST_MODE ("exact": fp64, index-identical to the
reference's st_cy); batched calls default to "fast" (integer kinematics, fp64 labels; sequences match
the reference on >99% of states, cost within 1e-6 rel -- DESIGN.md §Parity).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .config import Settings
from .engine import MpcEngine, params_from_settings, params_key
from .prediction import BatchedState, HighwayState, tdiv

_engine: Optional[MpcEngine] = None
_engine_key = None


def get_engine(min_batch: int = 1) -> MpcEngine:
    """The process-wide engine for the current Settings snapshot (rebuilt when Settings or capacity change)."""
    global _engine, _engine_key
    p = params_from_settings(Settings)
    dev = int(getattr(Settings, "CUDA_DEVICE", 0))
    key = (params_key(p), dev)
    if _engine is None or key != _engine_key or _engine.max_batch < min_batch:
        # (the previous engine is NOT closed here: environments / agents created earlier may still hold it; it is destroyed with
        # its last reference.  refresh_engine() closes explicitly.)
        cap = max(min_batch, 64)
        _engine = MpcEngine(p, device=dev, max_batch=cap)
        _engine_key = key
    return _engine


def refresh_engine():
    """Drop the cached engine (call after mutating Settings in place with unusual values)."""
    global _engine, _engine_key
    if _engine is not None:
        _engine.close()
    _engine, _engine_key = None, None


def _t_values(eng: MpcEngine):
    return np.arange(eng.num_t) * float(Settings.T_DISCRETIZATION)


def _s_values(start_s: float, delta_s: float, num_s: int):
    s = start_s + np.arange(num_s) * delta_s
    if num_s > 1:
        s[1] = start_s + float(Settings.S_DISCRETIZATION)
    return s


def find_s_t_obstacles_from_state(current_state: HighwayState, *_ignored, **_kw):
    """(obstacles bool[num_t,num_s], s_values, t_values, ego_speed, distances f64[num_t,num_s]) -- reference st.py:25-70.
    Discretisation / horizon come from Settings (the reference passes the same Settings values positionally)."""
    eng = get_engine()
    bs = BatchedState.from_states([current_state], eng.device, eng.nmax)
    g = eng.build_grid(*bs.args())
    ns = int(g["num_s"].item())
    obstacles = g["obstacles"][0, :, :ns].cpu().numpy().astype(bool)
    distances = g["distances"][0, :, :ns].cpu().numpy()
    s_values = _s_values(g["start_s"].item(), g["delta_s"].item(), ns)
    return obstacles, s_values, _t_values(eng), current_state.ego_speed, distances


class PlanHint:
    """Cost and reached_t of the previous plan of the SAME episodes (row b = episode b), kept by a closed-loop caller that
    re-plans every tick like the reference's control.run_episode (control.py:229-340).  With Settings.PLAN_COST_HINTS the next
    plan_batch(..., hint=h) bounds each episode's first attempt by PLAN_HINT_SCALE x its previous cost (mpc_plan_hinted).
    Plans do not depend on the hint -- a stale row (episode reset, other batch) only costs that episode a retry."""

    def __init__(self):
        self.cost: Optional[torch.Tensor] = None
        self.reached: Optional[torch.Tensor] = None


def plan_batch(batch: BatchedState, mode: Optional[str] = None, out: Optional[dict] = None, hint: Optional[PlanHint] = None) -> dict:
    """Fused gap-evaluation for a batch: dict(idx, s_seq, cost, reached_t, crash, min_dist, start_s) of device tensors."""
    eng = get_engine(batch.batch)
    mode = mode or "fast"
    hinted = hint is not None and mode == "fast" and bool(getattr(Settings, "PLAN_COST_HINTS", False))
    if hinted and hint.cost is not None and hint.cost.shape[0] == batch.batch:
        r = eng.plan_hinted(*batch.args(), hint_cost=hint.cost, hint_reached=hint.reached, hint_full_t=eng.num_t - 1,
                            hint_scale=float(getattr(Settings, "PLAN_HINT_SCALE", 1.15)), mode=mode, out=out)
    else:
        r = eng.plan(*batch.args(), mode=mode, out=out)
    if hinted:                                      # remember this plan for the next tick (own tensors: never alias the outputs)
        if hint.cost is None or hint.cost.shape != r["cost"].shape:
            hint.cost, hint.reached = r["cost"].clone(), r["reached_t"].clone()
        else:
            hint.cost.copy_(r["cost"]); hint.reached.copy_(r["reached_t"])
    return r


def get_appropriate_base_st_path_and_obstacles(state):
    """Reference st.py:726-754.  HighwayState -> (s_sequence, obstacles, s_values, t_values, distances) as numpy
    (unreached layers are 0.0 like the reference).  BatchedState -> plan_batch(state)."""
    if isinstance(state, BatchedState):
        return plan_batch(state)
    obstacles, s_values, t_values, _v, distances = find_s_t_obstacles_from_state(state)
    eng = get_engine()
    bs = BatchedState.from_states([state], eng.device, eng.nmax)
    r = eng.plan(*bs.args(), mode=getattr(Settings, "ST_MODE", "exact"))
    return r["s_seq"][0].cpu().numpy(), obstacles, s_values, t_values, distances


solve = get_appropriate_base_st_path_and_obstacles          # the name BASELINE.json's north_star uses


def test_guaranteed_crash_from_state(state, hint: Optional[PlanHint] = None):
    """Reference st.py:790-802: True when the plan is incomplete or touches a cell closer than
    COMBINATION_MIN_DISTANCE - CAR_LENGTH to traffic.  BatchedState -> bool tensor [B] (hint: see PlanHint)."""
    if isinstance(state, BatchedState):
        return plan_batch(state, hint=hint)["crash"].bool()
    eng = get_engine()
    bs = BatchedState.from_states([state], eng.device, eng.nmax)
    return bool(eng.plan(*bs.args(), mode=getattr(Settings, "ST_MODE", "exact"))["crash"].item())


test_guaranteed_crash_from_state.__test__ = False            # not a pytest test


def finer_fit(s_sequence, delta_t=None, coarse_delta_t=None, start_speed=0.0, start_acceleration=0.0, before_after_cars=None):
    """Reference st.py:584-723: re-sample a coarse plan to the control tick and project it onto the speed / acceleration /
    jerk limits.  Same QP as the reference's (cvxopt) one, solved on the GPU by an interior-point kernel (mpc_finer_fit).
    numpy in -> numpy out (one plan).  `before_after_cars` (unused by every caller in the reference) is not supported."""
    if before_after_cars is not None:
        raise NotImplementedError("before_after_cars is never passed by the reference's callers")
    s = np.asarray(s_sequence, dtype=np.float64)
    if len(s) == 1:
        return s
    if delta_t is not None and abs(delta_t - float(Settings.TICK_LENGTH)) > 1e-12 or \
            coarse_delta_t is not None and abs(coarse_delta_t - float(Settings.T_DISCRETIZATION)) > 1e-12:
        raise ValueError("finer_fit runs with Settings.TICK_LENGTH / Settings.T_DISCRETIZATION")
    eng = get_engine()
    if len(s) > eng.num_t:
        raise ValueError("plan longer than the planning horizon")
    seq = torch.zeros((1, eng.num_t), dtype=torch.float64, device=eng.device)
    seq[0, :len(s)] = torch.from_numpy(s).to(eng.device)
    reached = torch.tensor([len(s) - 1], dtype=torch.int32, device=eng.device)
    ego = torch.tensor([[0.0, 0.0, float(start_speed), float(start_acceleration)]], dtype=torch.float64, device=eng.device)
    fine, n_fine, _speed, _it = eng.finer_fit(seq, reached, ego)
    return fine[0, :int(n_fine.item())].cpu().numpy()


def smoothed_speed(batch: BatchedState, plan: dict):
    """Speed command of reference st.py:757-783 for a batch: trim the 0.0 tail, finer_fit (TICK_LENGTH < T_DISCRETIZATION in
    every published config), then (x[1]-x[0])/TICK_LENGTH; plans with a single point keep the current speed."""
    eng = get_engine(batch.batch)
    if float(Settings.TICK_LENGTH) < float(Settings.T_DISCRETIZATION):
        fine, n_fine, speed, _it = eng.finer_fit(plan["s_seq"], plan["reached_t"], batch.ego)
        return speed, fine, n_fine
    s = plan["s_seq"]
    v = tdiv(s[:, 1] - s[:, 0], Settings.TICK_LENGTH)
    return torch.where(plan["reached_t"] >= 1, v, batch.ego[:, 2]), s, plan["reached_t"] + 1


def do_st_control(state, hint: Optional[PlanHint] = None):
    """Reference st.py:757-783.  Returns the commanded ego speed (float, or tensor [B] for a BatchedState).
    hint: see PlanHint (batched closed-loop callers)."""
    if isinstance(state, BatchedState):
        return smoothed_speed(state, plan_batch(state, hint=hint))[0]
    eng = get_engine()
    bs = BatchedState.from_states([state], eng.device, eng.nmax)
    plan = eng.plan(*bs.args(), mode=getattr(Settings, "ST_MODE", "exact"))
    return float(smoothed_speed(bs, plan)[0].item())


def do_st_control_masked(batch: BatchedState, mask: torch.Tensor, speed: torch.Tensor, scratch: Optional[dict] = None) -> dict:
    """do_st_control for the episodes with mask[b] set, written into speed[b] in place; the other rows of `speed` keep their
    values.  Nothing is read back to the host (mpc_plan_masked + mpc_finer_fit_masked).  Returns the scratch tensors for re-use."""
    eng = get_engine(batch.batch)
    B = batch.batch
    if scratch is None or scratch["cost"].shape[0] != B:
        scratch = eng._plan_out(B)
        scratch["fine"] = torch.zeros((B, int(eng.lib.mpc_finer_fit_max_points())), dtype=torch.float64, device=eng.device)
        scratch["n_fine"] = torch.zeros(B, dtype=torch.int32, device=eng.device)
    m = mask.contiguous()
    eng.plan_masked(m, *batch.args(), out=scratch)
    if float(Settings.TICK_LENGTH) < float(Settings.T_DISCRETIZATION):
        eng.finer_fit_masked(m, scratch["s_seq"], scratch["reached_t"], batch.ego, scratch["fine"], scratch["n_fine"], speed)
    else:
        s = scratch["s_seq"]
        v = torch.where(scratch["reached_t"] >= 1, tdiv(s[:, 1] - s[:, 0], Settings.TICK_LENGTH), batch.ego[:, 2])
        speed.copy_(torch.where(m.bool(), v, speed))
    return scratch


def get_path_mean_abs_jerk(s_sequence, ego_start_speed, ego_start_acceleration, delta_t):
    """Reference st.py:274-288 (host bookkeeping on a short path)."""
    s = np.asarray(s_sequence, dtype=np.float64)
    v = np.diff(s) / delta_t
    a = np.diff(np.concatenate([[ego_start_speed], v])) / delta_t
    j = np.diff(np.concatenate([[ego_start_acceleration], a])) / delta_t
    total = 0.0
    for x in j:                       # sequential accumulation like the reference's loop
        total += abs(float(x))
    return total / (len(s) - 1)


def evaluate_st_and_dump_crash(num_episodes=1000, num_envs=None, csv_path="run_data.csv"):
    """Reference st.py:822-824: closed-loop evaluation of the pure MPC controller; prints the statistics and appends the
    run_data.csv row.  (The crash replay pickle of the reference is debug tooling and not reproduced.)"""
    from . import control
    hint = PlanHint()                               # the same environments are re-planned every tick
    output = control.evaluate_control(lambda state: do_st_control(state, hint=hint), num_episodes=num_episodes,
                                      num_envs=num_envs or Settings.EVAL_NUM_ENVS)
    output.print_stats(csv_path)
    return output
