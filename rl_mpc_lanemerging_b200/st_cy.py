"""Drop-in for the reference's only native module, the Cython extension `st_cy` (reference st_cy.pyx).

`solve_s_t_path_fast` keeps the reference's positional signature (st_cy.pyx:315) and return value (a
float64 array of length num_t whose unreached tail is 0.0), but the Dijkstra runs as the layered-DP
CUDA kernel behind `mpc_solve_dense` (K2, exact fp64 mode: index-identical to the Cython result).
The non-default solvers of the reference module (no-jerk variants) are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import MpcEngine, params_key

_engines = {}
_MAX_ENGINES = 4            # one per (Settings snapshot, device); the least recently used one is closed beyond that


def _engine_for(p, device=0) -> MpcEngine:
    key = (params_key(p), device)
    e = _engines.pop(key, None)
    if e is None:
        e = MpcEngine(p, device=device, max_batch=16)
        while len(_engines) >= _MAX_ENGINES:
            _engines.pop(next(iter(_engines))).close()
    _engines[key] = e                                         # (re-inserted: most recently used last)
    return e


def solve_s_t_path_fast(obstacles_bool, s_values, t_indices, ego_start_speed, ego_start_acceleration, distances,
                        d_weight, v_weight, a_weight, j_weight, desired_speed, max_speed, negative_acceleration_limit,
                        positive_acceleration_limit, negative_jerk_limit, positive_jerk_limit, min_allowed_distance,
                        mode="exact", device=0):
    obstacles = np.ascontiguousarray(obstacles_bool).view(np.uint8) if obstacles_bool.dtype == np.bool_ \
        else np.ascontiguousarray(obstacles_bool, dtype=np.uint8)
    s_values = np.asarray(s_values, dtype=np.float64)
    t_indices = np.asarray(t_indices, dtype=np.float64)
    distances = np.ascontiguousarray(distances, dtype=np.float64)
    num_t, num_s = obstacles.shape
    if distances.shape != (num_t, num_s) or s_values.shape != (num_s,) or t_indices.shape != (num_t,):
        raise IndexError("obstacles / distances / s_values / t_indices shapes disagree")
    start_s, delta_s = float(s_values[0]), float(s_values[1] - s_values[0])
    delta_t = float(t_indices[1] - t_indices[0])
    # the kernel regenerates s_values from (start, delta, n) the way numpy.arange fills them (st.py:31)
    if not np.array_equal(s_values[2:], start_s + np.arange(2, num_s) * delta_s):
        raise ValueError("s_values must be an arithmetic grid as produced by numpy.arange (reference st.py:31)")
    p = _lib.default_params()
    p.t_disc = delta_t
    p.future_t = float(t_indices[-1])                     # arange(0, T + dt, dt) has num_t points when T = t[-1]
    p.s_disc = delta_s
    p.future_s = float(s_values[-1] - start_s)
    p.d_weight, p.v_weight, p.a_weight, p.j_weight = float(d_weight), float(v_weight), float(a_weight), float(j_weight)
    p.desired_speed, p.max_speed = float(desired_speed), float(max_speed)
    p.a_min, p.a_max = float(negative_acceleration_limit), float(positive_acceleration_limit)
    p.j_min, p.j_max = float(negative_jerk_limit), float(positive_jerk_limit)
    p.min_allowed_distance = float(min_allowed_distance)
    eng = _engine_for(p, device)
    if eng.num_t != num_t:
        # nudge FUTURE_T so that arange() yields exactly num_t layers
        p.future_t = delta_t * (num_t - 1) - 0.5 * delta_t
        eng = _engine_for(p, device)
    if eng.num_t != num_t or eng.num_s_max < num_s:
        raise ValueError("grid larger than the engine derived from it")
    dev = eng.device
    S = eng.num_s_max
    ob = torch.ones((1, num_t, S), dtype=torch.uint8, device=dev)
    di = torch.zeros((1, num_t, S), dtype=torch.float64, device=dev)
    ob[0, :, :num_s] = torch.from_numpy(obstacles).to(dev)
    di[0, :, :num_s] = torch.from_numpy(distances).to(dev)
    f = lambda x, dt=torch.float64: torch.tensor([x], dtype=dt, device=dev)  # noqa: E731
    out = eng.solve_dense(ob, di, f(start_s), f(delta_s), f(num_s, torch.int32), f(float(ego_start_speed)),
                          f(float(ego_start_acceleration)), mode=mode)
    return out["s_seq"][0].cpu().numpy()
