"""ctypes binding of the C oracle (oracle/mpc_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (rl_mpc_lanemerging_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmpc_oracle.so")
MAX_CARS = 64


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "s_disc", "t_disc", "future_s", "future_t", "start_uncertainty", "uncertainty_per_second",
        "d_weight", "v_weight", "a_weight", "j_weight", "desired_speed", "max_speed",
        "a_min", "a_max", "j_min", "j_max", "min_allowed_distance", "crash_min_s", "car_length",
        "max_predicted_decel", "tick_length", "sensor_radius", "combination_min_distance")]


class State(C.Structure):
    _fields_ = [("ego_x", C.c_double), ("ego_y", C.c_double), ("ego_v", C.c_double), ("ego_a", C.c_double),
                ("n", C.c_int), ("x", C.c_double * MAX_CARS), ("v", C.c_double * MAX_CARS),
                ("a", C.c_double * MAX_CARS)]


class Stats(C.Structure):
    _fields_ = [("pops", C.c_long), ("pushes", C.c_long), ("nodes", C.c_long), ("edges", C.c_long),
                ("coast_collisions", C.c_long), ("max_span", C.c_int), ("max_width", C.c_int),
                ("wmult_hist", C.c_long * 8)]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "mpc_oracle.c")
    hdr = os.path.join(HERE, "mpc_oracle.h")
    if os.path.exists(src) and (force or not os.path.exists(LIB) or
                                os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                               "-Wall", "-Wno-misleading-indentation", src, "-o", LIB, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_get_ego_s.restype = C.c_double
        L.orc_get_ego_s.argtypes = [C.c_double, C.c_double]
        L.orc_speed_from_jerk.restype = C.c_double
        L.orc_speed_from_jerk.argtypes = [C.POINTER(Params), C.c_double, C.c_double, C.c_double]
        L.orc_path_mean_abs_jerk.restype = C.c_double
        L.orc_path_cost.restype = C.c_double
    return _lib


def default_params(**over) -> Params:
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, float(v))
    return p


def horizon_params(H: int, **over) -> Params:
    """H+1 layers, FUTURE_S = MAX_SPEED*FUTURE_T kept as in the reference."""
    if H == 17:
        return default_params(**over)
    return default_params(future_t=0.3 * H - 0.15, future_s=30.0 * 0.3 * H, **over)


def make_state(ego_position, ego_speed, ego_acc, xs, vs, accs) -> State:
    s = State()
    s.ego_x, s.ego_y = float(ego_position[0]), float(ego_position[1])
    s.ego_v, s.ego_a, s.n = float(ego_speed), float(ego_acc), len(xs)
    for i in range(len(xs)):
        s.x[i], s.v[i], s.a[i] = float(xs[i]), float(vs[i]), float(accs[i])
    return s


def state_tuple(s: State):
    n = s.n
    return ((s.ego_x, s.ego_y), s.ego_v, s.ego_a, list(s.x[:n]), list(s.v[:n]), list(s.a[:n]))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def grid_dims(p: Params, start_s: float):
    nt, ns, ds = C.c_int(), C.c_int(), C.c_double()
    lib().orc_grid_dims(C.byref(p), C.c_double(start_s), C.byref(nt), C.byref(ns), C.byref(ds))
    return nt.value, ns.value, ds.value


def get_ego_s(x, y):
    return lib().orc_get_ego_s(x, y)


def predict_step_with_ego(p, st: State, speed, dt, min_crash=5.0):
    out = State()
    crashed = lib().orc_predict_step_with_ego(C.byref(p), C.byref(st), C.c_double(speed), C.c_double(dt),
                                              C.c_double(min_crash), C.byref(out))
    return out, bool(crashed)


def predict_step_without_ego(p, st: State, dt, min_crash=5.0):
    out = State()
    crashed = lib().orc_predict_step_without_ego(C.byref(p), C.byref(st), C.c_double(dt),
                                                 C.c_double(min_crash), C.byref(out))
    return out, bool(crashed)


def build_grid(p, st: State, want_obs_s=False):
    start_s = get_ego_s(st.ego_x, st.ego_y)
    nt, ns, _ = grid_dims(p, start_s)
    obstacles = np.zeros((nt, ns), np.uint8)
    distances = np.zeros((nt, ns), np.float64)
    s_values = np.zeros(ns, np.float64)
    obs_s = np.zeros((nt, MAX_CARS), np.float64) if want_obs_s else None
    lib().orc_build_grid(C.byref(p), C.byref(st), nt, ns, obstacles.ctypes.data_as(C.c_void_p),
                         _dp(distances), _dp(s_values), _dp(obs_s) if want_obs_s else None)
    if want_obs_s:
        return obstacles, distances, s_values, obs_s
    return obstacles, distances, s_values


def solve(p, obstacles, distances, s_values, delta_t, v0, a0, layered=False):
    nt, ns = obstacles.shape
    idx = np.zeros(nt, np.int32)
    seq = np.zeros(nt, np.float64)
    cost = C.c_double()
    stats = Stats()
    fn = lib().orc_solve_layered if layered else lib().orc_solve_dijkstra
    obstacles = np.ascontiguousarray(obstacles, dtype=np.uint8)
    distances = np.ascontiguousarray(distances, dtype=np.float64)
    s_values = np.ascontiguousarray(s_values, dtype=np.float64)
    r = fn(C.byref(p), nt, ns, obstacles.ctypes.data_as(C.c_void_p), _dp(distances), _dp(s_values),
           C.c_double(delta_t), C.c_double(v0), C.c_double(a0), _ip(idx), _dp(seq), C.byref(cost), C.byref(stats))
    return dict(reached_t=r, idx=idx, s_seq=seq, cost=cost.value, stats=stats)


def solve_fast_model(p, obstacles, distances, s_values, delta_t, v0, a0, f32_labels=False):
    """CPU model of the CUDA fast kernel (integer kinematics, 2^-18 fixed-point labels; f32_labels=True = the rejected
    fp32-label design, kept to document why); NOT the reference."""
    nt, ns = obstacles.shape
    idx = np.zeros(nt, np.int32)
    seq = np.zeros(nt, np.float64)
    cost = C.c_double()
    obstacles = np.ascontiguousarray(obstacles, dtype=np.uint8)
    distances = np.ascontiguousarray(distances, dtype=np.float64)
    s_values = np.ascontiguousarray(s_values, dtype=np.float64)
    r = lib().orc_solve_fast_model(C.byref(p), nt, ns, obstacles.ctypes.data_as(C.c_void_p), _dp(distances), _dp(s_values),
                                   C.c_double(delta_t), C.c_double(v0), C.c_double(a0), C.c_int(int(f32_labels)), _ip(idx), _dp(seq), C.byref(cost))
    return dict(reached_t=r, idx=idx, s_seq=seq, cost=cost.value)


def solve_fast_model_ex(p, obstacles, distances, s_values, delta_t, v0, a0, prune_cost=0.0, prune_fx=None):
    """Fast-kernel model with the cost bound (nodes with a label above prune_cost are dropped; 0 = none; prune_fx gives the
    bound in label units, 2^-18, exactly as the kernel holds it).  Also returns the node / push counts."""
    nt, ns = obstacles.shape
    idx = np.zeros(nt, np.int32); seq = np.zeros(nt, np.float64); cost = C.c_double(); counts = (C.c_int64 * 2)()
    obstacles = np.ascontiguousarray(obstacles, dtype=np.uint8)
    distances = np.ascontiguousarray(distances, dtype=np.float64)
    s_values = np.ascontiguousarray(s_values, dtype=np.float64)
    fxb = int(prune_fx) if prune_fx is not None else (int(round(prune_cost * 262144.0)) if prune_cost else 0)
    r = lib().orc_solve_fast_model_ex(C.byref(p), nt, ns, obstacles.ctypes.data_as(C.c_void_p), _dp(distances), _dp(s_values),
                                      C.c_double(delta_t), C.c_double(v0), C.c_double(a0), C.c_int(0), C.c_uint64(fxb),
                                      _ip(idx), _dp(seq), C.byref(cost), counts)
    return dict(reached_t=r, idx=idx, s_seq=seq, cost=cost.value, nodes=int(counts[0]), pushes=int(counts[1]))


def solve_fast_model_q(p, obstacles, distances, s_values, delta_t, v0, a0, frac_bits, key_shift, prune_fx=0):
    """Model of the 32-bit-key fast kernel: 2^-frac_bits labels, min-combine on (label >> key_shift, larger v')."""
    nt, ns = obstacles.shape
    idx = np.zeros(nt, np.int32); seq = np.zeros(nt, np.float64); cost = C.c_double(); counts = (C.c_int64 * 2)()
    obstacles = np.ascontiguousarray(obstacles, dtype=np.uint8)
    distances = np.ascontiguousarray(distances, dtype=np.float64)
    s_values = np.ascontiguousarray(s_values, dtype=np.float64)
    r = lib().orc_solve_fast_model_q(C.byref(p), nt, ns, obstacles.ctypes.data_as(C.c_void_p), _dp(distances), _dp(s_values),
                                     C.c_double(delta_t), C.c_double(v0), C.c_double(a0), C.c_int(frac_bits), C.c_int(key_shift),
                                     C.c_uint64(int(prune_fx)), _ip(idx), _dp(seq), C.byref(cost), counts)
    return dict(reached_t=r, idx=idx, s_seq=seq, cost=cost.value, nodes=int(counts[0]), pushes=int(counts[1]))


def solve_fast_ladder(p, obstacles, distances, s_values, delta_t, v0, a0, frac_bits, bound_fx):
    """The fast mode of mpc_plan as the device runs it (run_solve, mpc_api.cu): the 32-bit-key kernel's bounded attempt
    (2^-frac_bits labels under bound_fx; both from mpc_fast32_info / derive_params); when that does not reach the horizon,
    the 64-bit kernel's unbounded pass (2^-18 labels).  Returns the model's answer plus `stage` (32 or 64)."""
    H = obstacles.shape[0] - 1
    if frac_bits:
        r = solve_fast_model_q(p, obstacles, distances, s_values, delta_t, v0, a0, frac_bits, 0, bound_fx)
        if r["reached_t"] == H:
            r["stage"] = 32
            return r
    r = solve_fast_model(p, obstacles, distances, s_values, delta_t, v0, a0)
    r["stage"] = 64
    return r


def path_cost(p, idx, s_values, distances, delta_t, v0, a0):
    idx = np.ascontiguousarray(idx, np.int32)
    distances = np.ascontiguousarray(distances, np.float64)
    return lib().orc_path_cost(C.byref(p), C.c_int(len(idx)), _ip(idx), _dp(np.ascontiguousarray(s_values)),
                               _dp(distances), C.c_int(distances.shape[1]), C.c_double(delta_t),
                               C.c_double(v0), C.c_double(a0))


def plan(p, st: State, layered=False):
    start_s = get_ego_s(st.ego_x, st.ego_y)
    nt, ns, _ = grid_dims(p, start_s)
    idx = np.zeros(nt, np.int32)
    seq = np.zeros(nt, np.float64)
    cost, mind, s0, ds = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    crash, nso = C.c_int(), C.c_int()
    stats = Stats()
    r = lib().orc_plan(C.byref(p), C.byref(st), int(layered), _ip(idx), _dp(seq), C.byref(cost), C.byref(crash),
                       C.byref(mind), C.byref(s0), C.byref(ds), C.byref(nso), C.byref(stats))
    return dict(reached_t=r, idx=idx, s_seq=seq, cost=cost.value, crash=bool(crash.value),
                min_path_distance=mind.value, start_s=s0.value, delta_s=ds.value, num_s=nso.value, stats=stats)


def plan_batch(p, ego, cars_x, cars_v, cars_a, n_cars, num_t, layered=True, nthreads=1):
    """ego [B,4] f64; cars_* [B,nmax] f64; n_cars [B] i32.  Returns dict of arrays."""
    ego = np.ascontiguousarray(ego, np.float64)
    cars_x = np.ascontiguousarray(cars_x, np.float64)
    cars_v = np.ascontiguousarray(cars_v, np.float64)
    cars_a = np.ascontiguousarray(cars_a, np.float64)
    n_cars = np.ascontiguousarray(n_cars, np.int32)
    B, nmax = cars_x.shape
    idx = np.zeros((B, num_t), np.int32)
    seq = np.zeros((B, num_t), np.float64)
    cost = np.zeros(B, np.float64)
    reached = np.zeros(B, np.int32)
    crash = np.zeros(B, np.int32)
    mind = np.zeros(B, np.float64)
    lib().orc_plan_batch(C.byref(p), B, nmax, _dp(ego), _dp(cars_x), _dp(cars_v), _dp(cars_a), _ip(n_cars),
                         int(layered), int(nthreads), int(num_t), _ip(idx), _dp(seq), _dp(cost), _ip(reached),
                         _ip(crash), _dp(mind))
    return dict(idx=idx, s_seq=seq, cost=cost, reached_t=reached, crash=crash.astype(bool), min_path_distance=mind)


def speed_from_jerk(p, v, a, jerk):
    return lib().orc_speed_from_jerk(C.byref(p), C.c_double(v), C.c_double(a), C.c_double(jerk))


def state_vector(p, st: State):
    out = np.zeros(20, np.float64)
    lib().orc_state_vector(C.byref(p), C.byref(st), _dp(out))
    return out


def path_mean_abs_jerk(s, v0, a0, dt):
    s = np.ascontiguousarray(s, np.float64)
    return lib().orc_path_mean_abs_jerk(_dp(s), C.c_int(len(s)), C.c_double(v0), C.c_double(a0), C.c_double(dt))
