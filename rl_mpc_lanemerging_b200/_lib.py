"""ctypes binding of libmpcb200.so (include/mpcb200.h).  Thin: pointers and sizes only.

The library is the product.  If it is missing this module raises -- there is no CPU or PyTorch
fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmpcb200.so")

MODE_FAST, MODE_EXACT = 0, 1
E_INVALID, E_CUDA, E_CAPACITY, E_NODEVICE = -1, -2, -3, -4

PARAM_FIELDS = ("s_disc", "t_disc", "future_s", "future_t", "start_uncertainty", "uncertainty_per_second",
                "d_weight", "v_weight", "a_weight", "j_weight", "desired_speed", "max_speed",
                "a_min", "a_max", "j_min", "j_max", "min_allowed_distance", "crash_min_s", "car_length",
                "max_predicted_decel", "tick_length", "sensor_radius", "combination_min_distance")


class MpcParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in PARAM_FIELDS]


ENV_PARAM_FIELDS = ("tick", "a_min", "a_max", "max_speed", "min_crash_distance", "sensor_radius", "spawn_x", "other_speed", "interval",
                    "arrival_x", "ego_start_x", "ego_start_y", "start_speed", "start_speed_var", "min_start_speed", "max_start_speed",
                    "time_reward_step", "jerk_weight", "crash_reward", "success_reward", "invalid_action_step",
                    "krauss_accel", "krauss_decel", "krauss_tau", "krauss_min_gap")


class MpcEnvParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ENV_PARAM_FIELDS] + [("max_ticks", C.c_int32), ("auto_reset", C.c_int32), ("world", C.c_int32),
                                                              ("pad_", C.c_int32)]


class MpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmpcb200 error {code}: {msg}")
        self.code = code


# every symbol include/mpcb200.h declares: (name, restype, argtypes)
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
SYMBOLS = {
    "mpc_abi_version": (_i, []),
    "mpc_last_error": (C.c_char_p, []),
    "mpc_default_params": (None, [C.POINTER(MpcParams)]),
    "mpc_device_count": (_i, []),
    "mpc_create": (_i, [C.POINTER(MpcParams), _i, _i, _i, C.POINTER(_vp)]),
    "mpc_set_params": (_i, [_vp, C.POINTER(MpcParams)]),
    "mpc_destroy": (_i, [_vp]),
    "mpc_grid_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "mpc_grid_stride": (_i, [_vp]),
    "mpc_last_counters": (_i, [_vp, C.POINTER(C.c_int64)]),
    "mpc_fast32_info": (_i, [_vp, C.POINTER(C.c_int64)]),
    "mpc_krauss_step": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _d, _d, _d, _vp, _vp]),
    "mpc_set_timing": (_i, [_vp, _i]),
    "mpc_last_kernel_ms": (_i, [_vp, C.POINTER(C.c_float)]),
    "mpc_selftest_search": (_i, [_vp, _i, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64), _vp]),
    "mpc_build_grid": (_i, [_vp, _i] + [_vp] * 5 + [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "mpc_solve_dense": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "mpc_plan": (_i, [_vp, _i] + [_vp] * 5 + [_i] + [_vp] * 7 + [_vp]),
    "mpc_plan_masked": (_i, [_vp, _i, _vp] + [_vp] * 5 + [_i] + [_vp] * 7 + [_vp]),
    "mpc_finer_fit_masked": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "mpc_plan_hinted": (_i, [_vp, _i] + [_vp] * 5 + [_i] + [_vp, _vp, _i, _d] + [_vp] * 7 + [_vp]),
    "mpc_plan_probed": (_i, [_vp, _vp, _d, _i] + [_vp] * 5 + [_vp] * 7 + [_vp]),
    "mpc_plan_host": (_i, [_vp, _i] + [_vp] * 5 + [_i] + [_vp] * 7 + [_vp]),
    "mpc_plan_host_probed": (_i, [_vp, _vp, _d, _i] + [_vp] * 5 + [_vp] * 7 + [_vp]),
    "mpc_finer_fit": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "mpc_finer_fit_max_points": (_i, []),
    "mpc_predict_step_with_ego": (_i, [_vp, _i] + [_vp] * 6 + [_d, _d] + [_vp] * 5 + [_vp]),
    "mpc_predict_step_without_ego": (_i, [_vp, _i] + [_vp] * 5 + [_d, _d] + [_vp] * 5 + [_vp]),
    "mpc_state_vector": (_i, [_vp, _i] + [_vp] * 5 + [_vp, _i, _vp]),
    "mpc_speed_from_jerk": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "mpc_env_step": (_i, [_vp, C.POINTER(MpcEnvParams), _i] + [_vp] * 17 + [_vp]),
    "mpc_rollout_step": (_i, [_vp, _i] + [_vp] * 6 + [_d, _d, _d, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Loads the shared library (building is __graft_entry__.build()'s job) and types its symbols."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                              "the hot path has no fallback implementation")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the header and the binary disagree
            fn.restype = res
            fn.argtypes = args
        if lib.mpc_abi_version() != 1:
            raise ImportError("libmpcb200 ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise MpcError(rc, load().mpc_last_error().decode())
    return rc


def default_params() -> MpcParams:
    p = MpcParams()
    load().mpc_default_params(C.byref(p))
    return p
