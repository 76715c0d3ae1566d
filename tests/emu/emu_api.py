"""numpy driver of the EMULATED library (tests/emu/libmpc_emu_api.so): the C ABI of include/mpcb200.h with host arrays in
the place of device pointers.  TEST INFRASTRUCTURE ONLY -- the product binding is rl_mpc_lanemerging_b200/_lib.py + engine.py,
which never load this library."""
import ctypes as C
import os

import numpy as np

from rl_mpc_lanemerging_b200 import _lib
from tests.emu import build_emu

_emu = None


def lib():
    global _emu
    if _emu is None:
        L = C.CDLL(build_emu.build_api())
        for name, (res, args) in _lib.SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _emu = L
    return _emu


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def check(rc):
    if rc != 0:
        raise RuntimeError(f"emulated libmpcb200 error {rc}: {lib().mpc_last_error().decode()}")


class EmuEngine:
    """MpcEngine's planning calls on the emulated library.  env: MPC_FAST_* variables read by mpc_create (launch shape)."""

    def __init__(self, params, max_batch=64, nmax=32, env=None):
        self.L = lib()
        self.params, self.max_batch, self.nmax = params, max_batch, nmax
        old = {k: os.environ.get(k) for k in (env or {})}
        os.environ.update(env or {})
        try:
            h = C.c_void_p()
            check(self.L.mpc_create(C.byref(params), 0, max_batch, nmax, C.byref(h)))
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        self.h = h
        nt, ns = C.c_int(), C.c_int()
        check(self.L.mpc_grid_dims(h, C.byref(nt), C.byref(ns)))
        self.num_t, self.num_s_max = nt.value, ns.value

    def close(self):
        if self.h:
            self.L.mpc_destroy(self.h)
            self.h = None

    def _out(self, B):
        T = self.num_t
        return dict(idx=np.zeros((B, T), np.int32), s_seq=np.zeros((B, T)), cost=np.zeros(B), reached_t=np.zeros(B, np.int32),
                    crash=np.zeros(B, np.uint8), min_dist=np.zeros(B), start_s=np.zeros(B))

    def _outs(self, o):
        return [_p(o[k]) for k in ("idx", "s_seq", "cost", "reached_t", "crash", "min_dist", "start_s")]

    def counters(self):
        out = (C.c_int64 * 2)()
        check(self.L.mpc_last_counters(self.h, out))
        return {"kernels_launched": int(out[0]), "fallback_problems": int(out[1])}

    def fast32_info(self):
        out = (C.c_int64 * 6)()
        check(self.L.mpc_fast32_info(self.h, out))
        return {"in_use": bool(out[0]), "frac_bits": int(out[1]), "bound_fx": int(out[2]), "ring_cells": int(out[3]), "handed_on": int(out[4]),
                "first_shape_handed_on": int(out[5])}

    def plan(self, S, mode=0):
        B = S["ego"].shape[0]
        o = self._out(B)
        check(self.L.mpc_plan(self.h, B, _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["cars_a"]), _p(S["n_cars"]), mode,
                              *self._outs(o), None))
        return o

    def plan_hinted(self, S, hint_cost, hint_reached=None, hint_full_t=0, hint_scale=1.0, mode=0):
        B = S["ego"].shape[0]
        o = self._out(B)
        check(self.L.mpc_plan_hinted(self.h, B, _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["cars_a"]), _p(S["n_cars"]), mode,
                                     _p(hint_cost), _p(hint_reached), hint_full_t, hint_scale, *self._outs(o), None))
        return o

    def plan_probed(self, probe, S, margin=1.1, host=False):
        B = S["ego"].shape[0]
        o = self._out(B)
        fn = self.L.mpc_plan_host_probed if host else self.L.mpc_plan_probed
        check(fn(self.h, probe.h, margin, B, _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["cars_a"]), _p(S["n_cars"]),
                 *self._outs(o), None))
        return o

    def predict_step_without_ego(self, S, dt, mcd):
        B = S["ego"].shape[0]
        eo, xo, vo, ao = np.zeros((B, 4)), np.zeros((B, self.nmax)), np.zeros((B, self.nmax)), np.zeros((B, self.nmax))
        cr = np.zeros(B, np.uint8)
        check(self.L.mpc_predict_step_without_ego(self.h, B, _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["cars_a"]),
                                                  _p(S["n_cars"]), dt, mcd, _p(eo), _p(xo), _p(vo), _p(ao), _p(cr), None))
        return eo, xo, vo, ao, cr

    def build_grid(self, S, f32=False):
        B, T = S["ego"].shape[0], self.num_t
        stride = int(self.L.mpc_grid_stride(self.h))
        ob = np.zeros((B, T, stride), np.uint8)
        di = np.zeros((B, T, stride), np.float32 if f32 else np.float64)
        s0, ds, ns = np.zeros(B), np.zeros(B), np.zeros(B, np.int32)
        check(self.L.mpc_build_grid(self.h, B, _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["cars_a"]), _p(S["n_cars"]),
                                    _p(ob), _p(di), int(f32), _p(s0), _p(ds), _p(ns), None))
        return dict(obstacles=ob, distances=di, start_s=s0, delta_s=ds, num_s=ns)

    def solve_dense(self, g, v0, a0, mode=0):
        B, T, stride = g["obstacles"].shape
        o = dict(idx=np.zeros((B, T), np.int32), s_seq=np.zeros((B, T)), cost=np.zeros(B), reached_t=np.zeros(B, np.int32))
        v0, a0 = np.ascontiguousarray(v0), np.ascontiguousarray(a0)
        check(self.L.mpc_solve_dense(self.h, B, T, stride, _p(g["obstacles"]), _p(g["distances"]), int(g["distances"].dtype == np.float32),
                                     _p(g["start_s"]), _p(g["delta_s"]), _p(g["num_s"]), _p(v0), _p(a0), mode,
                                     _p(o["idx"]), _p(o["s_seq"]), _p(o["cost"]), _p(o["reached_t"]), None))
        return o

    def plan_host(self, S, mode=0):
        B = S["ego"].shape[0]
        o = self._out(B)
        check(self.L.mpc_plan_host(self.h, B, _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["cars_a"]), _p(S["n_cars"]), mode,
                                   *self._outs(o), None))
        return o

    def selftest_search(self, S):
        out = C.c_int64(-1)
        check(self.L.mpc_selftest_search(self.h, S["ego"].shape[0], _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["n_cars"]),
                                         C.byref(out), None))
        return int(out.value)

    def predict_step_with_ego(self, S, sel, dt, mcd):
        B = S["ego"].shape[0]
        eo, xo, vo, ao = np.zeros((B, 4)), np.zeros((B, self.nmax)), np.zeros((B, self.nmax)), np.zeros((B, self.nmax))
        cr = np.zeros(B, np.uint8)
        sel = np.ascontiguousarray(sel, np.float64)
        check(self.L.mpc_predict_step_with_ego(self.h, B, _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["cars_a"]),
                                               _p(S["n_cars"]), _p(sel), dt, mcd, _p(eo), _p(xo), _p(vo), _p(ao), _p(cr), None))
        return eo, xo, vo, ao, cr

    def state_vector(self, S):
        B = S["ego"].shape[0]
        out = np.zeros((B, 21), np.float32)
        check(self.L.mpc_state_vector(self.h, B, _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["cars_a"]), _p(S["n_cars"]),
                                      _p(out), 21, None))
        return out

    def speed_from_jerk(self, S, jerk):
        B = S["ego"].shape[0]
        out = np.zeros(B)
        jerk = np.ascontiguousarray(jerk, np.float64)
        check(self.L.mpc_speed_from_jerk(self.h, B, _p(S["ego"]), _p(jerk), _p(out), None))
        return out

    def finer_fit(self, s_seq, reached_t, ego):
        B = s_seq.shape[0]
        nf = int(self.L.mpc_finer_fit_max_points())
        fine, n_fine, speed, iters = np.zeros((B, nf)), np.zeros(B, np.int32), np.zeros(B), np.zeros(B, np.int32)
        check(self.L.mpc_finer_fit(self.h, B, _p(s_seq), _p(reached_t), _p(ego), _p(fine), nf, _p(n_fine), _p(speed), _p(iters), None))
        return fine, n_fine, speed, iters

    def plan_masked(self, S, mask, out, mode=0):
        B = S["ego"].shape[0]
        check(self.L.mpc_plan_masked(self.h, B, _p(mask), _p(S["ego"]), _p(S["cars_x"]), _p(S["cars_v"]), _p(S["cars_a"]), _p(S["n_cars"]),
                                     mode, _p(out["idx"]), _p(out["s_seq"]), _p(out["cost"]), _p(out["reached_t"]), _p(out["crash"]),
                                     _p(out["min_dist"]), None, None))
        return out

    def finer_fit_masked(self, mask, s_seq, reached_t, ego, fine, n_fine, speed):
        check(self.L.mpc_finer_fit_masked(self.h, s_seq.shape[0], _p(mask), _p(s_seq), _p(reached_t), _p(ego), _p(fine), fine.shape[1],
                                          _p(n_fine), _p(speed), None, None))
