"""Batched planner engine: the Python face of libmpcb200 (one engine per GPU / stream).

PyTorch is used for device memory and streams only; every computation on the hot path is a
hand-written CUDA kernel reached through the C ABI (include/mpcb200.h).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import MODE_EXACT, MODE_FAST, MpcParams


def params_from_settings(S) -> MpcParams:
    """Snapshot of the reference-style global Settings (reference config.py:30-37,94-110,143,153)."""
    p = MpcParams()
    p.s_disc, p.t_disc = float(S.S_DISCRETIZATION), float(S.T_DISCRETIZATION)
    p.future_s, p.future_t = float(S.FUTURE_S), float(S.FUTURE_T)
    p.start_uncertainty, p.uncertainty_per_second = float(S.START_UNCERTAINTY), float(S.UNCERTAINTY_PER_SECOND)
    p.d_weight, p.v_weight, p.a_weight, p.j_weight = (float(S.D_WEIGHT), float(S.V_WEIGHT), float(S.A_WEIGHT),
                                                      float(S.J_WEIGHT))
    p.desired_speed, p.max_speed = float(S.DESIRED_SPEED), float(S.MAX_SPEED)
    p.a_min, p.a_max = float(S.MAX_NEGATIVE_ACCELERATION), float(S.MAX_POSITIVE_ACCELERATION)
    p.j_min, p.j_max = float(S.MINIMUM_NEGATIVE_JERK), float(S.MAXIMUM_POSITIVE_JERK)
    p.min_allowed_distance, p.crash_min_s = float(S.MIN_ALLOWED_DISTANCE), float(S.CRASH_MIN_S)
    p.car_length = float(S.CAR_LENGTH)
    p.max_predicted_decel = float(S.MAX_PREDICTED_DECELERATION)
    p.tick_length, p.sensor_radius = float(S.TICK_LENGTH), float(S.SENSOR_RADIUS)
    p.combination_min_distance = float(S.COMBINATION_MIN_DISTANCE)
    return p


def params_key(p: MpcParams):
    return tuple(getattr(p, n) for n in _lib.PARAM_FIELDS)


def _mode(mode) -> int:
    if mode in (MODE_FAST, "fast"):
        return MODE_FAST
    if mode in (MODE_EXACT, "exact"):
        return MODE_EXACT
    raise ValueError(f"unknown mode {mode!r}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class MpcEngine:
    """Owns one mpc_handle.  All tensor arguments live on `device`; states are fp64:
    ego [B,4] = (x, y, speed, acceleration); cars_x/v/a [B,nmax]; n_cars [B] int32."""

    _pin_host = True             # host staging buffers of plan_host are page-locked

    def __init__(self, params: Optional[MpcParams] = None, device: int | str | torch.device = 0,
                 max_batch: int = 4096, nmax: int = 32):
        self.lib = _lib.load()
        if not torch.cuda.is_available() or self.lib.mpc_device_count() <= 0:
            raise _lib.MpcError(_lib.E_NODEVICE, "no CUDA device: the MPC hot path has no CPU fallback")
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        self.device = dev
        self.dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self.params = params if params is not None else _lib.default_params()
        self.max_batch, self.nmax = int(max_batch), int(nmax)
        torch.cuda.init()
        with self._device_ctx():
            torch.zeros(1, device=dev)            # make sure the primary context exists before the library uses it
            h = C.c_void_p()
            _lib.check(self.lib.mpc_create(C.byref(self.params), self.dev_index, self.max_batch, self.nmax, C.byref(h)))
        self.h = h
        nt, ns = C.c_int(), C.c_int()
        _lib.check(self.lib.mpc_grid_dims(self.h, C.byref(nt), C.byref(ns)))
        self.num_t, self.num_s_max = nt.value, ns.value
        self.num_s_stride = int(self.lib.mpc_grid_stride(self.h))     # row stride of dense grids (multiple of 8)
        self._pinned = {}

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.mpc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _is_dev(t) -> bool:
        """Tensor arguments must live in device memory: the library dereferences their data_ptr() on the GPU."""
        return t.is_cuda

    def _device_ctx(self):
        """Context in which the library is called: the handle's device made current."""
        return torch.cuda.device(self.dev_index)

    def _check_state(self, ego, cars_x, cars_v, cars_a, n_cars):
        B = ego.shape[0]
        assert ego.dtype == torch.float64 and ego.shape == (B, 4) and ego.is_contiguous() and self._is_dev(ego)
        for c in (cars_x, cars_v) + ((cars_a,) if cars_a is not None else ()):
            assert c.dtype == torch.float64 and c.shape == (B, self.nmax) and c.is_contiguous() and self._is_dev(c)
        assert n_cars.dtype == torch.int32 and n_cars.shape == (B,) and n_cars.is_contiguous() and self._is_dev(n_cars)
        return B

    def counters(self):
        out = (C.c_int64 * 2)()
        _lib.check(self.lib.mpc_last_counters(self.h, out))
        return {"kernels_launched": int(out[0]), "fallback_problems": int(out[1])}

    def fast32_info(self):
        """The 32-bit-key DP kernel of the fast mode's first attempt: in use?, label fraction bits, cost bound (label units),
        ring capacity of its first launch shape (cells), problems of the last call it handed to the 64-bit kernel / its first shape handed on."""
        out = (C.c_int64 * 6)()
        _lib.check(self.lib.mpc_fast32_info(self.h, out))
        return {"in_use": bool(out[0]), "frac_bits": int(out[1]), "bound_fx": int(out[2]), "ring_cells": int(out[3]),
                "handed_on": int(out[4]), "first_shape_handed_on": int(out[5])}

    def selftest_search(self, ego, cars_x, cars_v, cars_a, n_cars) -> int:
        """Cells whose sorted-structure lookup differs from the reference-order evaluation (must be 0)."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        out = C.c_int64(-1)
        with self._device_ctx():
            _lib.check(self.lib.mpc_selftest_search(self.h, B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(n_cars), C.byref(out), self._stream()))
        return int(out.value)

    def set_timing(self, enable=True):
        _lib.check(self.lib.mpc_set_timing(self.h, int(enable)))

    def last_kernel_ms(self):
        """(predictor, DP kernel, fallback DP) device milliseconds of the last plan()/solve_dense() call."""
        out = (C.c_float * 3)()
        _lib.check(self.lib.mpc_last_kernel_ms(self.h, out))
        return float(out[0]), float(out[1]), float(out[2])

    # ------------------------------------------------------------------------------------------
    def plan(self, ego, cars_x, cars_v, cars_a, n_cars, mode="fast", out: Optional[dict] = None):
        """Fused gap-evaluation (K3).  Returns dict(idx i32[B,T], s_seq f64[B,T], cost f64[B], reached_t i32[B],
        crash u8[B], min_dist f64[B], start_s f64[B]) of device tensors (re-used when `out` is passed)."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        T = self.num_t
        if out is None:
            o = dict(device=self.device)
            out = dict(idx=torch.empty((B, T), dtype=torch.int32, **o), s_seq=torch.empty((B, T), dtype=torch.float64, **o),
                       cost=torch.empty(B, dtype=torch.float64, **o), reached_t=torch.empty(B, dtype=torch.int32, **o),
                       crash=torch.empty(B, dtype=torch.uint8, **o), min_dist=torch.empty(B, dtype=torch.float64, **o),
                       start_s=torch.empty(B, dtype=torch.float64, **o))
        with self._device_ctx():
            _lib.check(self.lib.mpc_plan(self.h, B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a), _ptr(n_cars),
                                         _mode(mode), _ptr(out["idx"]), _ptr(out["s_seq"]), _ptr(out["cost"]),
                                         _ptr(out["reached_t"]), _ptr(out["crash"]), _ptr(out["min_dist"]),
                                         _ptr(out["start_s"]), self._stream()))
        return out

    def plan_masked(self, mask, ego, cars_x, cars_v, cars_a, n_cars, out: dict, mode="fast"):
        """mpc_plan_masked: plans only the episodes with mask[b] != 0 (u8 / bool [B]); rows of `out` (idx, s_seq, cost, reached_t,
        crash, min_dist) of the other episodes are left untouched.  No host round trip."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        m = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
        assert m.dtype == torch.uint8 and m.shape == (B,) and m.is_contiguous() and self._is_dev(m)
        with self._device_ctx():
            _lib.check(self.lib.mpc_plan_masked(self.h, B, _ptr(m), _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a), _ptr(n_cars),
                                                _mode(mode), _ptr(out["idx"]), _ptr(out["s_seq"]), _ptr(out["cost"]),
                                                _ptr(out["reached_t"]), _ptr(out["crash"]), _ptr(out["min_dist"]), None, self._stream()))
        return out

    def finer_fit_masked(self, mask, s_seq, reached_t, ego, fine, n_fine, speed):
        """mpc_finer_fit_masked: st.finer_fit for the episodes with mask[b] != 0; rows of fine / n_fine / speed of the others untouched."""
        B = s_seq.shape[0]
        m = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
        assert m.dtype == torch.uint8 and m.shape == (B,) and m.is_contiguous() and fine.is_contiguous()
        with self._device_ctx():
            _lib.check(self.lib.mpc_finer_fit_masked(self.h, B, _ptr(m), _ptr(s_seq), _ptr(reached_t), _ptr(ego), _ptr(fine),
                                                     fine.shape[1], _ptr(n_fine), _ptr(speed), None, self._stream()))

    def plan_hinted(self, ego, cars_x, cars_v, cars_a, n_cars, hint_cost, hint_reached=None, hint_full_t=0, hint_scale=1.0,
                    mode="fast", out: Optional[dict] = None):
        """plan() with a per-episode cost hint (mpc_plan_hinted): episode b is first solved under the bound
        hint_scale * hint_cost[b] (when hint_reached is None or hint_reached[b] == hint_full_t).  The hint is an
        ESTIMATE -- a coarse probe plan, or the previous tick's plan of a closed-loop controller; the outputs are
        identical to plan()'s whatever the hint, only the number of expanded nodes changes."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        assert hint_cost.dtype == torch.float64 and hint_cost.shape == (B,) and hint_cost.is_contiguous() and self._is_dev(hint_cost)
        if hint_reached is not None:
            assert hint_reached.dtype == torch.int32 and hint_reached.shape == (B,) and hint_reached.is_contiguous() and self._is_dev(hint_reached)
        out = self._plan_out(B) if out is None else out
        with self._device_ctx():
            _lib.check(self.lib.mpc_plan_hinted(self.h, B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a), _ptr(n_cars),
                                                _mode(mode), _ptr(hint_cost), _ptr(hint_reached), int(hint_full_t),
                                                float(hint_scale), _ptr(out["idx"]), _ptr(out["s_seq"]), _ptr(out["cost"]),
                                                _ptr(out["reached_t"]), _ptr(out["crash"]), _ptr(out["min_dist"]),
                                                _ptr(out["start_s"]), self._stream()))
        return out

    def make_probe(self, s_mult: int = 20, t_mult: int = 3) -> "MpcEngine":
        """A second engine on the same device whose grid is s_mult x t_mult coarser (same Settings otherwise): its
        plans are the cost estimates of plan_probed().  20 x 3 keeps the jerk resolution (ds/dt^3) of the published
        discretisation; H=50 -> an 18 x 451 grid instead of 51 x 9001."""
        p = MpcParams()
        for n in _lib.PARAM_FIELDS:
            setattr(p, n, getattr(self.params, n))
        p.s_disc, p.t_disc = self.params.s_disc * s_mult, self.params.t_disc * t_mult
        return MpcEngine(p, device=self.device, max_batch=self.max_batch, nmax=self.nmax)

    def plan_probed(self, probe: "MpcEngine", ego, cars_x, cars_v, cars_a, n_cars, margin: float = 1.1,
                    out: Optional[dict] = None):
        """Probe + plan (mpc_plan_probed, fast mode): the states are planned on `probe`'s coarse grid first and
        margin * (num_t-1)/(probe.num_t-1) * probe cost bounds the first attempt of the real solve.  Same outputs
        as plan()."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        out = self._plan_out(B) if out is None else out
        with self._device_ctx():
            _lib.check(self.lib.mpc_plan_probed(self.h, probe.h, float(margin), B, _ptr(ego), _ptr(cars_x), _ptr(cars_v),
                                                _ptr(cars_a), _ptr(n_cars), _ptr(out["idx"]), _ptr(out["s_seq"]),
                                                _ptr(out["cost"]), _ptr(out["reached_t"]), _ptr(out["crash"]),
                                                _ptr(out["min_dist"]), _ptr(out["start_s"]), self._stream()))
        return out

    def _plan_out(self, B):
        o, T = dict(device=self.device), self.num_t
        return dict(idx=torch.empty((B, T), dtype=torch.int32, **o), s_seq=torch.empty((B, T), dtype=torch.float64, **o),
                    cost=torch.empty(B, dtype=torch.float64, **o), reached_t=torch.empty(B, dtype=torch.int32, **o),
                    crash=torch.empty(B, dtype=torch.uint8, **o), min_dist=torch.empty(B, dtype=torch.float64, **o),
                    start_s=torch.empty(B, dtype=torch.float64, **o))

    def _pin(self, name, shape, dtype):
        t = self._pinned.get(name)
        if t is None or t.shape != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, pin_memory=self._pin_host)
            self._pinned[name] = t
        return t

    def plan_host(self, ego, cars_x, cars_v, cars_a, n_cars, mode="fast", probe: Optional["MpcEngine"] = None, margin: float = 1.1):
        """Same through HOST buffers (numpy or CPU tensors in, numpy out): the end-to-end call a CPU rollout loop makes.
        Inputs are staged through pinned memory unless they already are pinned CPU tensors; results come back in (re-used) pinned arrays.
        probe: a make_probe() engine -> mpc_plan_host_probed (fast mode; same outputs)."""
        B, T = int(ego.shape[0]), self.num_t
        f64, i32 = torch.float64, torch.int32

        def stage(name, a, shape, dt):
            # a pinned CPU tensor of the right type is handed to the copy engine as it is; anything else goes through a pinned buffer
            if torch.is_tensor(a) and a.device.type == "cpu" and a.is_pinned() and a.dtype == dt and a.is_contiguous() and tuple(a.shape) == shape:
                return a
            p = self._pin(name, shape, dt)
            p.numpy()[...] = a.numpy() if torch.is_tensor(a) else a
            return p
        pe, px = stage("ego", ego, (B, 4), f64), stage("cx", cars_x, (B, self.nmax), f64)
        pv, pn = stage("cv", cars_v, (B, self.nmax), f64), stage("n", n_cars, (B,), i32)
        o = dict(idx=self._pin("idx", (B, T), i32), s_seq=self._pin("seq", (B, T), f64), cost=self._pin("cost", (B,), f64),
                 reached_t=self._pin("reached", (B,), i32), crash=self._pin("crash", (B,), torch.uint8),
                 min_dist=self._pin("mind", (B,), f64), start_s=self._pin("s0", (B,), f64))
        with self._device_ctx():
            if probe is not None:
                _lib.check(self.lib.mpc_plan_host_probed(self.h, probe.h, float(margin), B, _ptr(pe), _ptr(px), _ptr(pv), None, _ptr(pn),
                                                         _ptr(o["idx"]), _ptr(o["s_seq"]), _ptr(o["cost"]), _ptr(o["reached_t"]),
                                                         _ptr(o["crash"]), _ptr(o["min_dist"]), _ptr(o["start_s"]), self._stream()))
            else:
                _lib.check(self.lib.mpc_plan_host(self.h, B, _ptr(pe), _ptr(px), _ptr(pv), None, _ptr(pn), _mode(mode),
                                                  _ptr(o["idx"]), _ptr(o["s_seq"]), _ptr(o["cost"]), _ptr(o["reached_t"]),
                                                  _ptr(o["crash"]), _ptr(o["min_dist"]), _ptr(o["start_s"]), self._stream()))
        return {k: v.numpy() for k, v in o.items()}

    # ------------------------------------------------------------------------------------------
    def build_grid(self, ego, cars_x, cars_v, cars_a, n_cars, dist_dtype=torch.float64):
        """K1: dense S-T grids in the layout the reference's solver consumes.
        Returns obstacles u8[B,T,S], distances [B,T,S], start_s[B], delta_s[B], num_s[B] (S = num_s_stride >= num_s_max)."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        T, S = self.num_t, self.num_s_stride
        o = dict(device=self.device)
        obstacles = torch.empty((B, T, S), dtype=torch.uint8, **o)
        distances = torch.empty((B, T, S), dtype=dist_dtype, **o)
        s0 = torch.empty(B, dtype=torch.float64, **o)
        ds = torch.empty(B, dtype=torch.float64, **o)
        ns = torch.empty(B, dtype=torch.int32, **o)
        with self._device_ctx():
            _lib.check(self.lib.mpc_build_grid(self.h, B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a), _ptr(n_cars),
                                               _ptr(obstacles), _ptr(distances), int(dist_dtype == torch.float32),
                                               _ptr(s0), _ptr(ds), _ptr(ns), self._stream()))
        return dict(obstacles=obstacles, distances=distances, start_s=s0, delta_s=ds, num_s=ns)

    def solve_dense(self, obstacles, distances, start_s, delta_s, num_s, v0, a0, mode="exact"):
        """K2: the drop-in for st_cy.solve_s_t_path_fast on caller-supplied grids [B,T,S]."""
        B, T, S = obstacles.shape
        assert obstacles.dtype == torch.uint8 and obstacles.is_contiguous() and distances.is_contiguous()
        assert distances.shape == obstacles.shape and distances.dtype in (torch.float64, torch.float32)
        o = dict(device=self.device)
        out = dict(idx=torch.empty((B, T), dtype=torch.int32, **o), s_seq=torch.empty((B, T), dtype=torch.float64, **o),
                   cost=torch.empty(B, dtype=torch.float64, **o), reached_t=torch.empty(B, dtype=torch.int32, **o))
        with self._device_ctx():
            _lib.check(self.lib.mpc_solve_dense(self.h, B, T, S, _ptr(obstacles), _ptr(distances),
                                                int(distances.dtype == torch.float32), _ptr(start_s), _ptr(delta_s),
                                                _ptr(num_s), _ptr(v0), _ptr(a0), _mode(mode), _ptr(out["idx"]),
                                                _ptr(out["s_seq"]), _ptr(out["cost"]), _ptr(out["reached_t"]), self._stream()))
        return out

    def finer_fit(self, s_seq, reached_t, ego):
        """st.finer_fit for a batch of plans: returns (fine [B,Nf] f64, n_fine [B] i32, speed [B] f64, iterations [B] i32)."""
        B = s_seq.shape[0]
        nf = int(self.lib.mpc_finer_fit_max_points())
        o = dict(device=self.device)
        fine = torch.zeros((B, nf), dtype=torch.float64, **o)
        n_fine = torch.empty(B, dtype=torch.int32, **o)
        speed = torch.empty(B, dtype=torch.float64, **o)
        iters = torch.empty(B, dtype=torch.int32, **o)
        with self._device_ctx():
            _lib.check(self.lib.mpc_finer_fit(self.h, B, _ptr(s_seq), _ptr(reached_t), _ptr(ego), _ptr(fine), nf, _ptr(n_fine),
                                              _ptr(speed), _ptr(iters), self._stream()))
        return fine, n_fine, speed, iters

    # ------------------------------------------------------------------------------------------
    def predict_step_with_ego(self, ego, cars_x, cars_v, cars_a, n_cars, selected_speed, dt, min_crash_distance=5.0,
                              inplace=False):
        """K4: HighwayState.predict_step_with_ego for the whole batch."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        if inplace:
            eo, xo, vo, ao = ego, cars_x, cars_v, cars_a
        else:
            eo, xo, vo, ao = torch.empty_like(ego), torch.empty_like(cars_x), torch.empty_like(cars_v), torch.empty_like(cars_x)
        crashed = torch.empty(B, dtype=torch.uint8, device=self.device)
        with self._device_ctx():
            _lib.check(self.lib.mpc_predict_step_with_ego(self.h, B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a),
                                                          _ptr(n_cars), _ptr(selected_speed), float(dt), float(min_crash_distance),
                                                          _ptr(eo), _ptr(xo), _ptr(vo), _ptr(ao), _ptr(crashed), self._stream()))
        return eo, xo, vo, ao, crashed

    def krauss_step(self, ego, cars_x, cars_v, cars_a, n_cars, selected_speed, dt, min_crash_distance, accel, decel, tau, min_gap, max_speed):
        """One tick of the Krauss world (mpc_krauss_step), in place; returns the crash flags u8[B]."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        crashed = torch.empty(B, dtype=torch.uint8, device=self.device)
        with self._device_ctx():
            _lib.check(self.lib.mpc_krauss_step(self.h, B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a), _ptr(n_cars),
                                                _ptr(selected_speed), float(dt), float(min_crash_distance), float(accel), float(decel),
                                                float(tau), float(min_gap), float(max_speed), _ptr(crashed), self._stream()))
        return crashed

    def predict_step_without_ego(self, ego, cars_x, cars_v, cars_a, n_cars, dt, min_crash_distance=5.0):
        """HighwayState.predict_step_without_ego (prediction.py:22-44) for the whole batch: (ego, x, v, a, crashed)."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        eo, xo, vo, ao = torch.empty_like(ego), torch.empty_like(cars_x), torch.empty_like(cars_v), torch.empty_like(cars_x)
        crashed = torch.empty(B, dtype=torch.uint8, device=self.device)
        with self._device_ctx():
            _lib.check(self.lib.mpc_predict_step_without_ego(self.h, B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a),
                                                             _ptr(n_cars), float(dt), float(min_crash_distance),
                                                             _ptr(eo), _ptr(xo), _ptr(vo), _ptr(ao), _ptr(crashed), self._stream()))
        return eo, xo, vo, ao, crashed

    def state_vector(self, ego, cars_x, cars_v, cars_a, n_cars, out=None):
        """dqn.get_state_vector_from_base_state -> f32 [B, >=20] (column 20, if present, is left for the time feature)."""
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        if out is None:
            out = torch.zeros((B, 21), dtype=torch.float32, device=self.device)
        with self._device_ctx():
            _lib.check(self.lib.mpc_state_vector(self.h, B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a), _ptr(n_cars),
                                                 _ptr(out), out.shape[1], self._stream()))
        return out

    def rollout_step(self, state_args, jerk, dt, min_crash_distance, stop_x, step, alive, selected_speed, roll_s, roll_len,
                     crash_predicted):
        """K4: one masked, in-place step of the combined controller's policy rollout (dqn.py:129-141)."""
        ego, cars_x, cars_v, cars_a, n_cars = state_args
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        assert jerk.dtype == torch.float64 and jerk.is_contiguous() and jerk.numel() == B
        assert alive.dtype == torch.uint8 and crash_predicted.dtype == torch.uint8 and roll_len.dtype == torch.int32
        assert roll_s.dtype == torch.float64 and roll_s.is_contiguous() and roll_s.shape[0] == B
        with self._device_ctx():
            _lib.check(self.lib.mpc_rollout_step(self.h, B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a), _ptr(n_cars),
                                                 _ptr(jerk), float(dt), float(min_crash_distance), float(stop_x), int(step),
                                                 _ptr(alive), _ptr(selected_speed), _ptr(roll_s), roll_s.shape[1], _ptr(roll_len),
                                                 _ptr(crash_predicted), self._stream()))

    def env_step(self, ep, state_args, prev_acc, delay, ticks, jerk, u_spawn, fresh, reward, flags, projected_jerk):
        """mpc_env_step: one fused, in-place tick of merge_gym.MergeEnv.  fresh = (gap_u [B,nmax], first_u [B], speed_z [B] or
        None, delay_u [B]) or None (no auto-reset); flags u8 [4,B] = (done, crashed, arrived, timeout)."""
        ego, cars_x, cars_v, cars_a, n_cars = state_args
        B = self._check_state(ego, cars_x, cars_v, cars_a, n_cars)
        for t in (prev_acc, delay, jerk, reward, projected_jerk):
            assert t.dtype == torch.float64 and t.shape == (B,) and t.is_contiguous() and self._is_dev(t)
        assert ticks.dtype == torch.int32 and flags.dtype == torch.uint8 and flags.shape == (4, B) and flags.is_contiguous()
        g, f, z, d = fresh if fresh is not None else (None, None, None, None)
        with self._device_ctx():
            _lib.check(self.lib.mpc_env_step(self.h, C.byref(ep), B, _ptr(ego), _ptr(cars_x), _ptr(cars_v), _ptr(cars_a), _ptr(n_cars),
                                             _ptr(prev_acc), _ptr(delay), _ptr(ticks), _ptr(jerk), _ptr(u_spawn), _ptr(g), _ptr(f),
                                             _ptr(z), _ptr(d), _ptr(reward), _ptr(flags), _ptr(projected_jerk), self._stream()))

    def speed_from_jerk(self, ego, jerk):
        B = ego.shape[0]
        out = torch.empty(B, dtype=torch.float64, device=self.device)
        with self._device_ctx():
            _lib.check(self.lib.mpc_speed_from_jerk(self.h, B, _ptr(ego), _ptr(jerk), _ptr(out), self._stream()))
        return out


def states_to_device(S: dict, device) -> dict:
    """numpy state dict (synthetic.make_states) -> dict of device tensors."""
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in S.items()}
