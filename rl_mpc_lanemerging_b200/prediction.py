"""Traffic state records.

`HighwayState` keeps the reference's constructor and field names (reference prediction.py:9-20) so code
written against the reference keeps working; `BatchedState` is the structure-of-arrays form the CUDA
kernels consume (fp64, cars sorted front->back, padded to nmax).  The prediction *methods* run on the
GPU through libmpcb200 (K4, include/mpcb200.h): they are batch-of-one calls of the same kernel the
batched rollout uses.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

NMAX = 32


@dataclass
class BatchedState:
    ego: torch.Tensor       # [B,4] f64: x, y, speed, acceleration
    cars_x: torch.Tensor    # [B,nmax] f64
    cars_v: torch.Tensor
    cars_a: torch.Tensor
    n_cars: torch.Tensor    # [B] i32

    @property
    def batch(self) -> int:
        return int(self.ego.shape[0])

    def args(self):
        return self.ego, self.cars_x, self.cars_v, self.cars_a, self.n_cars

    def clone(self) -> "BatchedState":
        return BatchedState(*(t.clone() for t in self.args()))

    @classmethod
    def from_numpy(cls, S: dict, device) -> "BatchedState":
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        return cls(f(S["ego"]), f(S["cars_x"]), f(S["cars_v"]), f(S["cars_a"]), f(S["n_cars"]))

    @classmethod
    def from_states(cls, states: Sequence["HighwayState"], device, nmax: int = NMAX) -> "BatchedState":
        B = len(states)
        ego = np.zeros((B, 4)); cx = np.zeros((B, nmax)); cv = np.zeros((B, nmax)); ca = np.zeros((B, nmax))
        n = np.zeros(B, np.int32)
        for b, s in enumerate(states):
            k = len(s.other_xs)
            if k > nmax:
                raise ValueError(f"{k} cars exceed nmax={nmax}")
            ego[b] = (s.ego_position[0], s.ego_position[1], s.ego_speed, s.ego_acceleration)
            cx[b, :k] = s.other_xs; cv[b, :k] = s.other_speeds; ca[b, :k] = s.other_accelerations; n[b] = k
        return cls.from_numpy(dict(ego=ego, cars_x=cx, cars_v=cv, cars_a=ca, n_cars=n), device)

    def to_states(self) -> List["HighwayState"]:
        e, x, v, a, n = (t.cpu().numpy() for t in self.args())
        return [HighwayState((float(e[b, 0]), float(e[b, 1])), float(e[b, 2]), float(e[b, 3]), x[b, :n[b]].tolist(),
                             v[b, :n[b]].tolist(), a[b, :n[b]].tolist()) for b in range(len(n))]


class HighwayState:
    """One traffic snapshot: ego pose/speed/acceleration and the sensed cars, front -> back."""

    ego_reaction_threshold = 8
    ego_crash_threshold = 11

    def __init__(self, ego_position, ego_speed, ego_acceleration, other_xs, other_speeds, other_accelerations):
        self.ego_position = ego_position
        self.ego_speed = ego_speed
        self.ego_acceleration = ego_acceleration
        self.other_xs = other_xs
        self.other_speeds = other_speeds
        self.other_accelerations = other_accelerations

    @classmethod
    def empty_state(cls):
        return cls(0, 0, 0, [], [], [])

    def predict_step_with_ego(self, selected_speed, delta_t, min_crash_distance=5) -> Tuple["HighwayState", bool]:
        """Reference prediction.py:46-105, evaluated by the K4 kernel (batch of one)."""
        from . import st
        eng = st.get_engine()
        bs = BatchedState.from_states([self], eng.device, eng.nmax)
        sel = torch.tensor([float(selected_speed)], dtype=torch.float64, device=eng.device)
        eo, xo, vo, ao, cr = eng.predict_step_with_ego(*bs.args(), sel, delta_t, min_crash_distance)
        out = BatchedState(eo, xo, vo, ao, bs.n_cars).to_states()[0]
        return out, bool(cr.item())

    def predict_step_without_ego(self, delta_t, min_crash_distance=5) -> Tuple["HighwayState", bool]:
        """Reference prediction.py:22-44 (the traffic-only step of the S-T grid builder), batch of one."""
        from . import st
        eng = st.get_engine()
        bs = BatchedState.from_states([self], eng.device, eng.nmax)
        eo, xo, vo, ao, cr = eng.predict_step_without_ego(*bs.args(), delta_t, min_crash_distance)
        out = BatchedState(eo, xo, vo, ao, bs.n_cars).to_states()[0]
        return out, bool(cr.item())

    def get_closest_cars(self) -> Tuple[Optional[tuple], Optional[tuple]]:
        """(car directly ahead, car directly behind) as (x, speed, acceleration) tuples or None (host bookkeeping)."""
        behind = next((i for i, x in enumerate(self.other_xs) if x < self.ego_position[0]), None)
        front = (behind - 1) if behind is not None else len(self.other_xs) - 1
        pick = lambda i: (self.other_xs[i], self.other_speeds[i], self.other_accelerations[i])  # noqa: E731
        return (pick(front) if front >= 0 else None), (pick(behind) if behind is not None else None)

    def __repr__(self):
        return (f"HighwayState(ego={self.ego_position}, v={self.ego_speed}, a={self.ego_acceleration}, "
                f"cars={len(self.other_xs)})")


def tdiv(t, scalar):
    """IEEE division of a tensor by a Python scalar.  `tensor / scalar` on a CUDA device multiplies by the rounded reciprocal
    (x / 0.2 becomes x * 5.0), which differs from the reference's numpy / float division in the last bit; dividing by a 0-dim
    tensor is a true division on every device.  (The divisors are cached per device: creating one copies from the host, which a
    CUDA-graph capture does not allow.)"""
    key = (float(scalar), t.dtype, t.device)
    d = _TDIV_CACHE.get(key)
    if d is None:
        import torch
        if t.is_cuda and torch.cuda.is_current_stream_capturing():
            d = torch.full((), float(scalar), dtype=t.dtype, device=t.device)          # (a fill kernel: capturable; not cached)
            return t / d
        d = _TDIV_CACHE[key] = torch.as_tensor(float(scalar), dtype=t.dtype, device=t.device)
    return t / d


_TDIV_CACHE: dict = {}

