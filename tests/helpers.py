"""Shared helpers for the parity tests: run the CPU oracle over a batch of synthetic states."""
import numpy as np


def oracle_state(O, S, b):
    n = int(S["n_cars"][b])
    e = S["ego"][b]
    return O.make_state((e[0], e[1]), e[2], e[3], S["cars_x"][b, :n], S["cars_v"][b, :n], S["cars_a"][b, :n])


def oracle_plan_batch(O, p, S, num_t, layered=True, nthreads=8):
    return O.plan_batch(p, S["ego"], S["cars_x"], S["cars_v"], S["cars_a"], S["n_cars"], num_t, layered=layered,
                        nthreads=nthreads)


def mpc_params_from_oracle(op):
    """oracle Params -> libmpcb200 MpcParams (same field names)."""
    from rl_mpc_lanemerging_b200 import _lib
    p = _lib.MpcParams()
    for name in _lib.PARAM_FIELDS:
        setattr(p, name, getattr(op, name))
    return p


def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
