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


def fast_mode_matches_model(O, info, op, st, idx, cost):
    """Is (idx, cost) of a fast-mode plan bit-identical to the CPU model of the kernel that produced it?  `info` =
    MpcEngine.fast32_info().  The 32-bit-key kernel makes the first attempt (2^-frac_bits labels under its bound); what it
    hands on -- bounded attempt failed, or the frontier outgrew its ring -- is solved by the 64-bit kernel (2^-18 labels):
    the model ladder is the normal case, the plain 2^-18 model the answer for a problem that was handed on for the ring."""
    ob, di, sv = O.build_grid(op, st)
    a = (op, ob, di, sv, op.t_disc, st.ego_v, st.ego_a)
    cands = [O.solve_fast_ladder(*a, info["frac_bits"] if info["in_use"] else 0, info["bound_fx"])]
    if info["in_use"]:
        cands.append(O.solve_fast_model(*a))
    return any(np.array_equal(m["idx"], idx) and m["cost"] == cost for m in cands), cands[0]
