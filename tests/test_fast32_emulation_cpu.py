"""fast32_kernel (csrc/mpc_fast32.cuh: 32-bit keys, native shared-memory min) executed on the CPU emulation of tests/emu and
compared bit for bit with its C model (oracle/mpc_oracle.c: orc_solve_fast_model_q) -- and through it with the oracle.

The chain is the one mpc_plan runs in fast mode (run_solve, mpc_api.cu): fast32_kernel on every state, fast_pull_kernel (64-bit
words) on what it hands on.  Cases cover the full-row and the wrapped ring, ring overflow, plans that cross a penalty zone /
do not reach the horizon (handed on with the "bounded attempt failed" flag), the same-bucket candidate list (about five
entries per H=50 problem) and random preemption at every shared-memory access (offers interleave across threads)."""
import ctypes as C

import numpy as np
import pytest

from oracle import cpu_oracle as O
from rl_mpc_lanemerging_b200 import synthetic
from rl_mpc_lanemerging_b200._lib import PARAM_FIELDS, MpcParams
from tests import helpers
from tests.emu import build_emu


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(build_emu.build())
    lib.emu_last_error.restype = C.c_char_p
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _plan32(emu, op, S, threads32, ring32, preempt=0, threads=256, ring=0):
    B, T = S["ego"].shape[0], O.grid_dims(op, 0.0)[0]
    p = MpcParams()
    for n in PARAM_FIELDS:
        setattr(p, n, getattr(op, n))
    out = dict(idx=np.zeros((B, T), np.int32), s_seq=np.zeros((B, T)), cost=np.zeros(B), reached_t=np.zeros(B, np.int32),
               crash=np.zeros(B, np.uint8), min_dist=np.zeros(B), handed=np.zeros(B, np.uint8))
    info = (C.c_int64 * 3)()
    emu.emu_set_preempt(C.c_ulonglong(preempt))
    try:
        rc = emu.emu_plan32(C.byref(p), B, S["cars_x"].shape[1], _ptr(S["ego"]), _ptr(S["cars_x"]), _ptr(S["cars_v"]), _ptr(S["n_cars"]),
                            threads32, ring32, threads, ring, _ptr(out["idx"]), _ptr(out["s_seq"]), _ptr(out["cost"]),
                            _ptr(out["reached_t"]), _ptr(out["crash"]), _ptr(out["min_dist"]), _ptr(out["handed"]), info)
    finally:
        emu.emu_set_preempt(C.c_ulonglong(0))
    assert rc == 0, emu.emu_last_error()
    out["frac_bits"], out["bound_fx"] = int(info[0]), int(info[1])
    return out


@pytest.mark.parametrize("H,traffic,kind,B,threads32,ring32,preempt", [
    (17, "moderate", "onramp", 24, 128, 0, 0),
    (17, "moderate", "mixed", 32, 192, 2048, 11),       # wrapped ring, zone-crossing / infeasible plans, preemption
    (17, "fast", "mixed", 16, 512, 1536, 0),            # ring overflows are handed on
    (25, "low", "onramp", 10, 256, 0, 5),
    (50, "moderate", "onramp", 8, 512, 5304, 7),        # 2^-17 labels; the same-bucket candidate list is exercised
])
def test_fast32_chain_equals_its_model_and_the_oracle(emu, H, traffic, kind, B, threads32, ring32, preempt):
    op = O.horizon_params(H)
    S = synthetic.make_states(B, traffic, seed=71, kind=kind)
    out = _plan32(emu, op, S, threads32, ring32, preempt)
    assert out["frac_bits"] == {17: 18, 25: 18, 50: 17}[H]
    ref = helpers.oracle_plan_batch(O, op, S, H + 1)
    assert np.array_equal(out["reached_t"], ref["reached_t"])
    n32 = 0
    for b in range(B):
        st = helpers.oracle_state(O, S, b)
        ob, di, sv = O.build_grid(op, st)
        a = (op, ob, di, sv, op.t_disc, st.ego_v, st.ego_a)
        m = O.solve_fast_ladder(*a, out["frac_bits"], out["bound_fx"])
        h = int(out["handed"][b])
        assert not h & 4                                           # nothing needed the exact kernel
        if h == 1:                                                 # handed on for the ring: the 64-bit kernel's own ladder
            m = O.solve_fast_model(*a)
        else:
            assert (m["stage"] == 32) == (h == 0), (b, h, m["stage"])
        n32 += h == 0
        assert np.array_equal(m["idx"], out["idx"][b]) and m["cost"] == out["cost"][b], (b, h, m["cost"], out["cost"][b])
        if ref["cost"][b] > 0:
            assert abs(out["cost"][b] - ref["cost"][b]) <= 1e-6 * ref["cost"][b]
        assert np.array_equal(out["idx"][b], ref["idx"][b])        # and the oracle's sequence
    assert n32 > 0
    same = (out["idx"] == ref["idx"]).all(1)
    assert np.array_equal(out["crash"].astype(bool)[same], ref["crash"][same])


def test_same_bucket_candidates_are_resolved_on_exact_labels():
    """Labels that agree above their low 8 bits are ordered by v' alone in the 32-bit key: the model with that rule differs from
    the exact comparison on some states (which is why the kernel keeps the candidate list), and the exact rule at the kernel's
    precision reproduces the oracle's sequences."""
    op = O.horizon_params(50)
    S = synthetic.make_states(48, "moderate", seed=11)
    differ = 0
    for b in range(48):
        st = helpers.oracle_state(O, S, b)
        ob, di, sv = O.build_grid(op, st)
        a = (op, ob, di, sv, op.t_disc, st.ego_v, st.ego_a)
        exact = O.solve_fast_model_q(*a, 17, 0)
        coarse = O.solve_fast_model_q(*a, 17, 10)                  # comparisons ignore the low 10 bits
        ref = O.solve(*a, layered=True)
        assert np.array_equal(exact["idx"], ref["idx"])
        differ += not np.array_equal(coarse["idx"], exact["idx"])
    assert differ <= 2
