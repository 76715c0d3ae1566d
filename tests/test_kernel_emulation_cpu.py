"""The library's kernel SOURCES executed on the CPU (tests/emu: g++ + a fiber shim for the CUDA execution model) and
compared with the CPU oracle / the C model of the fast kernel.

Why: device code written when no GPU time is left can still be RUN.  The emulation executes the very text nvcc compiles
(predict_layers_kernel, reach_caps_kernel, fast_pull_kernel incl. its HINT instances, predict_step_without_ego_kernel): control
flow, index arithmetic, shared-memory layout, the retry ladder, the pruning rules.  It cannot show anything about the
hardware (inline PTX, memory model, races -- a CAS never loses here -- occupancy, timing): the GPU tests remain the gate.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import bound_model as BM
from oracle import cpu_oracle as O
from rl_mpc_lanemerging_b200 import synthetic
from rl_mpc_lanemerging_b200._lib import PARAM_FIELDS, MpcParams
from tests.emu import build_emu

FX = 262144.0


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(build_emu.build())
    lib.emu_last_error.restype = C.c_char_p
    lib.emu_node_count.restype = C.c_longlong
    return lib


def _params(op):
    p = MpcParams()
    for n in PARAM_FIELDS:
        setattr(p, n, getattr(op, n))
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _plan(emu, op, S, threads, ring=0, use_bound=1, hint=None, scale=1.0, retry=1.36, caps=1):
    B = S["ego"].shape[0]
    T = O.grid_dims(op, 0.0)[0]
    out = dict(idx=np.zeros((B, T), np.int32), s_seq=np.zeros((B, T)), cost=np.zeros(B), reached_t=np.zeros(B, np.int32),
               crash=np.zeros(B, np.uint8), min_dist=np.zeros(B), start_s=np.zeros(B), fallback=np.zeros(B, np.uint8))
    nt = C.c_int()
    p = _params(op)
    rc = emu.emu_plan(C.byref(p), B, S["cars_x"].shape[1], _ptr(S["ego"]), _ptr(S["cars_x"]), _ptr(S["cars_v"]), _ptr(S["n_cars"]),
                      threads, ring, use_bound, _ptr(hint), C.c_double(scale), C.c_double(retry), caps,
                      _ptr(out["idx"]), _ptr(out["s_seq"]), _ptr(out["cost"]), _ptr(out["reached_t"]), _ptr(out["crash"]),
                      _ptr(out["min_dist"]), _ptr(out["start_s"]), _ptr(out["fallback"]), C.byref(nt))
    assert rc == 0, emu.emu_last_error()
    assert nt.value == T
    out["nodes"] = int(emu.emu_node_count())
    return out


def _state(S, b):
    e, n = S["ego"][b], S["n_cars"][b]
    return O.make_state((e[0], e[1]), e[2], e[3], S["cars_x"][b, :n], S["cars_v"][b, :n], S["cars_a"][b, :n])


def _bound_fx(op):
    return int(np.rint(BM.zone_bound(op) * FX)) - 2            # DevParams::bound_fx (mpc_derive.h)


def _model_ladder(op, st, hint_fx, mid_fx, use_h=True):
    """The kernel's attempts on the C model, in label units: returns (answer, nodes over all attempts)."""
    ob, di, sv = O.build_grid(op, st)
    H = ob.shape[0] - 1
    bound = _bound_fx(op)
    a = (op, ob, di, sv, op.t_disc, st.ego_v, st.ego_a)
    nodes = 0
    if hint_fx is not None and hint_fx >= bound:              # zone-crossing plan expected: general loop under the hint, then unbounded
        r = O.solve_fast_model_ex(*a, prune_fx=hint_fx)
        nodes += r["nodes"]
        if r["reached_t"] < H:
            r = O.solve_fast_model_ex(*a, prune_fx=0)
            nodes += r["nodes"]
        return r, nodes
    h = BM.heuristic_table(op, ob, di) if (use_h and hint_fx is not None) else None
    rungs = ([hint_fx] if hint_fx is not None else []) + ([mid_fx] if hint_fx is not None and mid_fx is not None and hint_fx < mid_fx < bound else []) + [bound]
    for U in rungs:
        r = (BM.solve_with_heuristic(op, ob, di, sv, st.ego_v, st.ego_a, 0.0, h, U_fx=U) if h is not None
             else O.solve_fast_model_ex(*a, prune_fx=U))
        nodes += r["nodes"]
        if r["reached_t"] == H:
            return r, nodes
    r = O.solve_fast_model_ex(*a, prune_fx=0)
    return r, nodes + r["nodes"]


@pytest.mark.parametrize("H,traffic,kind,B,threads,ring", [(17, "moderate", "mixed", 12, 192, 0), (17, "fast", "mixed", 6, 128, 1400),
                                                            (50, "moderate", "onramp", 3, 512, 6568), (25, "low", "mixed", 4, 384, 0)])
def test_emulated_fused_plan_matches_the_model_and_the_oracle(emu, H, traffic, kind, B, threads, ring):
    """Un-hinted path (the one measured on the GPU): the emulated kernels must reproduce the C model bit for bit -- which is
    what the GPU tests assert of the real device -- so this also validates the emulation itself."""
    op = O.horizon_params(H)
    S = synthetic.make_states(B, traffic, seed=3, kind=kind)
    got = _plan(emu, op, S, threads, ring)
    nodes = 0
    for b in range(B):
        st = _state(S, b)
        ref, n = _model_ladder(op, st, None, None)
        nodes += n
        if got["fallback"][b]:
            continue
        assert got["reached_t"][b] == ref["reached_t"] and np.array_equal(got["idx"][b], ref["idx"]) and got["cost"][b] == ref["cost"], b
        orc = O.plan(op, st, layered=True)                      # crash verdict / minimum distance / start_s: the reference's own
        assert got["start_s"][b] == orc["start_s"]
        if np.array_equal(ref["idx"], orc["idx"]):
            assert bool(got["crash"][b]) == orc["crash"] and got["min_dist"][b] == orc["min_path_distance"], b
    if not got["fallback"].any():
        assert got["nodes"] == nodes                            # same nodes finalised, attempt by attempt


@pytest.mark.parametrize("H,traffic,kind,B,threads,ring", [(17, "moderate", "mixed", 10, 192, 0), (50, "moderate", "onramp", 3, 512, 6568),
                                                            (50, "fast", "onramp", 2, 384, 3800)])
def test_emulated_hinted_plan_is_identical_and_expands_what_the_model_says(emu, H, traffic, kind, B, threads, ring):
    op = O.horizon_params(H)
    S = synthetic.make_states(B, traffic, seed=4, kind=kind)
    plain = _plan(emu, op, S, threads, 0 if ring == 3800 else ring)
    assert not plain["fallback"].any()
    for scale in (1.1, 1.0, 0.93, 0.5, 3.0, 400.0):
        hint = plain["cost"].copy()
        got = _plan(emu, op, S, threads, ring, hint=hint, scale=scale)
        if ring == 3800:
            # the three-blocks-per-SM ring of H=50: wide enough for the pruned frontier of a good hint; a retry
            # under the standard bound overflows it and hands the problem to the full-row re-solve (not emulated)
            assert not got["fallback"].any() or scale not in (1.1, 1.0)
            if got["fallback"].any():
                continue
        assert not got["fallback"].any()
        for k in ("idx", "s_seq", "cost", "reached_t", "crash", "min_dist", "start_s"):
            assert np.array_equal(got[k], plain[k]), (scale, k)
        nodes = 0
        for b in range(B):
            hc = hint[b] * scale
            ok = 0.0 < hc < 1e9
            hint_fx = int(np.rint(hc * FX)) if ok else None
            mid_fx = int(np.rint(hc * 1.36 * FX)) if ok else None
            nodes += _model_ladder(op, _state(S, b), hint_fx, mid_fx)[1]
        assert got["nodes"] == nodes, (scale, got["nodes"], nodes)
        if scale == 1.1 and H == 50:
            assert got["nodes"] < 0.8 * plain["nodes"]
    if ring == 3800:
        return
    # without the reachability caps (MPC_FAST_HEUR=0) and without the middle rung: still identical
    got = _plan(emu, op, S, threads, ring, hint=plain["cost"].copy(), scale=0.9, retry=0.0, caps=0)
    for k in ("idx", "cost", "reached_t", "crash"):
        assert np.array_equal(got[k], plain[k]), k
    junk = plain["cost"].copy()
    junk[0::3] = np.nan; junk[1::3] = -5.0
    got = _plan(emu, op, S, threads, ring, hint=junk, scale=1.0)
    for k in ("idx", "cost", "reached_t", "crash"):
        assert np.array_equal(got[k], plain[k]), k


@pytest.mark.parametrize("seed", [1, 2])
def test_result_does_not_depend_on_thread_interleaving(emu, seed):
    """With random preemption at every shared-memory access the threads of a block interleave inside a phase: CAS operations
    lose, the retry / redo paths of the min-combine run, cells are finalised while neighbours still push.  The answer and the
    number of finalised nodes must not change (integer labels: the min-combine is order independent)."""
    op = O.horizon_params(17)
    S = synthetic.make_states(6, "moderate", seed=30 + seed, kind="mixed")
    emu.emu_set_preempt(C.c_ulonglong(0))
    plain = _plan(emu, op, S, 192, 1600)
    hinted = _plan(emu, op, S, 192, 1600, hint=plain["cost"].copy(), scale=0.9)
    try:
        emu.emu_set_preempt(C.c_ulonglong(0x9E3779B97F4A7C15 * seed % (1 << 64)))
        lost0 = emu.emu_cas_lost()
        for ref, kw in ((plain, {}), (hinted, dict(hint=plain["cost"].copy(), scale=0.9))):
            got = _plan(emu, op, S, 192, 1600, **kw)
            for k in ("idx", "s_seq", "cost", "reached_t", "crash", "min_dist", "fallback"):
                assert np.array_equal(got[k], ref[k]), k
            assert got["nodes"] == ref["nodes"]
        assert emu.emu_cas_lost() > lost0                     # races did happen
    finally:
        emu.emu_set_preempt(C.c_ulonglong(0))


def test_emulated_probe_grid_runs_the_same_kernels(emu):
    """The 20x3 coarse probe grid of mpc_plan_probed is just another Settings snapshot for the same kernels."""
    op = O.horizon_params(50)
    q = BM.probe_params(op, 20, 3)
    S = synthetic.make_states(8, "moderate", seed=0)
    got = _plan(emu, q, S, 192)
    assert got["idx"].shape[1] == 18
    for b in range(8):
        st = _state(S, b)
        ob, di, sv = O.build_grid(q, st)
        ref = O.solve_fast_model(q, ob, di, sv, q.t_disc, st.ego_v, st.ego_a)
        assert got["reached_t"][b] == ref["reached_t"] and np.array_equal(got["idx"][b], ref["idx"]) and got["cost"][b] == ref["cost"]


@pytest.mark.parametrize("traffic,kind,dt,mcd", [("moderate", "mixed", 0.3, 5.0), ("default", "mixed", 0.2, 5.1)])
def test_emulated_predict_step_without_ego_matches_oracle(emu, traffic, kind, dt, mcd):
    B, nmax = 96, 32
    op = O.default_params()
    S = synthetic.make_states(B, traffic, seed=21, kind=kind)
    S["n_cars"][:3] = 0
    S["ego"][5] = (60.0, -1.6, 12.0, 0.0)                      # the ego leads every car (prediction.py:28-31)
    S["cars_x"][5, :3] = (40.0, 20.0, 0.0); S["cars_v"][5, :3] = 11.0; S["cars_a"][5, :3] = 0.0; S["n_cars"][5] = 3
    p = _params(op)
    eo, xo, vo, ao, cr = np.zeros((B, 4)), np.zeros((B, nmax)), np.zeros((B, nmax)), np.zeros((B, nmax)), np.zeros(B, np.uint8)
    rc = emu.emu_predict_step_without_ego(C.byref(p), B, nmax, _ptr(S["ego"]), _ptr(S["cars_x"]), _ptr(S["cars_v"]), _ptr(S["cars_a"]),
                                          _ptr(S["n_cars"]), C.c_double(dt), C.c_double(mcd), _ptr(eo), _ptr(xo), _ptr(vo), _ptr(ao), _ptr(cr))
    assert rc == 0
    branches = set()
    for b in range(B):
        st = _state(S, b)
        out, crashed = O.predict_step_without_ego(op, st, dt, mcd)
        n = st.n
        assert tuple(eo[b]) == (out.ego_x, out.ego_y, out.ego_v, out.ego_a), b
        assert np.array_equal(xo[b, :n], np.array(out.x[:n])) and np.array_equal(vo[b, :n], np.array(out.v[:n])), b
        assert np.array_equal(ao[b, :n], np.array(out.a[:n])) and bool(cr[b]) == crashed, b
        es = O.get_ego_s(st.ego_x, st.ego_y)
        branches.add("stay" if (es < 8 or n == 0) else ("lead" if st.x[0] < st.ego_x else "follow"))
    assert branches == {"stay", "lead", "follow"}


@pytest.mark.parametrize("over", [dict(desired_speed=20.0), dict(start_uncertainty=1.0, uncertainty_per_second=0.5),
                                  dict(max_speed=27.3), dict(min_allowed_distance=3.0, v_weight=0.1, d_weight=50.0)])
def test_hinted_kernel_over_other_settings(emu, over):
    """Settings away from the published ones (cheapest speed below MAX_SPEED, growing uncertainty, a non-integer speed clamp,
    other weights / safety distance): the un-hinted kernel still equals the C model, hints still change nothing."""
    op = O.horizon_params(17, **over)
    S = synthetic.make_states(6, "moderate", seed=41, kind="mixed")
    plain = _plan(emu, op, S, 192)
    for b in range(6):
        st = _state(S, b)
        ob, di, sv = O.build_grid(op, st)
        ref = O.solve_fast_model(op, ob, di, sv, op.t_disc, st.ego_v, st.ego_a)
        assert plain["reached_t"][b] == ref["reached_t"] and np.array_equal(plain["idx"][b], ref["idx"]) and plain["cost"][b] == ref["cost"]
    for scale in (1.1, 0.8, 50.0):
        got = _plan(emu, op, S, 192, hint=plain["cost"].copy(), scale=scale)
        for k in ("idx", "cost", "reached_t", "crash", "min_dist"):
            assert np.array_equal(got[k], plain[k]), (scale, k)
        if scale == 1.1:
            assert got["nodes"] < plain["nodes"]
