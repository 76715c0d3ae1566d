"""The WHOLE library -- C ABI, handle set-up, launch configuration, the fallback chain, every kernel -- compiled by g++ against
the CUDA shim of tests/emu and driven with numpy arrays in the place of device pointers (tests/emu/emu_api.py).

Covers on the CPU what otherwise only a device run can: the host-side orchestration of the entry points written without GPU
access (mpc_plan_hinted, mpc_plan_probed, mpc_plan_host_probed, mpc_predict_step_without_ego), the ring-overflow re-solve
of hinted problems, the fp32 dense solve.  See tests/emu/cuda_emu.h for what an emulation can and cannot show.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import cpu_oracle as O
from rl_mpc_lanemerging_b200 import _lib, synthetic
from tests import helpers
from tests.emu import emu_api as EA

KEYS = ("idx", "s_seq", "cost", "reached_t", "crash", "min_dist", "start_s")


def _params(H, **over):
    op = O.horizon_params(H, **over)
    return op, helpers.mpc_params_from_oracle(op)


def _same(a, b, keys=KEYS):
    for k in keys:
        assert np.array_equal(a[k], b[k]), k


def test_plan_exact_and_fast_against_the_oracle():
    op, p = _params(17)
    eng = EA.EmuEngine(p, max_batch=16)
    S = synthetic.make_states(12, "moderate", seed=5, kind="mixed")
    ref = helpers.oracle_plan_batch(O, op, S, 18)
    ex = eng.plan(S, mode=_lib.MODE_EXACT)
    assert np.array_equal(ex["idx"], ref["idx"]) and np.array_equal(ex["cost"], ref["cost"]) and np.array_equal(ex["s_seq"], ref["s_seq"])
    assert np.array_equal(ex["crash"].astype(bool), ref["crash"]) and np.array_equal(ex["min_dist"], ref["min_path_distance"])
    fa = eng.plan(S)
    assert np.array_equal(fa["reached_t"], ref["reached_t"]) and (fa["idx"] == ref["idx"]).all(1).sum() >= 11
    ok = ref["cost"] > 0
    assert np.all(helpers.rel(fa["cost"][ok], ref["cost"][ok]) < 1e-6)
    eng.close()


@pytest.mark.parametrize("H,B,mult", [(50, 4, (20, 3)), (17, 8, (8, 2))])
def test_probed_and_hinted_plans_equal_the_plain_plan(H, B, mult):
    _op, p = _params(H)
    # hinted / probed solves run on the 64-bit kernel: the plain plan they must equal bit for bit is that kernel's too
    # (at H=50 the 32-bit-key kernel of the default path carries 2^-17 labels: same sequences, cost equal to ~1e-8)
    eng = EA.EmuEngine(p, max_batch=B, env={"MPC_FAST32": "0"})
    q = _lib.MpcParams()
    for n in _lib.PARAM_FIELDS:
        setattr(q, n, getattr(p, n))
    q.s_disc, q.t_disc = p.s_disc * mult[0], p.t_disc * mult[1]
    probe = EA.EmuEngine(q, max_batch=B)
    S = synthetic.make_states(B, "moderate", seed=8, kind="mixed" if H == 17 else "onramp")
    ref = eng.plan(S)
    dflt = EA.EmuEngine(p, max_batch=B)                        # the default path: 32-bit-key kernel first
    got = dflt.plan(S)
    _same(got, ref, keys=tuple(k for k in KEYS if k != "cost"))
    assert np.all(np.abs(got["cost"] - ref["cost"]) <= 1e-7 * np.maximum(ref["cost"], 1.0))
    dflt.close()
    _same(eng.plan_probed(probe, S, margin=1.1), ref)
    assert eng.counters()["kernels_launched"] >= 7            # probe: predictor + DP (+ re-solves); plan: predictor + caps + DP (+ re-solves)
    _same(eng.plan_probed(probe, S, margin=0.4), ref)         # every first bound too low
    _same(eng.plan_probed(probe, S, margin=1.1, host=True), ref)
    for scale in (1.05, 0.7, 500.0):
        _same(eng.plan_hinted(S, ref["cost"].copy(), hint_scale=scale), ref)
    junk = ref["cost"].copy()
    junk[::3] = np.nan; junk[1::3] = -2.0
    _same(eng.plan_hinted(S, junk), ref)
    reached = np.where(np.arange(B) % 2 == 0, H, H - 1).astype(np.int32)
    _same(eng.plan_hinted(S, ref["cost"] * 0.5, hint_reached=reached, hint_full_t=H), ref)
    ex = eng.plan(S, mode=_lib.MODE_EXACT)
    _same(eng.plan_hinted(S, junk, mode=_lib.MODE_EXACT), ex)
    with pytest.raises(RuntimeError):                          # the hint must not alias the cost output
        o = eng._out(B)
        EA.check(eng.L.mpc_plan_hinted(eng.h, B, EA._p(S["ego"]), EA._p(S["cars_x"]), EA._p(S["cars_v"]), None, EA._p(S["n_cars"]), 0,
                                       EA._p(o["cost"]), None, 0, 1.0, EA._p(o["idx"]), EA._p(o["s_seq"]), EA._p(o["cost"]),
                                       EA._p(o["reached_t"]), None, None, None, None))
    probe.close(); eng.close()


def test_hinted_problems_that_overflow_the_ring_are_resolved():
    """Five blocks per SM at H=50 leave a ring of ~2.1 k cells: many frontiers outgrow it and go through the library's re-solve
    chain (full-row fast kernel -- HINT instance for hinted calls -- then the exact kernel): same answers."""
    _op, p = _params(50)
    S = synthetic.make_states(5, "fast", seed=2)
    wide = EA.EmuEngine(p, max_batch=8, env={"MPC_FAST32": "0"})
    ref = wide.plan(S)
    assert wide.counters()["fallback_problems"] == 0
    small = EA.EmuEngine(p, max_batch=8, env={"MPC_FAST_BLOCKS": "160", "MPC_FAST32": "0"})
    got = small.plan(S)
    assert small.counters()["fallback_problems"] > 0
    _same(got, ref)
    got = small.plan_hinted(S, ref["cost"].copy(), hint_scale=0.6)      # low hints: retries under wider bounds overflow the ring
    assert small.counters()["fallback_problems"] > 0
    _same(got, ref)
    _same(small.plan_hinted(S, ref["cost"].copy(), hint_scale=1.1), ref)
    # the same through the 32-bit-key kernel with a ring of ~4.8 k cells: what outgrows it goes to its wide-ring launch shape
    small32 = EA.EmuEngine(p, max_batch=8, env={"MPC_F32_BLOCKS": "96"})
    got = small32.plan(S)
    info = small32.fast32_info()
    assert info["in_use"] and info["ring_cells"] < 5000 and info["first_shape_handed_on"] > 0 and info["handed_on"] == 0
    _same(got, ref, keys=tuple(k for k in KEYS if k != "cost"))
    assert np.all(np.abs(got["cost"] - ref["cost"]) <= 1e-7 * np.maximum(ref["cost"], 1.0))
    wide.close(); small.close(); small32.close()


def test_dense_solve_on_fp32_and_fp64_grids():
    """K2 (mpc_solve_dense) fed by mpc_build_grid: fp64 grids give the fused plan exactly, fp32 distances within tolerance."""
    op, p = _params(17)
    eng = EA.EmuEngine(p, max_batch=8)
    S = synthetic.make_states(6, "moderate", seed=6, kind="mixed")
    ref = eng.plan(S)
    v0, a0 = S["ego"][:, 2].copy(), S["ego"][:, 3].copy()
    g64 = eng.build_grid(S)
    for b in range(6):                                         # the grids themselves: bit-identical to the oracle's
        ob, di, sv = O.build_grid(op, helpers.oracle_state(O, S, b))
        n = sv.size
        assert g64["num_s"][b] == n and np.array_equal(g64["obstacles"][b, :, :n], ob) and np.array_equal(g64["distances"][b, :, :n], di)
    d64 = eng.solve_dense(g64, v0, a0)
    _same(d64, ref, ("idx", "cost", "reached_t"))
    d32 = eng.solve_dense(eng.build_grid(S, f32=True), v0, a0)
    assert np.array_equal(d32["reached_t"], ref["reached_t"])
    ok = ref["cost"] > 0
    assert np.all(helpers.rel(d32["cost"][ok], ref["cost"][ok]) < 1e-4)
    eng.close()


def test_predict_step_without_ego_entry_point():
    op, p = _params(17)
    eng = EA.EmuEngine(p, max_batch=64)
    S = synthetic.make_states(48, "default", seed=12, kind="mixed")
    eo, xo, vo, ao, cr = eng.predict_step_without_ego(S, 0.3, 5.0)
    for b in range(48):
        st = helpers.oracle_state(O, S, b)
        out, crashed = O.predict_step_without_ego(op, st, 0.3, 5.0)
        n = st.n
        assert tuple(eo[b]) == (out.ego_x, out.ego_y, out.ego_v, out.ego_a) and bool(cr[b]) == crashed, b
        assert np.array_equal(xo[b, :n], np.array(out.x[:n])) and np.array_equal(vo[b, :n], np.array(out.v[:n])), b
        assert np.array_equal(ao[b, :n], np.array(out.a[:n])), b
    eng.close()


def test_remaining_entry_points_against_the_oracle():
    """The rest of the C ABI on the emulated library (all of it verified on a device already): guards the host-side
    refactors made without GPU access (shared plan / plan_host implementations, derive_params header, launch macro)."""
    op, p = _params(17)
    eng = EA.EmuEngine(p, max_batch=32)
    S = synthetic.make_states(24, "moderate", seed=15, kind="mixed")
    assert eng.selftest_search(S) == 0
    _same(eng.plan_host(S), eng.plan(S))
    rng = np.random.default_rng(0)
    sel, jerk = rng.uniform(0, 30, 24), rng.uniform(-6, 6, 24)
    eo, xo, vo, ao, cr = eng.predict_step_with_ego(S, sel, 0.2, 5.1)
    sv = eng.state_vector(S)
    sp = eng.speed_from_jerk(S, jerk)
    for b in range(24):
        st = helpers.oracle_state(O, S, b)
        out, crashed = O.predict_step_with_ego(op, st, sel[b], 0.2, 5.1)
        n = st.n
        assert tuple(eo[b]) == (out.ego_x, out.ego_y, out.ego_v, out.ego_a) and bool(cr[b]) == crashed
        assert np.array_equal(xo[b, :n], np.array(out.x[:n])) and np.array_equal(vo[b, :n], np.array(out.v[:n]))
        assert np.array_equal(ao[b, :n], np.array(out.a[:n]))
        assert np.allclose(sv[b, :20], O.state_vector(op, st).astype(np.float32), rtol=0, atol=1e-7)
        assert sp[b] == O.speed_from_jerk(op, st.ego_v, st.ego_a, jerk[b])
    plan = eng.plan(S)
    fine, n_fine, speed, iters = eng.finer_fit(plan["s_seq"], plan["reached_t"], S["ego"])
    tick = op.tick_length
    for b in range(24):                                        # the projected plan obeys the limits of st.py:608-668
        n = n_fine[b]
        if n < 4:
            continue
        x = fine[b, :n]
        v = np.diff(x) / tick
        assert v.min() > -1e-6 and v.max() < op.max_speed + 1e-6
        a = np.diff(np.concatenate(([S["ego"][b, 2]], v))) / tick
        assert a.min() > op.a_min - 1e-5 and a.max() < op.a_max + 1e-5
        assert abs(speed[b] - (x[1] - x[0]) / tick) < 1e-12
    eng.close()


def test_finer_fit_survives_infeasible_histories():
    """an ego that has just braked to a standstill with a0 < ~-1, or sits near MAX_SPEED with a0 > 0, makes
    the first rows of the QP (st.py:626-668) contradict each other; the interior-point iterate used to turn into NaN and the NaN
    speed command poisoned the episode.  The kernel now replaces a non-finite / infeasible iterate by the plan tracked through
    the limits step by step: the speed command is finite, inside [0, MAX_SPEED] and reachable under the jerk / acceleration clamp."""
    op, p = _params(17)
    eng = EA.EmuEngine(p, max_batch=8)
    S = synthetic.make_states(6, "moderate", seed=4)
    bad = [(0.5, -6.0), (0.0, -6.0), (29.9, 4.5), (0.3, -1.5), (30.0, 2.0), (12.0, 0.3)]
    plan = eng.plan(S)                                         # ordinary plans ...
    assert (plan["reached_t"] == 17).all()
    for b, (v0, a0) in enumerate(bad):                         # ... fitted from histories that contradict the limits
        S["ego"][b, 2], S["ego"][b, 3] = v0, a0
    fine, n_fine, speed, iters = eng.finer_fit(plan["s_seq"], plan["reached_t"], S["ego"])
    assert np.isfinite(speed).all() and all(np.isfinite(fine[b, :n_fine[b]]).all() for b in range(6))
    tick = op.tick_length
    assert (iters[:3] == 41).all() and iters[5] <= 40          # the three states of the report fall back, the ordinary one converges
    for b, (v0, a0) in enumerate(bad):
        assert -1e-9 <= speed[b] <= op.max_speed + 1e-9
        if iters[b] == 41:                                     # the fallback equals control.get_ego_speed_from_jerk of the tracked step
            acc = (speed[b] - v0) / tick
            assert op.a_min - 1e-9 <= acc <= op.a_max + 1e-9 or speed[b] in (0.0, op.max_speed)
    eng.close()


@pytest.mark.parametrize("name,H,n", [("plan_h17.npz", 17, 40), ("plan_h50.npz", 50, 6), ("plan_h25.npz", 25, 12), ("plan_h50b.npz", 50, 6)])
def test_emulated_library_against_the_reference_golden_vectors(name, H, n):
    """The committed outputs of the UNMODIFIED reference (tests/golden/make_golden.py): exact mode reproduces positions, crash
    verdict and start_s bit for bit, fast mode the same plans with the cost within 1e-6; the hinted path returns the same."""
    import os
    G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)))
    S = {k: np.ascontiguousarray(G[k][:n]) for k in ("ego", "cars_x", "cars_v", "cars_a", "n_cars")}
    _op, p = _params(H)
    eng = EA.EmuEngine(p, max_batch=n)
    ex = eng.plan(S, mode=_lib.MODE_EXACT)
    assert np.array_equal(ex["s_seq"], G["s_seq"][:n]) and np.array_equal(ex["crash"].astype(bool), G["crash"][:n])
    assert np.array_equal(ex["start_s"], G["start_s"][:n])
    ok = G["cost"][:n] > 0
    assert np.all(helpers.rel(ex["cost"][ok], G["cost"][:n][ok]) < 1e-12)
    fa = eng.plan(S)
    assert (fa["s_seq"] == G["s_seq"][:n]).all(1).sum() >= n - 1 and np.all(helpers.rel(fa["cost"][ok], G["cost"][:n][ok]) < 1e-6)
    hi = eng.plan_hinted(S, fa["cost"].copy(), hint_scale=1.1)   # (64-bit kernel, 2^-18 labels; the plain plan at H=50: 2^-17)
    _same(hi, fa, keys=tuple(k for k in KEYS if k != "cost"))
    assert np.all(np.abs(hi["cost"] - fa["cost"]) <= 1e-7 * np.maximum(fa["cost"], 1.0))
    eng.close()


def test_error_conventions_and_set_params():
    """Status codes of the C ABI (include/mpcb200.h) and a Settings refresh through mpc_set_params (scratch re-allocated,
    incl. the reachability caps of hinted solves)."""
    _op, p = _params(17)
    eng = EA.EmuEngine(p, max_batch=4)
    L = eng.L
    S = synthetic.make_states(4, "moderate", seed=1, kind="mixed")
    o = eng._out(4)
    a = (EA._p(S["cars_x"]), EA._p(S["cars_v"]), None, EA._p(S["n_cars"]))
    assert L.mpc_plan(eng.h, 0, EA._p(S["ego"]), *a, 0, *eng._outs(o), None) == _lib.MODE_FAST        # B = 0 is not an error
    assert L.mpc_plan(eng.h, 5, EA._p(S["ego"]), *a, 0, *eng._outs(o), None) == _lib.E_CAPACITY
    assert L.mpc_plan(eng.h, -1, EA._p(S["ego"]), *a, 0, *eng._outs(o), None) == _lib.E_INVALID
    assert L.mpc_plan(eng.h, 4, None, *a, 0, *eng._outs(o), None) == _lib.E_INVALID and b"null pointer" in L.mpc_last_error()
    assert L.mpc_plan(eng.h, 4, EA._p(S["ego"]), *a, 7, *eng._outs(o), None) == _lib.E_INVALID
    assert L.mpc_plan_hinted(eng.h, 4, EA._p(S["ego"]), *a, 0, EA._p(np.ones(4)), None, 0, -1.0, *eng._outs(o), None) == _lib.E_INVALID
    assert L.mpc_plan_probed(eng.h, eng.h, 1.1, 4, EA._p(S["ego"]), *a, *eng._outs(o), None) == _lib.E_INVALID
    ref = eng.plan(S)
    _same(eng.plan_hinted(S, ref["cost"].copy(), hint_scale=1.1), ref)
    p2 = _lib.MpcParams()
    for n in _lib.PARAM_FIELDS:
        setattr(p2, n, getattr(p, n))
    p2.crash_min_s = 12.0                                       # the reference's default (config.py:110)
    EA.check(L.mpc_set_params(eng.h, C.byref(p2)))
    ref2 = eng.plan(S)
    assert not np.array_equal(ref2["cost"], ref["cost"])        # other Settings, other plans
    _same(eng.plan_hinted(S, ref2["cost"].copy(), hint_scale=1.1), ref2)
    lone = {k: v[:1].copy() for k, v in S.items()}
    lone["n_cars"][0] = 0                                       # an empty road
    assert eng.plan(lone)["reached_t"][0] == 17
    eng.close()
    # a grid wider than the per-layer lookup tables (288 buckets of 64 cells) is refused, not silently mis-planned
    p3 = _lib.MpcParams()
    for n in _lib.PARAM_FIELDS:
        setattr(p3, n, getattr(p, n))
    p3.future_s = 1000.0                                        # 20002 cells
    h = C.c_void_p()
    assert L.mpc_create(C.byref(p3), 0, 4, 32, C.byref(h)) == _lib.E_CAPACITY and b"lookup tables" in L.mpc_last_error()


@pytest.mark.parametrize("mode", [_lib.MODE_FAST, _lib.MODE_EXACT])
def test_masked_plan_and_fit_touch_only_the_masked_episodes(mode):
    """mpc_plan_masked / mpc_finer_fit_masked: the masked rows equal the plain calls, every other row keeps its contents;
    empty and full masks; 1100 episodes so that the one-block compaction loops."""
    _op, p = _params(17)
    eng = EA.EmuEngine(p, max_batch=64)
    S = synthetic.make_states(40, "moderate", seed=19, kind="mixed")
    ref = eng.plan(S, mode=mode)
    fine_ref, n_ref, speed_ref, _ = eng.finer_fit(ref["s_seq"], ref["reached_t"], S["ego"])
    rng = np.random.default_rng(3)
    for mask in (rng.random(40) < 0.3, np.zeros(40, bool), np.ones(40, bool)):
        m = mask.astype(np.uint8)
        out = eng._out(40)
        for k in ("idx", "s_seq", "cost", "reached_t", "crash", "min_dist"):
            out[k][...] = 77
        eng.plan_masked(S, m, out, mode=mode)
        for k in ("idx", "s_seq", "cost", "reached_t", "crash", "min_dist"):
            assert np.array_equal(out[k][mask], ref[k][mask]), k
            assert np.all(out[k][~mask] == 77), k
        fine, n_fine, speed = np.full_like(fine_ref, -5.0), np.full_like(n_ref, -5), np.full_like(speed_ref, -5.0)
        sseq, reached = np.where(mask[:, None], ref["s_seq"], 1e9), np.where(mask, ref["reached_t"], 999).astype(np.int32)   # junk in unmasked rows
        eng.finer_fit_masked(m, sseq, reached, S["ego"], fine, n_fine, speed)
        assert np.array_equal(speed[mask], speed_ref[mask]) and np.array_equal(n_fine[mask], n_ref[mask])
        for b in np.flatnonzero(mask):
            assert np.array_equal(fine[b, :n_ref[b]], fine_ref[b, :n_ref[b]])
        assert np.all(speed[~mask] == -5.0) and np.all(n_fine[~mask] == -5) and np.all(fine[~mask] == -5.0)
    eng.close()
    if mode == _lib.MODE_FAST:                                   # the compaction itself, across several 1024-thread rounds
        big = EA.EmuEngine(p, max_batch=1100)
        Sb = synthetic.make_states(1100, "moderate", seed=2, kind="mixed")
        mask = np.zeros(1100, np.uint8); mask[[3, 500, 1023, 1024, 1099]] = 1
        out = big._out(1100)
        out["reached_t"][...] = -9
        big.plan_masked(Sb, mask, out)
        assert set(np.flatnonzero(out["reached_t"] != -9)) == {3, 500, 1023, 1024, 1099}
        sub = {k: np.ascontiguousarray(v[[3, 500, 1023, 1024, 1099]]) for k, v in Sb.items()}
        assert np.array_equal(out["cost"][[3, 500, 1023, 1024, 1099]], big.plan(sub)["cost"])
        big.close()


@pytest.mark.parametrize("H", [17, 50])
def test_the_two_row_rasterisers_write_the_same_grids(H, monkeypatch):
    """mpc_build_grid: the segment kernel (default) against the 8-cells-per-thread kernel (MPC_RASTER_ROWS=0) and the oracle, fp32
    and fp64 distances, block sizes 32..256, episodes without cars / with one car / with a grid shorter than the row stride."""
    op, p = _params(H)
    eng = EA.EmuEngine(p, max_batch=16)
    for traffic in ("moderate", "fast"):
        S = synthetic.make_states(10, traffic, seed=11, kind="mixed")
        S["n_cars"][0] = 0
        S["n_cars"][1] = 1
        for f32 in (False, True):
            monkeypatch.setenv("MPC_RASTER_ROWS", "0")
            ref = eng.build_grid(S, f32=f32)
            monkeypatch.delenv("MPC_RASTER_ROWS")
            for threads in ("", "32", "96", "256"):
                if threads:
                    monkeypatch.setenv("MPC_RASTER_THREADS", threads)
                got = eng.build_grid(S, f32=f32)
                monkeypatch.delenv("MPC_RASTER_THREADS", raising=False)
                assert np.array_equal(got["obstacles"], ref["obstacles"]) and np.array_equal(got["distances"], ref["distances"]), (traffic, f32, threads)
        if traffic == "moderate":
            g64 = eng.build_grid(S)
            for b in range(4):
                ob, di, sv = O.build_grid(op, helpers.oracle_state(O, S, b))
                n = sv.size
                assert np.array_equal(g64["obstacles"][b, :, :n], ob) and np.array_equal(g64["distances"][b, :, :n], di)
                assert g64["obstacles"][b, :, n:].all() and not g64["distances"][b, :, n:].any()        # beyond the grid: blocked, 0
    eng.close()


@pytest.mark.parametrize("H,n,traffic", [(17, 20, "moderate"), (17, 12, "low"), (50, 5, "fast")])
def test_dense_solve_runs_the_32bit_key_kernel(H, n, traffic):
    """K2 goes through the same launch chain as the fused planner: fast32_kernel on the dense grids (blocked bits from the mask and
    the zone test d < MIN_ALLOWED_DISTANCE, penalty from the grid's distance), its hand-overs to the 64-bit kernel.  On fp64 grids
    the result is the fused plan bit for bit -- including the problems that were handed over."""
    op, p = _params(H)
    eng = EA.EmuEngine(p, max_batch=32)
    S = synthetic.make_states(n, traffic, seed=9, kind="mixed")
    ref = eng.plan(S)
    fused = eng.fast32_info()
    v0, a0 = S["ego"][:, 2].copy(), S["ego"][:, 3].copy()
    d64 = eng.solve_dense(eng.build_grid(S), v0, a0)
    dense = eng.fast32_info()
    assert dense["in_use"] and dense["handed_on"] == fused["handed_on"] and dense["first_shape_handed_on"] == fused["first_shape_handed_on"]
    assert fused["handed_on"] > 0                                        # the set holds plans that cross a penalty zone
    _same(d64, ref, ("idx", "cost", "reached_t"))
    d32 = eng.solve_dense(eng.build_grid(S, f32=True), v0, a0)
    assert np.array_equal(d32["reached_t"], ref["reached_t"])
    ok = ref["cost"] > 0
    assert np.all(helpers.rel(d32["cost"][ok], ref["cost"][ok]) < 1e-4)
    eng.close()
