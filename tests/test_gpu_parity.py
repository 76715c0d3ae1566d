"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on seeded synthetic states."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

CASES = [("moderate", "onramp"), ("moderate", "mixed"), ("default", "onramp"), ("fast", "mixed"), ("low", "onramp")]


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


@pytest.fixture(scope="module")
def engines(oracle, torch_mod):
    from rl_mpc_lanemerging_b200.engine import MpcEngine
    out = {}
    for H in (17, 50):
        op = oracle.horizon_params(H)
        out[H] = (op, MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=512, nmax=32))
    yield out
    for _, e in out.values():
        e.close()


def _states(traffic, kind, B, seed):
    from rl_mpc_lanemerging_b200 import synthetic
    return synthetic.make_states(B, traffic, seed=seed, kind=kind)


def _dev(S, torch):
    from rl_mpc_lanemerging_b200.engine import states_to_device
    return states_to_device(S, "cuda:0")


@pytest.mark.parametrize("H,B", [(17, 96), (50, 24)])
@pytest.mark.parametrize("traffic,kind", CASES)
def test_plan_exact_is_index_identical(oracle, engines, torch_mod, H, B, traffic, kind):
    op, eng = engines[H]
    S = _states(traffic, kind, B, seed=11)
    ref = helpers.oracle_plan_batch(oracle, op, S, eng.num_t)
    D = _dev(S, torch_mod)
    out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode="exact")
    torch_mod.cuda.synchronize()
    idx = out["idx"].cpu().numpy()
    assert np.array_equal(out["reached_t"].cpu().numpy(), ref["reached_t"])
    assert np.array_equal(idx, ref["idx"]), f"first mismatch at problem {np.argwhere((idx != ref['idx']).any(1))[:5].ravel()}"
    assert np.array_equal(out["s_seq"].cpu().numpy(), ref["s_seq"])          # bit-identical fp64 positions
    assert np.array_equal(out["cost"].cpu().numpy(), ref["cost"])            # bit-identical DP cost
    assert np.array_equal(out["crash"].cpu().numpy().astype(bool), ref["crash"])
    md = out["min_dist"].cpu().numpy()
    assert np.array_equal(md, ref["min_path_distance"])


@pytest.mark.parametrize("H,B", [(17, 96), (50, 24)])
@pytest.mark.parametrize("traffic,kind", CASES)
def test_plan_fast_within_tolerance(oracle, engines, torch_mod, H, B, traffic, kind):
    """north_star tolerance: DP cost within 1e-4 rel; identical cell sequence, or the oracle-evaluated
    cost of our sequence within 1e-4 rel of the oracle's optimum; identical reached_t / zero pattern."""
    op, eng = engines[H]
    S = _states(traffic, kind, B, seed=12)
    ref = helpers.oracle_plan_batch(oracle, op, S, eng.num_t)
    D = _dev(S, torch_mod)
    out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode="fast")
    torch_mod.cuda.synchronize()
    idx, cost = out["idx"].cpu().numpy(), out["cost"].cpu().numpy()
    assert np.array_equal(out["reached_t"].cpu().numpy(), ref["reached_t"])
    assert np.array_equal(idx < 0, ref["idx"] < 0)
    ok = ref["cost"] > 0
    assert np.all(helpers.rel(cost[ok], ref["cost"][ok]) <= 1e-4), helpers.rel(cost[ok], ref["cost"][ok]).max()
    same = (idx == ref["idx"]).all(1)
    n_diff = int((~same).sum())
    for b in np.nonzero(~same)[0]:
        st = helpers.oracle_state(oracle, S, b)
        obst, dist, sv = oracle.build_grid(op, st)
        n = int(ref["reached_t"][b]) + 1
        ours = oracle.path_cost(op, idx[b, :n], sv, dist, op.t_disc, st.ego_v, st.ego_a)
        assert abs(ours - ref["cost"][b]) <= 1e-4 * ref["cost"][b], (b, ours, ref["cost"][b])
        # and the path must be feasible on the oracle's grid
        assert not obst[np.arange(1, n), idx[b, 1:n]].any()
    assert n_diff <= max(1, B // 16), f"{n_diff} of {B} sequences differ from the oracle"
    # the kernels' arithmetic is modelled on the CPU (oracle/mpc_oracle.c: orc_solve_fast_model_q / orc_solve_fast_model): bit-identical
    info = eng.fast32_info()
    for b in range(0, B, 7):
        ok, m = helpers.fast_mode_matches_model(oracle, info, op, helpers.oracle_state(oracle, S, b), idx[b], cost[b])
        assert ok, (b, m["cost"], cost[b])
    # first-step acceleration (what the controller acts on)
    a_ref = ((ref["s_seq"][:, 1] - ref["s_seq"][:, 0]) / op.t_disc - S["ego"][:, 2]) / op.t_disc
    s = out["s_seq"].cpu().numpy()
    a_our = ((s[:, 1] - s[:, 0]) / op.t_disc - S["ego"][:, 2]) / op.t_disc
    m = same & (ref["reached_t"] >= 1)
    assert np.all(np.abs(a_our[m] - a_ref[m]) <= 1e-4 * np.maximum(np.abs(a_ref[m]), 1.0))
    assert np.array_equal(out["crash"].cpu().numpy().astype(bool)[same], ref["crash"][same])


@pytest.mark.parametrize("traffic,kind", CASES[:3])
def test_build_grid_matches_oracle(oracle, engines, torch_mod, traffic, kind):
    op, eng = engines[17]
    S = _states(traffic, kind, 12, seed=13)
    D = _dev(S, torch_mod)
    g = eng.build_grid(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    torch_mod.cuda.synchronize()
    ob, di, ns = g["obstacles"].cpu().numpy(), g["distances"].cpu().numpy(), g["num_s"].cpu().numpy()
    for b in range(12):
        o2, d2, sv = oracle.build_grid(op, helpers.oracle_state(oracle, S, b))
        assert ns[b] == sv.size
        assert g["start_s"][b].item() == sv[0] and g["delta_s"][b].item() == sv[1] - sv[0]
        assert np.array_equal(ob[b, :, :ns[b]], o2)
        assert np.array_equal(di[b, :, :ns[b]], d2)              # bit-identical fp64 distance field
    # fp32 distances = the fp64 field rounded once; the 8-cells-per-thread kernel (MPC_RASTER_ROWS=0) writes the same bytes
    g32 = eng.build_grid(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], dist_dtype=torch_mod.float32)
    assert np.array_equal(g32["distances"].cpu().numpy(), di.astype(np.float32)) and np.array_equal(g32["obstacles"].cpu().numpy(), ob)
    import os
    os.environ["MPC_RASTER_ROWS"] = "0"
    try:
        for dt, want in ((torch_mod.float64, di), (torch_mod.float32, di.astype(np.float32))):
            g0 = eng.build_grid(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], dist_dtype=dt)
            assert np.array_equal(g0["distances"].cpu().numpy(), want) and np.array_equal(g0["obstacles"].cpu().numpy(), ob)
    finally:
        del os.environ["MPC_RASTER_ROWS"]


@pytest.mark.parametrize("H", [17, 50])
@pytest.mark.parametrize("traffic,kind", CASES)
def test_sorted_search_structure_is_exact(engines, torch_mod, H, traffic, kind):
    """The fast kernel's O(1) obstacle/distance lookup equals the reference-order evaluation on every cell, and the
    blocked intervals of the lean bounded pass are exactly the cells in a band or inside a penalty zone (device self-test)."""
    _, eng = engines[H]
    D = _dev(_states(traffic, kind, 64, seed=16), torch_mod)
    assert eng.selftest_search(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"]) == 0


@pytest.mark.parametrize("H,env", [(17, {"MPC_FAST_BOUND": "0"}), (17, {"MPC_FAST_BLOCKS": "64", "MPC_FAST_THREADS": "512"}),
                                   (17, {"MPC_FAST_BLOCKS": "32", "MPC_FAST_THREADS": "1024"}), (50, {"MPC_FAST_BOUND": "0"}),
                                   (50, {"MPC_FAST_BLOCKS": "96", "MPC_FAST_THREADS": "384"}),
                                   (17, {"MPC_FAST32": "0"}), (50, {"MPC_FAST32": "0"}),
                                   (17, {"MPC_F32_BLOCKS": "64", "MPC_F32_THREADS": "512"}), (50, {"MPC_F32_BLOCKS": "64", "MPC_F32_THREADS": "1024"}),
                                   (50, {"MPC_F32_BLOCKS": "96", "MPC_F32_THREADS": "256", "MPC_FAST_BLOCKS": "96"}),
                                   (50, {"MPC_HANDOVER_ALT": "2"}), (17, {"MPC_HANDOVER_ALT": "2"}), (50, {"MPC_SIDE_STREAM": "0"})])
def test_fast_result_does_not_depend_on_bound_or_launch_shape(oracle, engines, torch_mod, monkeypatch, H, env):
    """The 32-bit-key kernel followed by the 64-bit kernel (default), the 64-bit kernel alone (MPC_FAST32=0: lean bounded first
    pass, retry without the bound inside the kernel), its plain unbounded pass (MPC_FAST_BOUND=0, which also switches the
    32-bit-key kernel off) and every launch shape / ring size of either (H=50 with 3 blocks per SM: some frontiers outgrow the ring
    and are handed on / re-solved with a full row) give identical plans.  Bit for bit except the cost at H > 25, where the
    32-bit-key kernel carries 2^-17 labels and the 64-bit kernel 2^-18 (DevParams::f32_frac): equal to 1e-7 there."""
    from rl_mpc_lanemerging_b200.engine import MpcEngine
    op, eng = engines[H]
    S = _states("moderate", "mixed", 64, seed=21)
    D = _dev(S, torch_mod)
    a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    ref = {k: v.cpu().numpy() for k, v in eng.plan(*a, mode="fast").items()}
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    other = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=64, nmax=32)   # reads the overrides at create time
    try:
        out = {k: v.cpu().numpy() for k, v in other.plan(*a, mode="fast").items()}
    finally:
        other.close()
    for k in ("idx", "s_seq", "reached_t", "crash", "min_dist"):
        assert np.array_equal(out[k], ref[k]), k
    if eng.fast32_info()["frac_bits"] == 18:
        assert np.array_equal(out["cost"], ref["cost"])
    else:
        assert np.all(np.abs(out["cost"] - ref["cost"]) <= 1e-7 * np.maximum(ref["cost"], 1.0))
    assert (ref["reached_t"] < H).any() and (ref["reached_t"] == H).any()


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_solve_dense_dropin(oracle, engines, torch_mod, mode):
    """K2: the st_cy.solve_s_t_path_fast replacement on dense grids built by the oracle."""
    torch = torch_mod
    op, eng = engines[17]
    S = _states("moderate", "mixed", 16, seed=14)
    T, Smax = eng.num_t, eng.num_s_max
    ob = np.ones((16, T, Smax), np.uint8)
    di = np.zeros((16, T, Smax), np.float64)
    s0 = np.zeros(16); ds = np.zeros(16); ns = np.zeros(16, np.int32)
    refs = []
    for b in range(16):
        st = helpers.oracle_state(oracle, S, b)
        o2, d2, sv = oracle.build_grid(op, st)
        ob[b, :, :sv.size] = o2; di[b, :, :sv.size] = d2
        s0[b], ds[b], ns[b] = sv[0], sv[1] - sv[0], sv.size
        refs.append(oracle.solve(op, o2, d2, sv, op.t_disc, st.ego_v, st.ego_a))
    dev = "cuda:0"
    out = eng.solve_dense(torch.from_numpy(ob).to(dev), torch.from_numpy(di).to(dev), torch.from_numpy(s0).to(dev),
                          torch.from_numpy(ds).to(dev), torch.from_numpy(ns).to(dev),
                          torch.from_numpy(S["ego"][:, 2].copy()).to(dev), torch.from_numpy(S["ego"][:, 3].copy()).to(dev), mode=mode)
    torch.cuda.synchronize()
    idx = out["idx"].cpu().numpy(); cost = out["cost"].cpu().numpy(); seq = out["s_seq"].cpu().numpy()
    for b, r in enumerate(refs):
        assert out["reached_t"][b].item() == r["reached_t"]
        if mode == "exact":
            assert np.array_equal(idx[b], r["idx"]) and np.array_equal(seq[b], r["s_seq"]) and cost[b] == r["cost"]
        elif r["cost"] > 0:
            assert abs(cost[b] - r["cost"]) <= 1e-4 * r["cost"]


def test_rollout_step_pieces(oracle, engines, torch_mod):
    """K4: predict_step_with_ego, observation vector, jerk->speed against the oracle."""
    torch = torch_mod
    op, eng = engines[17]
    S = _states("moderate", "mixed", 64, seed=15)
    rng = np.random.default_rng(0)
    sel = rng.uniform(0, 30, 64)
    jerk = rng.uniform(-6, 6, 64)
    D = _dev(S, torch)
    eo, xo, vo, ao, cr = eng.predict_step_with_ego(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"],
                                                   torch.from_numpy(sel).to("cuda:0"), 0.2, 5.1)
    sv = eng.state_vector(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    sp = eng.speed_from_jerk(D["ego"], torch.from_numpy(jerk).to("cuda:0"))
    torch.cuda.synchronize()
    eo, xo, vo, ao, cr, sv, sp = [t.cpu().numpy() for t in (eo, xo, vo, ao, cr, sv, sp)]
    for b in range(64):
        st = helpers.oracle_state(oracle, S, b)
        nxt, crashed = oracle.predict_step_with_ego(op, st, sel[b], 0.2, 5.1)
        n = st.n
        assert np.array_equal(eo[b], [nxt.ego_x, nxt.ego_y, nxt.ego_v, nxt.ego_a])
        assert np.array_equal(xo[b, :n], nxt.x[:n]) and np.array_equal(vo[b, :n], nxt.v[:n]) and np.array_equal(ao[b, :n], nxt.a[:n])
        assert bool(cr[b]) == crashed
        assert np.allclose(sv[b, :20], oracle.state_vector(op, st).astype(np.float32), rtol=0, atol=1e-7)
        assert sp[b] == oracle.speed_from_jerk(op, st.ego_v, st.ego_a, jerk[b])


@pytest.mark.parametrize("H,traffic,n", [(17, "moderate", 192), (17, "low", 96), (50, "moderate", 96), (50, "fast", 64)])
def test_dense_solve_equals_the_fused_plan(engines, torch_mod, H, traffic, n):
    """K2 through the 32-bit-key kernel on dense grids (mpc_build_grid output): on fp64 grids bit-identical to the fused plan --
    hand-overs to the 64-bit kernel included --, on fp32 distances same horizon reached and cost within 1e-4."""
    torch = torch_mod
    _, eng = engines[H]
    D = _dev(_states(traffic, "mixed", n, seed=23), torch)
    a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    ref = {k: v.clone() for k, v in eng.plan(*a, mode="fast").items()}
    fused = eng.fast32_info()
    v0, a0 = D["ego"][:, 2].contiguous(), D["ego"][:, 3].contiguous()
    g = eng.build_grid(*a)
    d64 = eng.solve_dense(g["obstacles"], g["distances"], g["start_s"], g["delta_s"], g["num_s"], v0, a0, mode="fast")
    dense = eng.fast32_info()
    assert dense["in_use"] and dense["handed_on"] == fused["handed_on"] > 0
    for k in ("idx", "s_seq", "cost", "reached_t"):
        assert torch.equal(d64[k], ref[k]), k
    g = eng.build_grid(*a, dist_dtype=torch.float32)
    d32 = eng.solve_dense(g["obstacles"], g["distances"], g["start_s"], g["delta_s"], g["num_s"], v0, a0, mode="fast")
    assert torch.equal(d32["reached_t"], ref["reached_t"])
    ok = ref["cost"] > 0
    assert bool(((d32["cost"][ok] - ref["cost"][ok]).abs() <= 1e-4 * ref["cost"][ok]).all())


def test_wide_ring_launch_next_to_the_handover_launch(oracle, engines, torch_mod, monkeypatch):
    """The problems whose frontier outgrew the first ring (second launch shape of the 32-bit-key kernel, side stream) and the
    flagged problems (64-bit kernel, caller's stream) are solved at the same time, each launch with its own scratch: low traffic at
    H=50 with a forced small ring has dozens of both; plans equal to the 64-bit kernel's alone, on every repetition."""
    from rl_mpc_lanemerging_b200.engine import MpcEngine
    op, _ = engines[50]
    S = _states("low", "onramp", 512, seed=31)
    D = _dev(S, torch_mod)
    a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    monkeypatch.setenv("MPC_FAST32", "0")
    alone = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=512, nmax=32)
    ref = {k: v.clone() for k, v in alone.plan(*a, mode="fast").items()}
    alone.close()
    monkeypatch.delenv("MPC_FAST32")
    monkeypatch.setenv("MPC_F32_BLOCKS", "96")
    eng = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=512, nmax=32)
    try:
        for rep in range(4):
            out = eng.plan(*a, mode="fast")
            info = eng.fast32_info()
            for k in ("idx", "s_seq", "reached_t", "crash", "min_dist"):
                assert torch_mod.equal(out[k], ref[k]), (k, rep)
            assert bool(((out["cost"] - ref["cost"]).abs() <= 1e-6 * ref["cost"].clamp(min=1.0)).all())      # 2^-17 against 2^-18 labels
            if rep == 0:
                wide = info["first_shape_handed_on"] - info["handed_on"]           # solved by the wide ring
                assert wide >= 10 and info["handed_on"] >= 10, info
    finally:
        eng.close()
