"""-m gpu: edge cases the reference's semantics define (empty / full car lists, states past the merge, infeasible
plans, tiny and ragged batches) and size-independent properties at BASELINE.json's full sizes (4096 episodes, H=50)."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng17(oracle):
    from rl_mpc_lanemerging_b200.engine import MpcEngine
    op = oracle.default_params()
    e = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=64)
    yield op, e
    e.close()


def _batch(states, nmax=32):
    B = len(states)
    S = dict(ego=np.zeros((B, 4)), cars_x=np.zeros((B, nmax)), cars_v=np.zeros((B, nmax)), cars_a=np.zeros((B, nmax)),
             n_cars=np.zeros(B, np.int32))
    for b, (pos, v, a, xs, vs, acs) in enumerate(states):
        S["ego"][b] = [pos[0], pos[1], v, a]
        n = len(xs)
        S["cars_x"][b, :n] = xs; S["cars_v"][b, :n] = vs; S["cars_a"][b, :n] = acs; S["n_cars"][b] = n
    return S


EDGE_STATES = [
    ((-200.0, 21.7), 15.0, 0.0, [], [], []),                                            # no cars at all (SURVEY anchor A)
    ((-60.0, 3.0), 0.0, 0.0, [], [], []),                                               # standing start: speed clamp at 0
    ((-60.0, 3.0), 30.0, 4.5, [], [], []),                                              # already at MAX_SPEED: upper clamp on every node
    ((10.0, -1.6), 11.0, 0.0, [30.0, -8.0, -30.0], [11.0] * 3, [0.0] * 3),             # merged, in traffic (anchor C)
    ((5.0, -1.6), 11.0, 0.0, [7.0, 2.0], [11.0, 11.0], [0.0, 0.0]),                    # boxed in: plan is infeasible
    ((-30.0, -1.0), 20.0, -2.0, list(np.linspace(60, -95, 32)), [9.0] * 32, [0.0] * 32),   # the full 32 cars, 5 m apart (bands touch)
    ((60.0, -1.6), 25.0, 0.0, [55.0, 40.0], [5.0, 5.0], [-4.0, -4.0]),                 # ego ahead of every car, braking followers
    ((-49.0, 1.0), 12.0, 1.0, [-48.0], [0.0], [0.0]),                                   # stopped car right at the merge point
    ((-120.0, 11.0), 8.0, -6.0, [100.0, 99.9, 99.8], [11.0, 11.0, 11.0], [0.0] * 3),   # overlapping cars far ahead, hard braking start
]


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_edge_states(oracle, eng17, mode):
    import torch
    from rl_mpc_lanemerging_b200.engine import states_to_device
    op, eng = eng17
    S = _batch(EDGE_STATES)
    ref = helpers.oracle_plan_batch(oracle, op, S, 18, layered=False)
    D = states_to_device(S, "cuda:0")
    out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode=mode)
    torch.cuda.synchronize()
    assert np.array_equal(out["reached_t"].cpu().numpy(), ref["reached_t"])
    assert (ref["reached_t"] < 17).any() and (ref["reached_t"] == 17).any()            # both feasible and infeasible plans present
    idx = out["idx"].cpu().numpy()
    if mode == "exact":
        assert np.array_equal(idx, ref["idx"]) and np.array_equal(out["cost"].cpu().numpy(), ref["cost"])
    else:
        ok = ref["cost"] > 0
        assert np.all(helpers.rel(out["cost"].cpu().numpy()[ok], ref["cost"][ok]) < 1e-6)
        assert (idx == ref["idx"]).all(1).sum() >= len(EDGE_STATES) - 1
    # the crash verdict (st.py:790-802) is a function of the chosen sequence: wherever the sequence is the oracle's, so is the verdict
    same = (idx == ref["idx"]).all(1)
    assert same.all() or mode == "fast"
    assert np.array_equal(out["crash"].cpu().numpy().astype(bool)[same], ref["crash"][same])
    assert eng.selftest_search(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"]) == 0


def test_batch_sizes_and_capacity(oracle, eng17):
    import torch
    from rl_mpc_lanemerging_b200 import _lib, synthetic
    from rl_mpc_lanemerging_b200.engine import states_to_device
    op, eng = eng17
    S = synthetic.make_states(64, "default", seed=31, kind="mixed")
    ref = helpers.oracle_plan_batch(oracle, op, S, 18)
    for B in (1, 2, 33, 64):                                                            # ragged batches through the same handle
        Sb = {k: np.ascontiguousarray(v[:B]) for k, v in S.items()}
        D = states_to_device(Sb, "cuda:0")
        out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode="exact")
        assert np.array_equal(out["idx"].cpu().numpy(), ref["idx"][:B])
    D = states_to_device({k: np.concatenate([v, v]) for k, v in S.items()}, "cuda:0")    # 128 > max_batch 64
    with pytest.raises(_lib.MpcError) as e:
        eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    assert e.value.code == _lib.E_CAPACITY
    host = eng.plan_host(S["ego"], S["cars_x"], S["cars_v"], S["cars_a"], S["n_cars"], mode="exact")   # host-buffer entry point
    assert np.array_equal(host["idx"], ref["idx"]) and np.array_equal(host["s_seq"], ref["s_seq"])
    torch.cuda.synchronize()


def test_full_size_properties_h50(oracle):
    """BASELINE.json configs[1] (4096 episodes, moderate traffic, H=50): properties that do not need the oracle on every
    episode -- run-to-run determinism, exact == fast sequences, kinematic feasibility of every step, paths avoid the
    obstacle bands of an independently built grid -- plus the oracle itself on a sample."""
    import torch
    from rl_mpc_lanemerging_b200 import synthetic
    from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
    op = oracle.horizon_params(50)
    B = 4096
    eng = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=B)
    S = synthetic.make_states(B, "moderate", seed=0)
    D = states_to_device(S, "cuda:0")
    a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    f1 = {k: v.clone() for k, v in eng.plan(*a, mode="fast").items()}
    f2 = eng.plan(*a, mode="fast")
    for k in f1:
        assert torch.equal(f1[k], f2[k]), f"fast mode is not deterministic in {k}"
    ex = eng.plan(*a, mode="exact")
    same = (f1["idx"] == ex["idx"]).all(1)
    assert same.float().mean().item() >= 0.999
    assert torch.equal(f1["reached_t"], ex["reached_t"])
    okc = ex["cost"] > 0
    assert ((f1["cost"] - ex["cost"]).abs()[okc] / ex["cost"][okc]).max().item() < 1e-6
    # kinematic feasibility of every planned step in cell units (st_cy.pyx:65-75): 0 <= v <= 180, -10 <= a <= 8, |jerk| <= 2
    idx = f1["idx"].long(); T = idx.shape[1]
    valid = idx >= 0
    v = idx[:, 1:] - idx[:, :-1]
    vv = valid[:, 1:] & valid[:, :-1]
    assert bool(((v >= 0) & (v <= 180))[vv].all())
    acc = v[:, 2:] - v[:, 1:-1]                                                          # from layer 2 on every point is a grid cell
    va = vv[:, 2:] & vv[:, 1:-1]
    assert bool(((acc >= -10) & (acc <= 8))[va].all())
    jerk = acc[:, 1:] - acc[:, :-1]
    assert bool((jerk.abs() <= 2)[va[:, 1:] & va[:, :-1]].all())
    # a sample against the oracle and against an independently rasterised grid
    sel = np.arange(0, B, 128)
    sub = {k: np.ascontiguousarray(vv_[sel]) for k, vv_ in S.items()}
    ref = helpers.oracle_plan_batch(oracle, op, sub, 51)
    assert np.array_equal(ex["idx"].cpu().numpy()[sel], ref["idx"]) and np.array_equal(ex["cost"].cpu().numpy()[sel], ref["cost"])
    Dsub = states_to_device(sub, "cuda:0")
    g = eng.build_grid(Dsub["ego"], Dsub["cars_x"], Dsub["cars_v"], Dsub["cars_a"], Dsub["n_cars"], dist_dtype=torch.float32)
    ob = g["obstacles"].cpu().numpy(); fi = f1["idx"].cpu().numpy()[sel]
    for i in range(len(sel)):
        n = int(ref["reached_t"][i]) + 1
        assert not ob[i, np.arange(1, n), fi[i, 1:n]].any()
    eng.close()


@pytest.mark.parametrize("H,traffic,B", [(25, "low", 16), (25, "fast", 16), (100, "default", 6), (100, "moderate", 6)])
def test_horizon_and_density_sweep(oracle, H, traffic, B):
    """BASELINE.json configs[4]: traffic-density sweep x horizon in {25, 100} (50 and 17 are covered elsewhere).  Exact mode is
    index-identical to the oracle, fast mode within tolerance (H=100 runs the fast kernel behind a ring window)."""
    import torch
    from rl_mpc_lanemerging_b200 import synthetic
    from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
    op = oracle.horizon_params(H)
    eng = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=B)
    assert eng.num_t == H + 1
    S = synthetic.make_states(B, traffic, seed=37, kind="mixed")
    ref = helpers.oracle_plan_batch(oracle, op, S, H + 1)
    D = states_to_device(S, "cuda:0")
    a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    ex = eng.plan(*a, mode="exact")
    assert np.array_equal(ex["idx"].cpu().numpy(), ref["idx"]) and np.array_equal(ex["cost"].cpu().numpy(), ref["cost"])
    fa = eng.plan(*a, mode="fast")
    torch.cuda.synchronize()
    assert np.array_equal(fa["reached_t"].cpu().numpy(), ref["reached_t"])
    ok = ref["cost"] > 0
    assert np.all(helpers.rel(fa["cost"].cpu().numpy()[ok], ref["cost"][ok]) < 1e-6)
    assert (fa["idx"].cpu().numpy() == ref["idx"]).all(1).sum() >= B - 1
    eng.close()


@pytest.mark.parametrize("over", [dict(start_uncertainty=0.5, uncertainty_per_second=0.3), dict(min_allowed_distance=6.0, crash_min_s=15.0),
                                  dict(d_weight=100.0, v_weight=1.0), dict(car_length=4.0, min_allowed_distance=3.0)])
def test_other_settings_keep_parity(oracle, over):
    """Settings away from the published configs (prediction uncertainty widens the bands, other safety distances / weights --
    the reference's grid-search ranges, main.py:43-52): exact mode index-identical, fast mode bit-identical to its CPU model, and
    the blocked intervals of the lean bounded pass still equal 'in a band or inside a penalty zone' on every cell."""
    import torch
    from rl_mpc_lanemerging_b200 import synthetic
    from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
    op = oracle.horizon_params(17, **over)
    B = 48
    eng = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=B)
    S = synthetic.make_states(B, "moderate", seed=41, kind="mixed")
    ref = helpers.oracle_plan_batch(oracle, op, S, 18)
    D = states_to_device(S, "cuda:0")
    a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    assert eng.selftest_search(*a) == 0
    ex = eng.plan(*a, mode="exact")
    assert np.array_equal(ex["idx"].cpu().numpy(), ref["idx"]) and np.array_equal(ex["cost"].cpu().numpy(), ref["cost"])
    fa = {k: v.cpu().numpy() for k, v in eng.plan(*a, mode="fast").items()}
    torch.cuda.synchronize()
    assert np.array_equal(fa["reached_t"], ref["reached_t"])
    ok = ref["cost"] > 0
    assert np.all(helpers.rel(fa["cost"][ok], ref["cost"][ok]) <= 1e-4)
    info = eng.fast32_info()
    for b in range(0, B, 5):
        ok, m = helpers.fast_mode_matches_model(oracle, info, op, helpers.oracle_state(oracle, S, b), fa["idx"][b], fa["cost"][b])
        assert ok, (b, m["cost"], fa["cost"][b])
    eng.close()
