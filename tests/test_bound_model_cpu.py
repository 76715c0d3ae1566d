"""CPU checks of the fast kernel's pruning rules (oracle/bound_model.py, the C model in oracle/mpc_oracle.c).

The GPU kernel is tested bit for bit against the same C model (tests/test_gpu_parity.py); these tests establish,
without a GPU, that the pruning rules themselves never change the answer:
  * any cost bound is exact when the bounded pass reaches the horizon, and the retry ladder
    (hint -> zone bound -> unbounded) always ends on the unbounded answer;
  * the cell-only reachability heuristic h (label + h > U dropped) is exact in the same sense;
  * the coarse probe plan (mpc_plan_probed) is a usable estimate of the plan's cost.
"""
import numpy as np
import pytest

from oracle import bound_model as BM
from oracle import cpu_oracle as O
from rl_mpc_lanemerging_b200 import synthetic


def _states(n, traffic, seed, kind="onramp"):
    S = synthetic.make_states(n, traffic, seed=seed, kind=kind)
    out = []
    for i in range(n):
        e, k = S["ego"][i], S["n_cars"][i]
        out.append(O.make_state((e[0], e[1]), e[2], e[3], S["cars_x"][i, :k], S["cars_v"][i, :k], S["cars_a"][i, :k]))
    return out


def _same(a, b):
    return a["reached_t"] == b["reached_t"] and np.array_equal(a["idx"], b["idx"]) and a["cost"] == b["cost"]


@pytest.mark.parametrize("H,traffic,kind", [(17, "moderate", "mixed"), (25, "default", "onramp"), (50, "moderate", "onramp")])
def test_any_bound_that_reaches_the_horizon_is_exact(H, traffic, kind):
    p = O.horizon_params(H)
    n_bounded_ok = 0
    for st in _states(10, traffic, seed=11, kind=kind):
        ob, di, sv = O.build_grid(p, st)
        ref = O.solve_fast_model_ex(p, ob, di, sv, p.t_disc, st.ego_v, st.ego_a, prune_cost=0.0)
        for f in (0.3, 0.8, 0.999, 1.0, 1.001, 1.1, 2.0, 300.0):
            U = max(ref["cost"] * f, 1e-3)
            r = O.solve_fast_model_ex(p, ob, di, sv, p.t_disc, st.ego_v, st.ego_a, prune_cost=U)
            if r["reached_t"] == H:                       # reached the horizon under the bound: must be THE answer
                assert _same(r, ref), (f, r["cost"], ref["cost"])
                n_bounded_ok += 1
                assert r["nodes"] <= ref["nodes"]
            else:                                          # cut short: only allowed when the bound is below the answer
                assert ref["reached_t"] < H or U < ref["cost"]
            # the kernel's ladder always ends on the unbounded answer, whatever the hint
            lad, _, tries = BM.solve_with_ladder(p, ob, di, sv, st.ego_v, st.ego_a, hint=U)
            assert _same(lad, ref) and 1 <= tries <= 4
        lad, _, _ = BM.solve_with_ladder(p, ob, di, sv, st.ego_v, st.ego_a, hint=None)
        assert _same(lad, ref)
    assert n_bounded_ok >= 20


@pytest.mark.parametrize("H,traffic", [(17, "moderate"), (50, "moderate"), (50, "fast")])
def test_reachability_heuristic_is_exact_and_prunes(H, traffic):
    p = O.horizon_params(H)
    zb = BM.zone_bound(p)
    plain = pruned = 0
    checked = 0
    for st in _states(8, traffic, seed=5):
        ob, di, sv = O.build_grid(p, st)
        ref = O.solve_fast_model_ex(p, ob, di, sv, p.t_disc, st.ego_v, st.ego_a, prune_cost=0.0)
        if ref["reached_t"] < H or ref["cost"] >= zb:      # (plans that cross a penalty zone never use the heuristic)
            continue
        h = BM.heuristic_table(p, ob, di)
        # admissible along the answer: label + h never exceeds the final cost
        lab = 0.0
        for t in range(1, H + 1):
            lab = O.path_cost(p, ref["idx"][:t + 1], sv, di, p.t_disc, st.ego_v, st.ego_a)
            assert lab + float(h[t, ref["idx"][t]]) / BM.FX_ONE <= ref["cost"] * (1 + 1e-6) + 1e-6   # (fp64 path cost vs fixed-point label: 1e-7 rel)
        for f in (1.0, 1.02, 1.1, 1.5):
            U = ref["cost"] * f + 2.0 / BM.FX_ONE
            r = BM.solve_with_heuristic(p, ob, di, sv, st.ego_v, st.ego_a, U, h)
            assert r["reached_t"] == H and _same(r, ref), (f, r["reached_t"], r["cost"], ref["cost"])
            checked += 1
            if f == 1.1:
                pruned += r["nodes"]
                plain += O.solve_fast_model_ex(p, ob, di, sv, p.t_disc, st.ego_v, st.ego_a, prune_cost=U)["nodes"]
    assert checked >= 16
    # measured at U = 1.1 C*: 0.76 of the nodes at H=50 moderate (60 states); at H=17 the horizon ends before the traffic confines the ego (0.97)
    assert pruned <= plain and (H < 50 or pruned < 0.9 * plain), (pruned, plain)


def test_reach_caps_kernel_algorithm_matches_the_mask_based_caps():
    """reach_caps_kernel works on the blocked INTERVALS of a layer (LayerDesc::blk); its loops, restated in
    bound_model.bucket_caps_like_kernel, must give the caps heuristic_table derives from the dense blocked mask."""
    for H, traffic, kind in ((50, "moderate", "onramp"), (17, "default", "mixed"), (25, "fast", "onramp")):
        p = O.horizon_params(H)
        for st in _states(5, traffic, seed=9, kind=kind):
            ob, di, sv = O.build_grid(p, st)
            nb = (ob.shape[1] + 63) // 64
            caps = np.full((ob.shape[0], nb), -1, np.int64)
            BM.heuristic_table(p, ob, di, caps_out=caps)
            assert np.array_equal(caps, BM.bucket_caps_like_kernel(p, ob, di))


def test_probe_plan_estimates_the_cost():
    H = 50
    p = O.horizon_params(H)
    q = BM.probe_params(p, 20, 3)
    zb = BM.zone_bound(p)
    ratios, base_nodes, hinted_nodes, first_ok, tried = [], 0, 0, 0, 0
    for st in _states(40, "moderate", seed=0):
        ob, di, sv = O.build_grid(p, st)
        ref, nb, _ = BM.solve_with_ladder(p, ob, di, sv, st.ego_v, st.ego_a, hint=None)
        hint, _ = BM.probe_hint(p, q, st, margin=1.1)
        lad, nh, tries = BM.solve_with_ladder(p, ob, di, sv, st.ego_v, st.ego_a, hint=hint)
        assert _same(lad, ref)                              # the hint never changes the answer
        base_nodes += nb
        hinted_nodes += nh
        if hint is not None and hint < zb and ref["reached_t"] == H and ref["cost"] < zb:
            ratios.append(hint / 1.1 / ref["cost"])
            tried += 1
            first_ok += int(tries == 1)
    ratios = np.array(ratios)
    assert len(ratios) >= 25
    assert 0.9 < np.median(ratios) < 1.1, np.median(ratios)         # measured: 1.00 (IQR 0.97-1.04)
    assert first_ok >= 0.6 * tried, (first_ok, tried)               # measured: ~0.8
    assert hinted_nodes < 0.92 * base_nodes, (hinted_nodes, base_nodes)   # measured: ~0.77 over 150 states


def test_kernel_division_by_the_layer_constant_is_exact():
    """The hinted kernel divides D (<= 250 * 127 cells) by n = steps to the horizon as (int)((float)D * (1.0f / n)) followed by
    one correction step (mpc_fast.cu); that must equal D div n, D mod n for every D and n it can meet."""
    D = np.arange(0, 250 * 127 + 1, dtype=np.int64)
    for n in range(1, 128):
        rcp = np.float32(1.0) / np.float32(n)
        q = (D.astype(np.float32) * rcp).astype(np.int64)          # truncation, like the cast in the kernel
        r = D - q * n
        lo, hi = r < 0, r >= n
        q = q - lo + hi
        r = r + lo * n - hi * n
        assert np.array_equal(q, D // n) and np.array_equal(r, D % n), n
