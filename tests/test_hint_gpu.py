"""GPU tests of the cost-hint entry points (mpc_plan_hinted, mpc_plan_probed).

The property under test is the one tests/test_bound_model_cpu.py establishes for the CPU model: a hint changes how
many nodes the fast kernel expands, never its outputs.

Hinted solves run on the 64-bit kernel (2^-18 labels).  The plain plan of the default path comes from the 32-bit-key kernel, whose
labels carry 2^-17 at H > 25 (DevParams::f32_frac): every output is compared bit for bit, the cost to 1e-7 where the precisions differ.
"""
import numpy as np
import pytest
import torch

from rl_mpc_lanemerging_b200 import engine as E
from rl_mpc_lanemerging_b200 import synthetic


def _engine(H, B):
    p = E._lib.default_params()
    p.future_t, p.future_s = synthetic.horizon_settings(H)
    return E.MpcEngine(p, device=0, max_batch=B)


def _same(a, b, cost_exact=True):
    for k in ("idx", "s_seq", "reached_t", "crash", "min_dist", "start_s"):
        assert torch.equal(a[k], b[k]), k
    if cost_exact:
        assert torch.equal(a["cost"], b["cost"])
    else:
        assert bool(((a["cost"] - b["cost"]).abs() <= 1e-7 * b["cost"].clamp(min=1.0)).all())


@pytest.mark.gpu
@pytest.mark.parametrize("H,traffic,kind", [(17, "moderate", "mixed"), (50, "moderate", "onramp"), (25, "fast", "onramp")])
def test_hint_never_changes_the_plan(H, traffic, kind):
    B = 512
    eng = _engine(H, B)
    S = E.states_to_device(synthetic.make_states(B, traffic, seed=3, kind=kind), eng.device)
    args = (S["ego"], S["cars_x"], S["cars_v"], S["cars_a"], S["n_cars"])
    ref = {k: v.clone() for k, v in eng.plan(*args).items()}
    ce = eng.fast32_info()["frac_bits"] == 18
    g = torch.Generator(device="cpu").manual_seed(0)
    for scale in (0.25, 0.9, 1.0, 1.05, 1.5, 1e4):
        _same(eng.plan_hinted(*args, hint_cost=ref["cost"].clone(), hint_scale=scale), ref, ce)
    # garbage hints: zeros, negatives, NaN, inf, random
    junk = torch.rand(B, generator=g, dtype=torch.float64).to(eng.device) * 3e4
    junk[::7] = float("nan"); junk[1::7] = -1.0; junk[2::7] = 0.0; junk[3::7] = float("inf")
    _same(eng.plan_hinted(*args, hint_cost=junk), ref, ce)
    # the validity mask: only episodes whose hint plan was complete use their hint
    reached = torch.where(torch.arange(B, device=eng.device) % 2 == 0, H, H - 1).to(torch.int32)
    _same(eng.plan_hinted(*args, hint_cost=ref["cost"] * 0.5, hint_reached=reached, hint_full_t=H), ref, ce)
    # exact mode ignores hints
    ex = {k: v.clone() for k, v in eng.plan(*args, mode="exact").items()}
    _same(eng.plan_hinted(*args, hint_cost=junk, mode="exact"), ex)


@pytest.mark.gpu
@pytest.mark.parametrize("mult", [(20, 3), (8, 2)])
def test_probed_plan_equals_plain_plan(mult):
    B = 1024
    eng = _engine(50, B)
    probe = eng.make_probe(*mult)
    assert probe.num_t < eng.num_t and probe.num_s_max < eng.num_s_max
    S_np = synthetic.make_states(B, "moderate", seed=0)
    S = E.states_to_device(S_np, eng.device)
    args = (S["ego"], S["cars_x"], S["cars_v"], S["cars_a"], S["n_cars"])
    ref = {k: v.clone() for k, v in eng.plan(*args).items()}
    _same(eng.plan_probed(probe, *args, margin=1.1), ref, False)
    _same(eng.plan_probed(probe, *args, margin=0.5), ref, False)   # every first attempt too low: the ladder recovers
    if mult == (20, 3):                                            # the host-buffer form of the same call
        host = eng.plan_host(S_np["ego"], S_np["cars_x"], S_np["cars_v"], S_np["cars_a"], S_np["n_cars"], probe=probe)
        for k in ("idx", "s_seq", "reached_t", "crash", "min_dist", "start_s"):
            assert np.array_equal(host[k], ref[k].cpu().numpy()), k
        assert np.all(np.abs(host["cost"] - ref["cost"].cpu().numpy()) <= 1e-7 * np.maximum(ref["cost"].cpu().numpy(), 1.0))
    # the probe plan is a usable estimate (CPU model: median 1.00, IQR 0.97-1.04 for 20x3)
    pr = probe.plan(*args)
    ok = (pr["reached_t"] == probe.num_t - 1) & (ref["reached_t"] == eng.num_t - 1) & (ref["cost"] < 1e6)
    ratio = (pr["cost"] * (eng.num_t - 1) / (probe.num_t - 1) / ref["cost"])[ok].cpu().numpy()
    assert 0.85 < np.median(ratio) < 1.15
