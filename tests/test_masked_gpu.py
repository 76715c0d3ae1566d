"""mpc_plan_masked / mpc_finer_fit_masked and Settings.SYNC_FREE_TAKEOVER on the device.

Also run under the CPU emulation (tests/test_library_emulation_cpu.py, tests/test_closed_loop_emulated_cpu.py); passed on a B200 at the
end of round 1."""
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("mode", ["fast", "exact"])
def test_masked_plan_and_fit(mode):
    import torch
    from rl_mpc_lanemerging_b200 import _lib, synthetic
    from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
    B = 3000
    eng = MpcEngine(_lib.default_params(), device=0, max_batch=B)
    D = states_to_device(synthetic.make_states(B, "moderate", seed=19, kind="mixed"), "cuda:0")
    a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    ref = {k: v.clone() for k, v in eng.plan(*a, mode=mode).items()}
    fine_ref, n_ref, speed_ref, _ = eng.finer_fit(ref["s_seq"], ref["reached_t"], D["ego"])
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for frac in (0.15, 0.0, 1.0):
        mask = torch.rand(B, generator=g, device="cuda:0") < frac
        out = eng._plan_out(B)
        for k in ("idx", "s_seq", "cost", "reached_t", "crash", "min_dist"):
            out[k].fill_(77)
        eng.plan_masked(mask, *a, out=out, mode=mode)
        for k in ("idx", "s_seq", "cost", "reached_t", "crash", "min_dist"):
            assert torch.equal(out[k][mask], ref[k][mask]), k
            assert bool((out[k][~mask] == 77).all()), k
        fine, n_fine, speed = torch.full_like(fine_ref, -5.0), torch.full_like(n_ref, -5), torch.full_like(speed_ref, -5.0)
        eng.finer_fit_masked(mask, ref["s_seq"], ref["reached_t"], D["ego"], fine, n_fine, speed)
        assert torch.equal(speed[mask], speed_ref[mask]) and torch.equal(n_fine[mask], n_ref[mask])
        assert bool((speed[~mask] == -5.0).all()) and bool((n_fine[~mask] == -5).all())
    eng.close()


def test_sync_free_takeover_same_decisions():
    import torch
    from rl_mpc_lanemerging_b200 import ddpg, st, synthetic
    from rl_mpc_lanemerging_b200.config import Settings
    from rl_mpc_lanemerging_b200.prediction import BatchedState
    Settings.reset()
    Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.TEST_ST_STRICTLY_BETTER = 20, 11.0, False
    st.refresh_engine()
    try:
        S = synthetic.make_states(512, "moderate", seed=17, kind="mixed")
        runs = {}
        for flag in (False, True):
            Settings.SYNC_FREE_TAKEOVER = flag
            agent = ddpg.DDPGAgent(device="cuda:0", seed=3)
            speed, take = agent.do_combined_control(BatchedState.from_numpy(S, "cuda:0"))
            runs[flag] = (speed.clone(), take.clone())
        assert 0 < int(runs[True][1].sum()) < 512
        assert torch.equal(runs[False][0], runs[True][0]) and torch.equal(runs[False][1], runs[True][1])
    finally:
        Settings.reset()
        st.refresh_engine()
