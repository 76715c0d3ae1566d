"""mpc_env_step (Settings.FUSED_ENV_STEP) on the device against the tensor version of MergeEnv.step.

Bit-identical under the CPU emulation too (tests/test_closed_loop_emulated_cpu.py).  Its first device run (round 2) failed on the
reward: `tensor / python_float` on a CUDA device multiplies by the rounded reciprocal, the kernel -- like the reference's float
arithmetic (merge_gym.py:83-96) -- divides; the tensor version now divides as well (prediction.tdiv).  On a GPU the fresh-episode traffic of the tensor version goes through torch.cumsum (a parallel scan),
the kernel sums left to right: rows that were reset agree to rounding, everything else bit for bit."""
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("world", ["predictor", "krauss"])
@pytest.mark.parametrize("auto_reset", [False, True])
def test_fused_env_step_matches_tensor_version(auto_reset, world):
    import torch
    from rl_mpc_lanemerging_b200 import merge_gym, st
    from rl_mpc_lanemerging_b200.config import Settings
    Settings.reset()
    Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.BASE_TRAFFIC_INTERVAL = 20, 11.0, 1.2
    Settings.MAX_EPISODE_LENGTH = 3.0
    Settings.WORLD_MODEL = world
    st.refresh_engine()
    try:
        B = 256
        envs = {}
        for fused in (False, True):
            e = merge_gym.MergeEnv(B, seed=9, auto_reset=auto_reset)
            e.reset()
            envs[fused] = e
        g = torch.Generator(device="cuda:0").manual_seed(4)
        for tick in range(40):
            jerk = (torch.rand(B, generator=g, dtype=torch.float64, device="cuda:0") - 0.45) * 14.0
            outs = {}
            for fused in (False, True):
                Settings.FUSED_ENV_STEP = fused
                outs[fused] = envs[fused].step(jerk.clone())
            Settings.FUSED_ENV_STEP = False
            (o0, r0, d0, i0), (o1, r1, d1, i1) = outs[False], outs[True]
            assert torch.equal(d0, d1) and torch.equal(i0["crashed"], i1["crashed"]) and torch.equal(i0["timeout"], i1["timeout"]), tick
            assert torch.equal(r0, r1) and torch.equal(i0["projected_jerk"], i1["projected_jerk"]), tick
            a, b = envs[False], envs[True]
            assert torch.equal(a.state.n_cars, b.state.n_cars) and torch.equal(a.ticks, b.ticks), tick
            for x, y in zip(a.state.args()[:4], b.state.args()[:4]):
                if auto_reset:
                    assert torch.allclose(x, y, rtol=1e-12, atol=1e-9), tick
                else:
                    assert torch.equal(x, y), tick
            assert torch.allclose(o0, o1, rtol=0, atol=1e-6), tick
    finally:
        Settings.reset()
        st.refresh_engine()
