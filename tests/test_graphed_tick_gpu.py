"""The CUDA-graph tick (rl_mpc_lanemerging_b200/graphed_tick.py) against the same tick launched kernel by kernel: bit-identical
worlds, rewards, flags and take-overs over 120 ticks, both world models.  (The eager tick with the two switches on is itself
compared with the plain tick in tests/test_fused_env_gpu.py and tests/test_masked_gpu.py.)"""
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(torch, graphed, world, n_envs=768, warm=32, ticks=120):
    from rl_mpc_lanemerging_b200 import ddpg, merge_gym, st
    from rl_mpc_lanemerging_b200.config import Settings
    from rl_mpc_lanemerging_b200.graphed_tick import GraphedTick
    Settings.reset()
    Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.BASE_TRAFFIC_INTERVAL = 20, 11.0, 1.2      # combined_moderate_1.json
    Settings.TEST_ST_STRICTLY_BETTER, Settings.ALT_J_WEIGHT, Settings.WORLD_MODEL = False, 0.1, world
    Settings.FUSED_ENV_STEP = Settings.SYNC_FREE_TAKEOVER = True
    st.refresh_engine()
    try:
        env = merge_gym.MergeEnv(n_envs, seed=21)
        agent = ddpg.DDPGAgent.load_npz(os.path.join(ROOT, "tests", "golden", "policy_moderate1.npz"), device="cuda:0")
        env.reset()
        gt = GraphedTick(env, agent)
        if graphed:
            gt.capture(warmup_ticks=warm)
        else:
            gt.eager(warm)
        hist = []
        for _ in range(ticks):
            gt.replay() if graphed else gt.eager()
            hist.append(torch.cat([gt.reward, gt.done.double(), gt.takeover.double(), gt.crashed.double(), gt.merged.double()]).clone())
        torch.cuda.synchronize()
        S = env.state
        final = [t.clone() for t in (S.ego, S.cars_x, S.cars_v, S.cars_a, S.n_cars, env.ticks, env.prev_acc, env.delay, agent.timestep)]
        return torch.stack(hist).cpu(), [t.cpu() for t in final]
    finally:
        Settings.reset()
        st.refresh_engine()


@pytest.mark.parametrize("world", ["krauss", "predictor"])
def test_graph_replay_equals_the_eager_tick(world):
    import torch
    h_e, f_e = _run(torch, False, world)
    h_g, f_g = _run(torch, True, world)
    assert torch.equal(h_e, h_g), f"first differing tick: {int((h_e != h_g).any(1).nonzero()[0])}"
    for a, b in zip(f_e, f_g):
        assert torch.equal(a, b)
    done, take = h_e[:, 768:1536], h_e[:, 1536:2304]
    assert done.sum() > 50 and take.sum() > 0 and bool(torch.isfinite(h_e).all())       # episodes ended and the planner took over
