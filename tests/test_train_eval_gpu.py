"""-m gpu: the closed loops around the hot path -- DDPG training on the batched environment, the
statistics / run_data.csv row of an evaluation (f-4), and the reference's task entry points."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def settings():
    from rl_mpc_lanemerging_b200 import st
    from rl_mpc_lanemerging_b200.config import Settings
    Settings.reset()
    Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.BASE_TRAFFIC_INTERVAL = 20, 11.0, 1.2     # the "moderate" configs
    Settings.ST_MODE = "fast"
    st.refresh_engine()
    yield Settings
    Settings.reset()
    st.refresh_engine()


def test_train_save_load_evaluate(settings, tmp_path):
    import torch
    from rl_mpc_lanemerging_b200 import ddpg, merge_gym, trainer
    env = merge_gym.MergeEnv(256, seed=3)
    tr = trainer.DDPGTrainer(env, lr=2e-4, seed=0, minibatch_size=512, replay_start_size=1024, replay_buffer_size=65536)
    w0 = torch.nn.utils.parameters_to_vector(tr.policy.parameters()).detach().clone()
    tr.train(256 * 40)
    assert tr.frames == 256 * 40 and tr.grad_steps >= 30
    assert torch.isfinite(tr.last["q_loss"]) and torch.isfinite(tr.last["pi_loss"])
    assert float((torch.nn.utils.parameters_to_vector(tr.policy.parameters()).detach() - w0).abs().max()) > 0
    r = tr.replay
    assert r.size == 256 * 40 and float(r.obs[:r.size, 20].max()) > 0 and bool((r.rew[:r.size] < 0).any())    # time feature, step cost
    tr.save(str(tmp_path))
    agent = ddpg.DDPGAgent.load(str(tmp_path))
    x = torch.rand(7, 21, device=agent.device)
    assert torch.equal(agent.policy(x), tr.policy(x))
    settings.MAX_EPISODE_LENGTH = 8
    out = agent.evaluate(24, num_envs=64, csv_path=os.path.join(tmp_path, "run_data.csv"))
    st = out.get_stat_averages()
    assert out.episodes == 24 and 0.0 <= st["crashed"] <= 1.0 and st["mean_speed"] > 0 and st["clock_time_per_step"] > 0
    import pandas as pd
    row = pd.read_csv(os.path.join(tmp_path, "run_data.csv")).iloc[0]
    assert row["TRAFFIC_DESCRIPTION"] == "uniform-11.0-1.2-varying" and "time_to_merge_std" in row.index


def test_st_and_combined_evaluation_tasks(settings, tmp_path):
    """TASK ST (st.evaluate_st_and_dump_crash) and EVALUATE_COMBINED_DDPG through main.do_task, short episodes."""
    import torch
    from rl_mpc_lanemerging_b200 import ddpg, main, st
    settings.MAX_EPISODE_LENGTH = 8
    out = st.evaluate_st_and_dump_crash(16, num_envs=32, csv_path=os.path.join(tmp_path, "st.csv"))
    avg = out.get_stat_averages()
    assert out.episodes == 16 and avg["crashed"] == 0.0                      # the planner does not crash in its own world model
    assert avg["mean_speed"] > 5.0 and np.isfinite(avg["mean_abs_jerk"])
    # a checkpoint in the reference layout, then the combined task through the dispatcher
    agent = ddpg.DDPGAgent(seed=1)
    torch.save(agent.policy.state_dict(), os.path.join(tmp_path, "policy.pt"))
    settings.TASK, settings.MODEL_NAME, settings.NUM_EPISODES, settings.EVAL_NUM_ENVS = "EVALUATE_COMBINED_DDPG", str(tmp_path), 12, 32
    settings.TEST_ST_STRICTLY_BETTER = False
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        out2 = main.do_task()
    finally:
        os.chdir(cwd)
    assert out2.episodes == 12 and "percent st solver" in out2.get_stat_averages()
    assert os.path.exists(os.path.join(tmp_path, "run_data.csv"))
