"""The closed loops around the hot path on the EMULATED library (small sizes): TASK ST evaluation with its run_data.csv
row, the combined controller through the task dispatcher, a few DDPG training ticks with save / load.  Same code paths as
tests/test_train_eval_gpu.py (which stays the device gate), sized for the CPU suite."""
import os

import numpy as np
import pytest


@pytest.fixture
def settings(emulated_engine):
    from rl_mpc_lanemerging_b200.config import Settings
    Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.BASE_TRAFFIC_INTERVAL = 20, 11.0, 1.2     # the "moderate" configs
    Settings.ST_MODE = "fast"
    return Settings


def test_st_and_combined_evaluation_loops(settings, tmp_path):
    """control.evaluate_control (the reference's episode loop, control.py:343-363, batched) under the ST planner and under
    the RL-proposes / MPC-vetoes controller; episodes cut at 3 s / 2 s so that the emulation finishes in seconds."""
    from rl_mpc_lanemerging_b200 import control, ddpg, st
    out = control.evaluate_control(st.do_st_control, num_episodes=6, max_episode_length=3, num_envs=8, seed=1)
    avg = out.get_stat_averages()
    assert out.episodes >= 6 and avg["crashed"] == 0.0 and avg["mean_speed"] > 5.0 and np.isfinite(avg["mean_abs_jerk"])
    assert avg["time_taken"] <= 3.0 + 0.2 + 1e-9 and avg["clock_time_per_step"] > 0
    out.print_stats = getattr(out, "print_stats", None)
    settings.TEST_ST_STRICTLY_BETTER = False
    agent = ddpg.DDPGAgent(device="cpu", seed=1)
    out2 = control.evaluate_control(agent.do_combined_control, num_episodes=4, max_episode_length=2, num_envs=8, seed=2,
                                    end_episode_callback=agent.reset_time)
    assert out2.episodes >= 4 and np.isfinite(out2.get_stat_averages()["mean_speed"])
    assert len(agent.takeover_history) > 0


def test_ddpg_training_ticks_save_load(settings, tmp_path):
    import torch
    from rl_mpc_lanemerging_b200 import ddpg, merge_gym, trainer
    env = merge_gym.MergeEnv(16, seed=3)
    tr = trainer.DDPGTrainer(env, lr=2e-4, seed=0, minibatch_size=32, replay_start_size=48, replay_buffer_size=1024)
    w0 = torch.nn.utils.parameters_to_vector(tr.policy.parameters()).detach().clone()
    tr.train(16 * 8)
    assert tr.frames == 16 * 8 and tr.grad_steps >= 3
    assert torch.isfinite(tr.last["q_loss"]) and torch.isfinite(tr.last["pi_loss"])
    assert float((torch.nn.utils.parameters_to_vector(tr.policy.parameters()).detach() - w0).abs().max()) > 0
    tr.save(str(tmp_path))
    agent = ddpg.DDPGAgent.load(str(tmp_path), device="cpu")
    x = torch.rand(7, 21)
    assert torch.equal(agent.policy(x), tr.policy(x))


def test_plan_cost_hints_do_not_change_the_closed_loop(settings):
    """Settings.PLAN_COST_HINTS: every tick's plans are bounded by the previous tick's costs (mpc_plan_hinted).  The closed
    loop -- commanded speeds, take-over decisions, episode statistics -- must be the same with and without."""
    import torch
    from rl_mpc_lanemerging_b200 import ddpg, merge_gym, st
    settings.TEST_ST_STRICTLY_BETTER = False
    runs = {}
    for hints in (False, True):
        settings.PLAN_COST_HINTS = hints
        env = merge_gym.MergeEnv(8, seed=5)
        env.reset()
        agent = ddpg.DDPGAgent(device="cpu", seed=2)
        hint = st.PlanHint()
        log = []
        for tick in range(12):
            if tick % 2 == 0:                                              # alternate the two closed-loop callers
                speed = st.do_st_control(env.state, hint=hint)
            else:
                speed, take = agent.do_combined_control(env.state)
                log.append(take.clone())
            jerk = ((speed - env.state.ego[:, 2]) / settings.TICK_LENGTH - env.state.ego[:, 3]) / settings.TICK_LENGTH
            _obs, _r, done, _info = env.step(jerk)
            agent.reset_time(done)
            log.append(speed.clone())
        assert (hint.cost is not None) == hints
        runs[hints] = log
    for a, b in zip(runs[False], runs[True]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("auto_reset,randomize,world", [(True, True, "predictor"), (True, False, "krauss"), (False, True, "krauss"),
                                                        (True, True, "krauss")])
def test_fused_env_step_equals_the_tensor_version(settings, auto_reset, randomize, world):
    """Settings.FUSED_ENV_STEP: MergeEnv.step as one kernel (mpc_env_step).  Two environments with the same seed, one stepped
    by the kernel, one by the ~150 tensor operations, fed the same jerks: states, rewards, flags and observations must stay
    bit-identical over resets, recycled and entering cars, clipped actions, crashes and timeouts."""
    import torch
    from rl_mpc_lanemerging_b200 import merge_gym
    settings.MAX_EPISODE_LENGTH = 2.4                            # 12 ticks: timeouts (and their resets) inside the run
    settings.RANDOMIZE_START_SPEED = randomize
    settings.WORLD_MODEL = world                                 # both dynamics: the reference's predictor, SUMO's Krauss model
    envs = {}
    for fused in (False, True):
        e = merge_gym.MergeEnv(12, seed=9, auto_reset=auto_reset)
        e.reset()
        envs[fused] = e
    g = torch.Generator().manual_seed(4)
    seen = dict(crashed=0, timeout=0, clipped=0, spawned=0, recycled=0)
    for tick in range(45):
        jerk = (torch.rand(12, generator=g, dtype=torch.float64) - 0.45) * 14.0
        if tick > 30:
            jerk = jerk.abs() + 3.0                              # floor it: run into the traffic ahead
        n_before = envs[True].state.n_cars.clone()
        outs = {}
        for fused in (False, True):
            settings.FUSED_ENV_STEP = fused
            outs[fused] = envs[fused].step(jerk.clone())
        settings.FUSED_ENV_STEP = False
        (o0, r0, d0, i0), (o1, r1, d1, i1) = outs[False], outs[True]
        assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(d0, d1), tick
        for k in ("crashed", "merged", "timeout", "projected_jerk"):
            assert torch.equal(i0[k], i1[k]), (tick, k)
        a, b = envs[False], envs[True]
        for x, y in zip(a.state.args(), b.state.args()):
            assert torch.equal(x, y), tick
        assert torch.equal(a.prev_acc, b.prev_acc) and torch.equal(a.delay, b.delay) and torch.equal(a.ticks, b.ticks), tick
        seen["crashed"] += int(i1["crashed"].sum()); seen["timeout"] += int(i1["timeout"].sum())
        seen["clipped"] += int((i1["projected_jerk"] != jerk).sum())
        seen["spawned"] += int((b.state.n_cars > n_before).sum()); seen["recycled"] += int((b.state.n_cars < n_before).sum())
    assert seen["timeout"] > 0 and seen["clipped"] > 0 and seen["spawned"] > 0
    if auto_reset:
        assert seen["crashed"] > 0 or seen["timeout"] > 12


def test_sync_free_takeover_gives_the_same_decisions(settings):
    """Settings.SYNC_FREE_TAKEOVER: the vetoed episodes go to the planner through the masked entry points (no nonzero(), no
    gather, no host sync).  Speeds and take-over masks of the combined controller must not change."""
    import torch
    from rl_mpc_lanemerging_b200 import ddpg, synthetic
    from rl_mpc_lanemerging_b200.prediction import BatchedState
    settings.TEST_ST_STRICTLY_BETTER = False
    S = synthetic.make_states(32, "moderate", seed=17, kind="mixed")
    runs = {}
    for flag in (False, True):
        settings.SYNC_FREE_TAKEOVER = flag
        agent = ddpg.DDPGAgent(device="cpu", seed=3)
        out = []
        for rep in range(2):                                     # twice: the second call re-uses the scratch tensors
            speed, take = agent.do_combined_control(BatchedState.from_numpy(S, "cpu"))
            out += [speed.clone(), take.clone()]
        runs[flag] = out
        assert 0 < int(out[1].sum()) < 32                        # some episodes were taken over, some not
    for a, b in zip(runs[False], runs[True]):
        assert torch.equal(a, b)
