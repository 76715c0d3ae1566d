"""Closed-loop statistics in the batched SUMO-free world (merge_gym.MergeEnv) next to the reference's published SUMO rows
(experiment_data/saved_data.csv:48 st_moderate, :46 combined_moderate_1), for both world models:

  predictor   the reference's own traffic predictor (prediction.py:46-105) applied as dynamics
  krauss      SUMO's Krauss car-following model with the vType of merge_impossible.rou.xml:3 (mpc_krauss_step)

    python tools/closed_loop_stats.py [episodes] > profiles/r02_closed_loop.json

Parity against SUMO itself cannot be pinned here (no SUMO): this shows how far the aggregate statistics of whole episodes are."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_mpc_lanemerging_b200 import ddpg, st
from rl_mpc_lanemerging_b200.config import Settings

ST_ROW = {"crashed": 0.0, "merged": 1.0, "mean_speed": 13.2447, "max_speed": 23.5486, "mean_abs_jerk": 1.2796,
          "closest_distance": 11.0500, "mean_closest_distance": 12.8850, "mean_disruption": 0.1885, "max_disruption": 4.7225,
          "time_to_merge": 20.34}
COMBINED_ROW = {"crashed": 0.0, "merged": 1.0, "mean_abs_jerk": 0.69, "time_to_merge": 20.25, "percent st solver": 0.037}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
res = {"episodes": n, "published_sumo": {"st_moderate (saved_data.csv:48)": ST_ROW, "combined_moderate_1 (saved_data.csv:46)": COMBINED_ROW}}
for world in ("predictor", "krauss"):
    Settings.reset()
    for k, v in dict(TASK="ST", BASE_TRAFFIC_INTERVAL=1.2, OTHER_CAR_SPEED=11.0, ALT_J_WEIGHT=0.1, CRASH_MIN_S=20, ST_MODE="fast",
                     WORLD_MODEL=world, SEED=100).items():
        setattr(Settings, k, v)                                  # configs/st_moderate.json
    st.refresh_engine()
    t0 = time.time()
    out = st.evaluate_st_and_dump_crash(n, num_envs=min(n, 2048), csv_path=os.path.join("gpurun_out", f"closed_loop_st_{world}.csv"))
    avg = out.get_stat_averages()
    res[f"st_moderate/{world}"] = dict({k: avg[k] for k in ST_ROW}, wall_s=time.time() - t0, episodes=out.episodes)
    for k, v in dict(TASK="EVALUATE_COMBINED_DDPG", ROLLOUT_LENGTH=5, ST_TEST_ROLLOUTS=5, LIMIT_DQN_SPEED=False, TEST_ST_STRICTLY_BETTER=False,
                     TEST_ROLLOUT_STATE=True, CHECK_ROLLOUT_CRASH=True, COMBINATION_MIN_DISTANCE=5.1, STOP_X=65).items():
        setattr(Settings, k, v)                                  # configs/combined_moderate_1.json
    agent = ddpg.DDPGAgent.load_npz(os.path.join(ROOT, "tests", "golden", "policy_moderate1.npz"))
    t0 = time.time()
    out = agent.evaluate_combined(n, num_envs=min(n, 2048), csv_path=os.path.join("gpurun_out", f"closed_loop_combined_{world}.csv"))
    avg = out.get_stat_averages()
    res[f"combined_moderate_1/{world}"] = dict({k: avg.get(k) for k in COMBINED_ROW}, mean_speed=avg["mean_speed"], wall_s=time.time() - t0,
                                                episodes=out.episodes, policy="published actor (tests/golden/policy_moderate1.npz)")
print(json.dumps(res, indent=1))
