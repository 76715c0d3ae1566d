"""Closed-loop statistics of the pure MPC controller in the batched world (merge_gym.MergeEnv) next to the reference's
published SUMO row (experiment_data/saved_data.csv:48, st_moderate).  The world model is the reference's own predictor used
as dynamics -- parity against SUMO is UNPINNED; this script only shows how far the aggregate statistics are."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_mpc_lanemerging_b200 import st
from rl_mpc_lanemerging_b200.config import Settings

PUBLISHED = {"crashed": 0.0, "merged": 1.0, "mean_speed": 13.2447, "max_speed": 23.5486, "mean_abs_jerk": 1.2796,
             "closest_distance": 11.0500, "mean_closest_distance": 12.8850, "mean_disruption": 0.1885, "max_disruption": 4.7225,
             "time_to_merge": 20.34}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Settings.reset()
for k, v in dict(TASK="ST", BASE_TRAFFIC_INTERVAL=1.2, OTHER_CAR_SPEED=11.0, ALT_J_WEIGHT=0.1, CRASH_MIN_S=20, ST_MODE="fast").items():
    setattr(Settings, k, v)                                  # configs/st_moderate.json
st.refresh_engine()
t0 = time.time()
out = st.evaluate_st_and_dump_crash(n, num_envs=min(n, 2048), csv_path=os.path.join("gpurun_out", "closed_loop_run_data.csv"))
avg = out.get_stat_averages()
res = {"episodes": out.episodes, "wall_s": time.time() - t0, "ours": {k: avg[k] for k in PUBLISHED}, "published_sumo": PUBLISHED}
print(json.dumps(res, indent=1))
