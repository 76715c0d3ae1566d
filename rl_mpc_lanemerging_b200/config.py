"""`Settings`: the global, mutable, class-attribute registry the reference reads everywhere
(reference config.py:7-170), restricted to what the MPC hot path and its callers consume.

`Settings.load_from_file(path)` accepts the reference's configs/*.json unchanged: every key becomes
a class attribute (unknown keys are kept, dict values get int keys like the reference's loader,
config.py:161-170).  The CUDA library snapshots the scalars at engine creation; after mutating
Settings call `st.refresh_engine()` (or just let the shims notice the changed snapshot key).
"""
from __future__ import annotations

import json

_HOT_PATH_DEFAULTS = {
    # task / bookkeeping
    "TASK": "ST", "NUM_EPISODES": 2000, "LOG_DIR": "last_run", "MODEL_NAME": "", "SEED": "Random",
    # simulation limits
    "TICK_LENGTH": 0.2, "MAX_POSITIVE_ACCELERATION": 4.5, "MAX_NEGATIVE_ACCELERATION": -6.0,
    "MINIMUM_NEGATIVE_JERK": -5.0, "MAXIMUM_POSITIVE_JERK": 5.0, "MAX_SPEED": 30, "CAR_LENGTH": 5.0,
    "MERGE_POINT_X": -50, "MAX_EPISODE_LENGTH": 100,
    # synthetic traffic (reference spawner parameters)
    "BASE_TRAFFIC_INTERVAL": 1.2, "OTHER_CAR_SPEED": 7.0, "VARY_TRAFFIC_START_TIMES": True,
    "START_SPEED": 15, "RANDOMIZE_START_SPEED": True, "START_SPEED_VARIANCE": 5, "MIN_START_SPEED": 5, "MAX_START_SPEED": 25,
    # sensing / observation vector
    "SENSOR_RADIUS": 125, "CARS_AHEAD": 2, "CARS_BEHIND": 2, "USE_ACCELERATION_OF_OTHER_CARS": True,
    "USE_SPEED_DIFFERENCE": True, "NORMALIZE_VECTOR_INPUT": True,
    # reward (every published config: "Slotted Jerk")
    "REWARD_FUNCTION": "Slotted Jerk", "CRASH_REWARD": -10, "SUCCESS_REWARD": 10, "TIME_REWARD": -0.1, "ALT_J_WEIGHT": 0.05,
    "INVALID_ACTION_PENALTY": 0.0,
    # S-T planner
    "DESIRED_SPEED": 30.0, "USE_CYTHON": True, "USE_FAST_ST_SOLVER": True, "S_DISCRETIZATION": 0.05, "T_DISCRETIZATION": 0.30,
    "FUTURE_S": 150.0, "FUTURE_T": 5.0, "START_UNCERTAINTY": 0.0, "UNCERTAINTY_PER_SECOND": 0.0,
    "V_WEIGHT": 0.5, "A_WEIGHT": 10.0, "J_WEIGHT": 10.0, "D_WEIGHT": 10.0, "MIN_ALLOWED_DISTANCE": 5, "CRASH_MIN_S": 12,
    # predictor
    "MAX_PREDICTED_DECELERATION": -4,
    # RL + MPC combination
    "ROLLOUT_LENGTH": 5, "ST_TEST_ROLLOUTS": 5, "LIMIT_DQN_SPEED": False, "TEST_ST_STRICTLY_BETTER": True,
    "TEST_ROLLOUT_STATE": True, "CHECK_ROLLOUT_CRASH": True, "COMBINATION_MIN_DISTANCE": 5.1, "STOP_X": 65,
    "REMEMBER_LAST_CHOICE_FOR_SWITCHING_COMBINED": False, "LEARNING_RATE": 2e-4,
    # this implementation only
    # training (reference ddpg.py:46-117).  The reference's preset makes one gradient step on a minibatch of 100 per environment frame
    # (update_frequency 1): 100 replayed samples per frame.  A tick here yields TRAIN_NUM_ENVS frames per rank, so it is followed by
    # TRAIN_NUM_ENVS * TRAIN_SAMPLES_PER_FRAME / TRAIN_MINIBATCH gradient steps (TRAIN_UPDATES_PER_TICK = None), which keeps that
    # update-to-data ratio with larger minibatches; an integer overrides it (throughput runs).
    "TRAIN_NUM_ENVS": 256, "TRAIN_UPDATES_PER_TICK": None, "TRAIN_MINIBATCH": 1024, "TRAIN_SAMPLES_PER_FRAME": 100, "EVAL_NUM_ENVS": 4096,
    "ST_MODE": "exact",       # arithmetic of the single-state drop-in calls: "exact" (fp64, st_cy-identical) or "fast"
    "CUDA_DEVICE": 0,
    # closed loop: bound each episode's plan by PLAN_HINT_SCALE x the cost of its previous plan (mpc_plan_hinted; the plans are
    # identical either way, DESIGN.md §3 "Cost hints"; hinted solves run on the 64-bit kernel).  Off by default.
    "PLAN_COST_HINTS": False, "PLAN_HINT_SCALE": 1.15,
    # MergeEnv.step as ONE kernel (mpc_env_step) instead of ~150 tensor operations; same random numbers, same trajectories
    # (bit-identical to the tensor version on the device and under the CPU emulation, both world models).  Off by default.
    "FUSED_ENV_STEP": False,
    # combined controller: hand the vetoed episodes to the planner through mpc_plan_masked / mpc_finer_fit_masked (episode list
    # built and counted on the device) instead of nonzero() + gather on the host side: no host sync in the tick.  Same speeds.
    "SYNC_FREE_TAKEOVER": False,
    # dynamics of merge_gym.MergeEnv: "predictor" = the reference's own traffic predictor (prediction.py:46-105) applied as the world,
    # "krauss" = SUMO's Krauss car-following model with the vType of merge_impossible.rou.xml:3 (mpc_krauss_step)
    "WORLD_MODEL": "krauss",
}


class _SettingsMeta(type):
    def __repr__(cls):
        return f"Settings({cls.export_settings()})"


class Settings(metaclass=_SettingsMeta):
    @classmethod
    def export_settings(cls):
        return {k: v for k, v in vars(cls).items() if k.isupper()}

    @classmethod
    def load_from_file(cls, filename):
        with open(filename, "rb") as f:
            contents = json.load(f)
        for key, value in contents.items():
            if isinstance(value, dict):
                value = {int(k): v for k, v in value.items()}
            setattr(cls, key, value)

    @classmethod
    def setup_logging(cls):
        """runs/<LOG_DIR>/ with settings.json and out.log (reference config.py:179-193; the source snapshot is not copied)."""
        import logging
        import os
        logdir = os.path.join(getattr(cls, "RUNS_DIR", "runs"), str(cls.LOG_DIR))
        os.makedirs(logdir, exist_ok=True)
        with open(os.path.join(logdir, "settings.json"), "w") as f:
            json.dump({k: v for k, v in cls.export_settings().items() if isinstance(v, (str, int, float, bool, list, dict))},
                      f, indent=4, sort_keys=True, default=str)
        for h in list(logging.getLogger().handlers):
            if isinstance(h, logging.FileHandler):
                logging.getLogger().removeHandler(h)
        logging.getLogger().addHandler(logging.FileHandler(os.path.join(logdir, "out.log")))
        logging.getLogger().setLevel(logging.INFO)
        cls.FULL_LOG_DIR = logdir
        return logdir

    @classmethod
    def seed_value(cls) -> int:
        """The run's seed as an integer: Settings.SEED, or fresh entropy when it is "Random" (reference main.py:93-97 seeds numpy /
        torch / random from it; here the simulated world and the learner are seeded explicitly).  The *_1 / *_2 / *_3 configs of the
        reference differ in nothing but SEED = 100 / 101 / 102."""
        seed = getattr(cls, "SEED", "Random")
        if isinstance(seed, str):
            import os
            return int.from_bytes(os.urandom(4), "little")
        return int(seed)

    @classmethod
    def reset(cls):
        for key in [k for k in vars(cls) if k.isupper()]:
            delattr(cls, key)
        for key, value in _HOT_PATH_DEFAULTS.items():
            setattr(cls, key, value)


Settings.reset()
