"""Task dispatch with the reference's command line (reference main.py:16-40, 84-102):

    python -m rl_mpc_lanemerging_b200.main configs/st_moderate.json

The reference's configs/*.json load unchanged.  SUMO is replaced by the batched world of merge_gym.MergeEnv;
Rainbow / tabular tasks are out of scope (no published config uses them).
"""
from __future__ import annotations

import argparse
import random

import numpy as np
import torch

from . import ddpg, st
from .config import Settings


def do_task():
    task = Settings.TASK
    if task == "ST":
        return st.evaluate_st_and_dump_crash(Settings.NUM_EPISODES)
    if task == "TRAIN_DDPG":
        return ddpg.train_ddpg_all_with_lr_drop(int(getattr(Settings, "TRAIN_FRAMES", 1e6)))
    if task == "RESUME_DDPG":
        return ddpg.DDPGAgent.resume_training(Settings.MODEL_NAME, int(getattr(Settings, "TRAIN_FRAMES", 1e6)))
    if task == "EVALUATE_DDPG":
        return ddpg.DDPGAgent.load(Settings.MODEL_NAME).evaluate(Settings.NUM_EPISODES)
    if task in ("EVALUATE_COMBINED_DDPG", "EVALUATE_COMBINED_DQN"):
        return ddpg.DDPGAgent.load(Settings.MODEL_NAME).evaluate_combined(Settings.NUM_EPISODES)
    raise NotImplementedError(f"TASK {task!r}: Rainbow / tabular trainers are out of scope (reference main.py:21-31)")


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("config", nargs="?", default=None)
    args = parser.parse_args(argv)
    if args.config is not None:
        Settings.load_from_file(args.config)
    Settings.setup_logging()
    if Settings.SEED != "Random":
        np.random.seed(Settings.SEED); torch.manual_seed(Settings.SEED); random.seed(Settings.SEED)
    # (the batched world and the learner take their seed from Settings.seed_value(): control.evaluate_control, ddpg.DDPGAgent._trainer)
    st.refresh_engine()
    return do_task()


if __name__ == "__main__":
    main()
