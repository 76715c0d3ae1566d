"""B200-native S-T MPC hot path for RL-MPC lane merging (drop-in for jlubars/RL-MPC-LaneMerging's planner path).

Sub-modules mirror the reference's flat module names: config (Settings), prediction (HighwayState), control,
st, st_cy, dqn, ddpg, merge_gym.  The computation lives in libmpcb200.so (hand-written CUDA for sm_100a,
C ABI in include/mpcb200.h); importing this package does not need a GPU, calling the planner does.
"""
__all__ = ["config", "prediction", "control", "st", "st_cy", "dqn", "ddpg", "merge_gym", "engine", "synthetic"]
__version__ = "0.1.0"
