"""DDPG actor for action proposal (reference ddpg.py:24-93).  The reference wraps the third-party
`autonomous-learning-library` preset; its deployed computation is a 21->400->300->1 MLP followed by
tanh * 5 on the 20-dim observation plus a time feature.  That forward stays in PyTorch (north_star), on
device tensors produced by the K4 observation kernel -- no host round trip.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Optional

import torch
import torch.nn as nn

from . import dqn, st
from .config import Settings
from .prediction import BatchedState, HighwayState

TIME_FEATURE_SCALE = 0.001          # autonomous-learning-library 0.5.3 TimeFeature default (unpinned)


class PolicyNet(nn.Module):
    """Same layout / state_dict keys as the reference checkpoints (`model.{0,2,4}.{weight,bias}`)."""

    def __init__(self, obs_dim=21, tanh_scale=5.0, tanh_mean=0.0):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(obs_dim, 400), nn.ReLU(), nn.Linear(400, 300), nn.ReLU(), nn.Linear(300, 1))
        self.tanh_scale, self.tanh_mean = tanh_scale, tanh_mean

    def forward(self, obs):
        return torch.tanh(self.model(obs)).squeeze(-1) * self.tanh_scale + self.tanh_mean


def _load_legacy_state_dict(path):
    """The published checkpoints are whole-module pickles of `all` classes; unpickle them against stand-ins."""
    names = {"all": {}, "all.policies": {}, "all.policies.deterministic": {"DeterministicPolicyNetwork": nn.Module},
             "all.approximation": {}, "all.approximation.q_continuous": {"QContinuousModule": nn.Module},
             "all.nn": {"Linear0": nn.Linear, "RLNetwork": nn.Module}}
    added = []
    for mod, attrs in names.items():
        if mod not in sys.modules:
            m = types.ModuleType(mod)
            for k, base in attrs.items():
                setattr(m, k, type(k, (base,), {}))
            sys.modules[mod] = m
            added.append(mod)
    try:
        obj = torch.load(path, map_location="cpu", weights_only=False)
        return obj.state_dict() if hasattr(obj, "state_dict") else obj
    finally:
        for mod in added:
            sys.modules.pop(mod, None)


class DDPGAgent(dqn.RLAgent):
    def __init__(self, device=None, seed: int = 0):
        super().__init__()
        self.device = torch.device(device if device is not None else f"cuda:{int(getattr(Settings, 'CUDA_DEVICE', 0))}")
        g = torch.Generator().manual_seed(seed)
        self.policy = PolicyNet()
        for p in self.policy.parameters():                     # deterministic random init (no checkpoint in the container)
            with torch.no_grad():
                p.copy_(torch.empty_like(p).uniform_(-0.05, 0.05, generator=g))
        self.policy.to(self.device).eval()
        self.timestep: Optional[torch.Tensor] = None
        self._obs: Optional[torch.Tensor] = None

    @classmethod
    def load(cls, path, device=None) -> "DDPGAgent":
        """`path` is a run directory holding policy.pt (reference ddpg.py:37-44)."""
        agent = cls(device)
        f = os.path.join(path, "policy.pt")
        sd = _load_legacy_state_dict(f)
        agent.policy.load_state_dict({k: v for k, v in sd.items() if k.startswith("model.")})
        agent.policy.to(agent.device).eval()
        return agent

    @classmethod
    def load_npz(cls, path, device=None) -> "DDPGAgent":
        """The same actor from a plain .npz of its tensors (keys model_{0,2,4}_{weight,bias}, tanh_scale, tanh_mean) -- the form in
        which the published pretrained_models/ddpg_moderate1_extended/policy.pt is committed as a test fixture
        (tests/golden/policy_moderate1.npz, written by tests/golden/make_golden_r2.py)."""
        import numpy as np
        z = np.load(path)
        agent = cls(device)
        sd = {f"model.{i}.{w}": torch.from_numpy(np.asarray(z[f"model_{i}_{w}"], np.float32)) for i in (0, 2, 4) for w in ("weight", "bias")}
        agent.policy.load_state_dict(sd)
        agent.policy.tanh_scale, agent.policy.tanh_mean = float(z["tanh_scale"]), float(z["tanh_mean"])
        agent.policy.to(agent.device).eval()
        return agent

    def reset_time(self, mask: Optional[torch.Tensor] = None):
        if self.timestep is not None:
            if mask is None:
                self.timestep.zero_()
            else:
                self.timestep[mask] = 0

    @torch.no_grad()
    def get_control(self, state):
        """Jerk proposed by the actor (reference ddpg.py:83-87).  The time feature advances on every call,
        like the reference's TimeFeature wrapper."""
        single = isinstance(state, HighwayState)
        eng = st.get_engine(1 if single else state.batch)
        bs = BatchedState.from_states([state], eng.device, eng.nmax) if single else state
        B = bs.batch
        if self._obs is None or self._obs.shape[0] != B:
            self._obs = torch.zeros((B, 21), dtype=torch.float32, device=self.device)
            self.timestep = torch.zeros(B, dtype=torch.float32, device=self.device)
        eng.state_vector(*bs.args(), out=self._obs)
        self._obs[:, 20] = self.timestep * TIME_FEATURE_SCALE
        self.timestep += 1
        jerk = self.policy(self._obs)
        return float(jerk.item()) if single else jerk

    def end_episode_callback(self, last_state=None):
        """Reference ddpg.py:89-90.  In the batched loops `last_state` is the mask of the episodes that just ended."""
        self.reset_time(last_state if torch.is_tensor(last_state) else None)

    # ---- training entry points with the reference's names (ddpg.py:46-81) ----
    @classmethod
    def _trainer(cls, seed=None, num_envs=None):
        from . import merge_gym, trainer
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        seed = Settings.seed_value() if seed is None else int(seed)       # (every rank must see the same Settings.SEED: "Random" is for 1 rank)
        dev = int(os.environ.get("LOCAL_RANK", getattr(Settings, "CUDA_DEVICE", 0)))
        Settings.CUDA_DEVICE = dev
        st.refresh_engine()
        n_envs, mb = int(num_envs or Settings.TRAIN_NUM_ENVS), int(Settings.TRAIN_MINIBATCH)
        upt = Settings.TRAIN_UPDATES_PER_TICK
        if upt is None:                                        # the reference's update-to-data ratio (config.py)
            upt = max(1, round(n_envs * float(Settings.TRAIN_SAMPLES_PER_FRAME) / mb))
        env = merge_gym.MergeEnv(n_envs, seed=seed + 104729 * rank)
        return trainer.DDPGTrainer(env, device=f"cuda:{dev}", lr=Settings.LEARNING_RATE, seed=seed, minibatch_size=mb, updates_per_tick=int(upt),
                                   replay_start_size=max(5000, 4 * mb))

    @classmethod
    def train(cls, num_frames: int, num_envs=None):
        """Train from scratch for `num_frames` frames per rank and leave policy.pt / q.pt in Settings.FULL_LOG_DIR."""
        tr = cls._trainer(num_envs=num_envs).train(int(num_frames))
        tr.save(Settings.FULL_LOG_DIR)
        return tr

    @classmethod
    def resume_training(cls, path, num_frames: int, num_envs=None):
        tr = cls._trainer(num_envs=num_envs).load(path)
        tr.train(int(num_frames))
        tr.save(Settings.FULL_LOG_DIR)
        return tr


def train_ddpg_all_with_lr_drop(num_frames, third=False, num_envs=None):
    """Reference ddpg.py:96-117: train, divide the learning rate by 10, train again from the first run's weights into
    <LOG_DIR>_extended; with `third`: evaluate that, divide the learning rate once more and train into <LOG_DIR>_extended2; then
    evaluate NUM_EPISODES episodes with the final weights."""
    if not hasattr(Settings, "FULL_LOG_DIR"):
        Settings.setup_logging()
    DDPGAgent.train(num_frames, num_envs)
    Settings.LEARNING_RATE /= 10
    old_log_dir = Settings.FULL_LOG_DIR
    Settings.LOG_DIR = Settings.LOG_DIR + "_extended"
    Settings.setup_logging()
    DDPGAgent.resume_training(old_log_dir, num_frames, num_envs)
    if third:                                                   # ddpg.py:104-113
        Settings.TASK = "EVALUATE_DDPG"
        Settings.MODEL_NAME = Settings.FULL_LOG_DIR
        DDPGAgent.load(Settings.FULL_LOG_DIR).evaluate(Settings.NUM_EPISODES)
        Settings.TASK = "TRAIN_DDPG"
        Settings.LEARNING_RATE /= 10
        old_log_dir = Settings.FULL_LOG_DIR
        Settings.LOG_DIR = Settings.LOG_DIR + "2"
        Settings.setup_logging()
        DDPGAgent.resume_training(old_log_dir, num_frames, num_envs)
    Settings.TASK = "EVALUATE_DDPG"
    Settings.MODEL_NAME = Settings.FULL_LOG_DIR
    eval_agent = DDPGAgent.load(Settings.FULL_LOG_DIR)
    return eval_agent.evaluate(Settings.NUM_EPISODES)
