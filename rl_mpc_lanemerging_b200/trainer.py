"""Data-parallel DDPG trainer: the replacement for the `autonomous-learning-library` preset the reference trains with
(reference ddpg.py:46-117; BASELINE.json configs[3]: train_medium_1.json, environments sharded over the GPUs of a node,
NCCL all-reduce for the gradient step only).

One process per GPU.  Every rank owns `num_envs` environments (merge_gym.MergeEnv: the K4 kernels step them on the
device), a device-resident replay ring, and a replica of actor / critic / targets.  A tick = one vectorised env step
(actor forward + exploration noise) + `updates_per_tick` gradient steps; the gradients of BOTH networks live in one
flat fp32 buffer (policy 129 401 + critic 129 801 parameters = 1.04 MB) that is all-reduced with a single NCCL call per
step, so the replicas stay bit-identical (same initial weights by broadcast, same averaged gradients, local Adam and
Polyak updates).  No collective touches the environment / planner path.

Network shapes, the files of a run (`<run dir>/policy.pt`, `q.pt`), their state_dict keys (`model.{0,2,4}.{weight,bias}`), the
learning rate (Settings.LEARNING_RATE) and the two-stage schedule are the reference's.  Checkpoints travel ONE WAY: this package
reads the reference's whole-module pickles (ddpg._load_legacy_state_dict) and its own files, but writes plain state_dicts -- the
reference's loaders (`torch.load(...).to(device).state_dict()`, GreedyAgent.load) expect pickled `all.*` module objects, which
cannot be produced without that library.  The remaining DDPG
hyper-parameters live inside the third-party library, not in the reference tree: the values below are that library's
published 0.5.3 defaults as far as known -- UNPINNED.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .config import Settings
from .ddpg import TIME_FEATURE_SCALE, PolicyNet

DEFAULTS = dict(discount_factor=0.98, minibatch_size=100, replay_start_size=5000, replay_buffer_size=1_000_000,
                polyak_rate=0.005, noise=0.1, update_frequency=1)


class QNet(nn.Module):
    """Critic 22 -> 400 -> 300 -> 1 on [observation(21), action(1)] (reference checkpoints q.pt)."""

    def __init__(self, obs_dim=21):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(obs_dim + 1, 400), nn.ReLU(), nn.Linear(400, 300), nn.ReLU(), nn.Linear(300, 1))

    def forward(self, obs, action):
        return self.model(torch.cat([obs, action.unsqueeze(-1)], -1)).squeeze(-1)


class ReplayRing:
    """Uniform replay memory resident on the device: a ring of [capacity] transitions written B at a time."""

    def __init__(self, capacity: int, obs_dim: int, device):
        f = dict(dtype=torch.float32, device=device)
        self.cap, self.size, self.pos = int(capacity), 0, 0
        self.obs, self.nxt = torch.zeros((self.cap, obs_dim), **f), torch.zeros((self.cap, obs_dim), **f)
        self.act, self.rew, self.done = torch.zeros(self.cap, **f), torch.zeros(self.cap, **f), torch.zeros(self.cap, **f)
        self._ar = None

    def push(self, obs, act, rew, nxt, done):
        B = obs.shape[0]
        if self._ar is None or self._ar.numel() != B:
            self._ar = torch.arange(B, device=obs.device)
        idx = (self._ar + self.pos) % self.cap
        self.obs[idx], self.act[idx], self.rew[idx], self.nxt[idx], self.done[idx] = obs, act, rew, nxt, done
        self.pos = (self.pos + B) % self.cap
        self.size = min(self.size + B, self.cap)

    def sample(self, n: int, generator=None):
        idx = torch.randint(0, self.size, (n,), device=self.obs.device, generator=generator)
        return self.obs[idx], self.act[idx], self.rew[idx], self.nxt[idx], self.done[idx]


def _flatten_grads(modules):
    """Point every parameter's .grad at a slice of ONE flat buffer: the whole gradient is one all-reduce."""
    params = [p for m in modules for p in m.parameters()]
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    return flat


class DDPGTrainer:
    def __init__(self, env, device=None, lr: Optional[float] = None, seed: int = 0, **hyper):
        self.env = env
        self.device = torch.device(device if device is not None else env.device)
        self.h = dict(DEFAULTS); self.h.update(hyper)
        self.lr = float(Settings.LEARNING_RATE if lr is None else lr)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        torch.manual_seed(seed)                                   # same initial weights on every rank ...
        self.policy, self.q = PolicyNet().to(self.device), QNet().to(self.device)
        self.policy_target, self.q_target = PolicyNet().to(self.device), QNet().to(self.device)
        if self.world > 1:                                        # ... and broadcast anyway (one flat buffer)
            for m in (self.policy, self.q):
                vec = nn.utils.parameters_to_vector(m.parameters())
                dist.broadcast(vec, 0)
                nn.utils.vector_to_parameters(vec, m.parameters())
        self.policy_target.load_state_dict(self.policy.state_dict()); self.q_target.load_state_dict(self.q.state_dict())
        for p in list(self.policy_target.parameters()) + list(self.q_target.parameters()):
            p.requires_grad_(False)
        self.flat_grad = _flatten_grads([self.policy, self.q])
        self.n_policy = sum(p.numel() for p in self.policy.parameters())
        self.opt_pi = torch.optim.Adam(self.policy.parameters(), lr=self.lr)
        self.opt_q = torch.optim.Adam(self.q.parameters(), lr=self.lr)
        self.gen = torch.Generator(device=self.device); self.gen.manual_seed(1000 + seed + 7919 * self.rank)   # exploration / sampling differ per rank
        self.replay = ReplayRing(self.h["replay_buffer_size"], 21, self.device)
        self.frames = 0
        self.grad_steps = 0
        self.allreduce_bytes = self.flat_grad.numel() * 4
        self.timestep = None
        self.last = {}

    # ---- observation with the time feature the reference's policy consumes (ddpg.py:83-87) ----
    def _with_time(self, obs20):
        B = obs20.shape[0]
        if self.timestep is None:
            self.timestep = torch.zeros(B, dtype=torch.float32, device=self.device)
        return torch.cat([obs20.to(self.device, torch.float32), (self.timestep * TIME_FEATURE_SCALE).unsqueeze(1)], 1)

    @torch.no_grad()
    def act(self, obs21, explore=True):
        a = self.policy(obs21)
        if explore:
            a = a + self.h["noise"] * self.policy.tanh_scale * torch.randn(a.shape, device=self.device, generator=self.gen)
        return a.clamp(float(Settings.MINIMUM_NEGATIVE_JERK), float(Settings.MAXIMUM_POSITIVE_JERK))

    def update(self):
        """One DDPG gradient step on a minibatch; the gradients of both networks are averaged over the ranks in one call."""
        o, a, r, n, d = self.replay.sample(self.h["minibatch_size"], self.gen)
        with torch.no_grad():
            target = r + self.h["discount_factor"] * (1.0 - d) * self.q_target(n, self.policy_target(n))
        self.flat_grad.zero_()
        q_loss = torch.nn.functional.mse_loss(self.q(o, a), target)
        q_loss.backward(inputs=list(self.q.parameters()))
        for p in self.q.parameters():
            p.requires_grad_(False)
        pi_loss = -self.q(o, self.policy(o)).mean()
        pi_loss.backward(inputs=list(self.policy.parameters()))
        for p in self.q.parameters():
            p.requires_grad_(True)
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            self.flat_grad.div_(self.world)
        self.opt_q.step(); self.opt_pi.step()
        tau = self.h["polyak_rate"]
        with torch.no_grad():
            for tgt, src in ((self.q_target, self.q), (self.policy_target, self.policy)):
                for pt, ps in zip(tgt.parameters(), src.parameters()):
                    pt.lerp_(ps, tau)
        self.grad_steps += 1
        self.last = {"q_loss": q_loss.detach(), "pi_loss": pi_loss.detach()}

    def train(self, num_frames: int):
        """Train until this rank has consumed `num_frames` environment frames (every tick yields num_envs frames per rank)."""
        obs = self._with_time(self.env.reset())
        target = self.frames + int(num_frames)
        while self.frames < target:
            action = self.act(obs)
            nxt20, reward, done, _info = self.env.step(action)
            self.timestep = torch.where(done.to(self.device), torch.zeros_like(self.timestep), self.timestep + 1)
            nxt = self._with_time(nxt20)
            self.replay.push(obs, action.float(), reward.to(self.device, torch.float32), nxt, done.to(self.device, torch.float32))
            obs = nxt
            self.frames += obs.shape[0]
            if self.replay.size >= min(self.h["replay_start_size"], self.replay.cap):
                for _ in range(max(1, int(self.h.get("updates_per_tick", 1)))):
                    self.update()
        return self

    # ---- checkpoints: runs/<LOG_DIR>/{policy,q}.pt like the reference (ddpg.py:56-57, 62-81), as plain state_dicts (one way, see above) ----
    def save(self, path: str):
        if self.rank == 0:
            os.makedirs(path, exist_ok=True)
            torch.save({k: v.cpu() for k, v in self.policy.state_dict().items()}, os.path.join(path, "policy.pt"))
            torch.save({k: v.cpu() for k, v in self.q.state_dict().items()}, os.path.join(path, "q.pt"))
        if self.world > 1:
            dist.barrier()

    def load(self, path: str):
        from .ddpg import _load_legacy_state_dict
        for name, net, tgt in (("policy.pt", self.policy, self.policy_target), ("q.pt", self.q, self.q_target)):
            sd = _load_legacy_state_dict(os.path.join(path, name))
            net.load_state_dict({k: v for k, v in sd.items() if k.startswith("model.")})
            tgt.load_state_dict(net.state_dict())
        return self

    def set_lr(self, lr: float):
        self.lr = float(lr)
        for opt in (self.opt_pi, self.opt_q):
            for g in opt.param_groups:
                g["lr"] = self.lr
