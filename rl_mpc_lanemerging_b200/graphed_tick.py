"""One closed-loop tick of the combined controller -- policy rollout, gap-evaluation, planner take-over, environment step --
as ONE CUDA-graph launch.

The tick of the reference is `RLAgent.do_combined_control(state)` -> `control.set_ego_speed` -> `control.step()`
(dqn.py:117-200, evaluated by control.evaluate_control, control.py:268-340).  With `Settings.FUSED_ENV_STEP` and
`Settings.SYNC_FREE_TAKEOVER` the batched version of that tick has no host synchronisation left (the episode list of the take-over
is built and counted on the device, the environment step is one kernel), which makes it capturable: ~70 launches of a few
microseconds each become one graph launch.  The random numbers of the environment are drawn inside the graph from the
environment's own generator (registered with the graph), so a replayed tick walks through exactly the trajectory the eager tick
walks through -- tests/test_graphed_tick_gpu.py compares the two bit for bit over 120 ticks.
"""
from __future__ import annotations

import torch

from .config import Settings
from .prediction import tdiv


class GraphedTick:
    def __init__(self, env, agent):
        if not (getattr(Settings, "FUSED_ENV_STEP", False) and getattr(Settings, "SYNC_FREE_TAKEOVER", False)):
            raise RuntimeError("GraphedTick needs Settings.FUSED_ENV_STEP and Settings.SYNC_FREE_TAKEOVER (a tick without host syncs)")
        if not env.auto_reset:
            raise RuntimeError("GraphedTick needs an auto-resetting environment (finished episodes restart inside the tick)")
        self.env, self.agent = env, agent
        B, dev = env.B, env.device
        self.reward = torch.zeros(B, dtype=torch.float64, device=dev)
        self.done = torch.zeros(B, dtype=torch.bool, device=dev)
        self.crashed = torch.zeros(B, dtype=torch.bool, device=dev)
        self.merged = torch.zeros(B, dtype=torch.bool, device=dev)
        self.takeover = torch.zeros(B, dtype=torch.bool, device=dev)
        self.graph = None
        self.ticks_run = 0

    def _tick(self):
        env, agent = self.env, self.agent
        speed, takeover = agent.do_combined_control(env.state)
        tick = float(Settings.TICK_LENGTH)
        jerk = tdiv(tdiv(speed - env.state.ego[:, 2], tick) - env.state.ego[:, 3], tick)      # the jerk that produces this speed
        _obs, reward, done, info = env.step(jerk)
        agent.reset_time(done)
        self.reward.copy_(reward); self.done.copy_(done); self.takeover.copy_(takeover)
        self.crashed.copy_(info["crashed"]); self.merged.copy_(info["merged"])

    def eager(self, n: int = 1):
        """n ticks launched kernel by kernel (also the warm-up before capture())."""
        for _ in range(n):
            self._tick()
            self.agent.takeover_history.clear()
        self.ticks_run += n

    def capture(self, warmup_ticks: int = 32):
        """Warm up (lazy module loading, the planner's launch-shape adaptation of its first calls, allocator) and capture one tick."""
        side = torch.cuda.Stream(device=self.env.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.eager(warmup_ticks)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        g.register_generator_state(self.env.gen)
        with torch.cuda.graph(g):
            self._tick()
        self.agent.takeover_history.clear()
        self.graph = g
        return self

    def replay(self, n: int = 1):
        if self.graph is None:
            raise RuntimeError("capture() first")
        for _ in range(n):
            self.graph.replay()
        self.ticks_run += n
