"""Dev-only: where does a closed-loop tick go? wall vs GPU time, launches per tick (torch profiler)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_mpc_lanemerging_b200 import ddpg, merge_gym, st
from rl_mpc_lanemerging_b200.config import Settings
from rl_mpc_lanemerging_b200.prediction import BatchedState
n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
Settings.reset()
Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.BASE_TRAFFIC_INTERVAL = 20, 11.0, 1.2
Settings.TEST_ST_STRICTLY_BETTER, Settings.ALT_J_WEIGHT = False, 0.1
st.refresh_engine()
env = merge_gym.MergeEnv(n_envs, seed=1); agent = ddpg.DDPGAgent(device="cuda:0", seed=1); env.reset()
st.do_st_control(BatchedState(*(t[:64].contiguous() for t in env.state.args())))
def tick():
    speed, takeover = agent.do_combined_control(env.state)
    jerk = ((speed - env.state.ego[:, 2]) / Settings.TICK_LENGTH - env.state.ego[:, 3]) / Settings.TICK_LENGTH
    _o, _r, done, _i = env.step(jerk); agent.reset_time(done)
for _ in range(30): tick()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(40): tick()
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 40
# split
def timed(fn, n=40):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n
ctrl = timed(lambda: agent.do_combined_control(env.state))
jerk0 = torch.zeros(n_envs, dtype=torch.float64, device="cuda:0")
stp = timed(lambda: env.step(jerk0))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10): tick()
    torch.cuda.synchronize()
ev = prof.key_averages()
gpu_us = sum(e.device_time_total for e in ev if e.device_type == torch.autograd.DeviceType.CUDA) / 10
n_k = sum(e.count for e in ev if e.device_type == torch.autograd.DeviceType.CUDA) / 10
print(f"envs {n_envs}: wall/tick {wall*1e3:.2f} ms ({n_envs/wall:.0f} env-steps/s); control {ctrl*1e3:.2f} ms, env.step {stp*1e3:.2f} ms; GPU busy/tick {gpu_us/1e3:.2f} ms over {n_k:.0f} kernels")
top = sorted([e for e in ev if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: -e.device_time_total)[:8]
for e in top: print(f"   {e.device_time_total/10/1e3:7.3f} ms/tick  x{e.count/10:5.1f}  {e.key[:90]}")
