"""Dev-only: quick timing of the planner on the GPU (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_lanemerging_b200 import synthetic, _lib
from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device

def run(H, B, mode, reps=3, threads=None, blocks=None, lst=0):
    for k, v in (("MPC_FAST_THREADS", threads), ("MPC_FAST_BLOCKS", blocks)):      # 0 / None = the library's default
        if v: os.environ[k] = str(v)
        else: os.environ.pop(k, None)
    p = _lib.default_params()
    p.future_t, p.future_s = synthetic.horizon_settings(H)
    eng = MpcEngine(p, 0, max_batch=B)
    D = states_to_device(synthetic.make_states(B, "moderate", seed=0), "cuda:0")
    out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode=mode)
    torch.cuda.synchronize()
    eng.set_timing(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9; km = None
    for _ in range(reps):
        ev[0].record()
        eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode=mode, out=out)
        ev[1].record(); torch.cuda.synchronize()
        t = ev[0].elapsed_time(ev[1])
        if t < best: best, km = t, eng.last_kernel_ms()
    c = eng.counters()
    print(f"H={H} B={B} mode={mode} threads={threads or 'default'} blocks={blocks or 'default'}: {best:.3f} ms -> {B/best*1e3:.0f} gap-evals/s  kernels(pred,dp,fb)={tuple(round(x,3) for x in km)} fallback={c['fallback_problems']}", flush=True)
    eng.close()

if __name__ == "__main__":
    cfgs = sys.argv[1:] or ["17:4096:fast:0", "17:4096:fast:256", "17:4096:fast:512", "17:4096:fast:1024", "50:4096:fast:512", "50:4096:fast:1024", "50:4096:fast:768", "17:4096:exact:0", "50:2048:exact:0"]
    for c in cfgs:
        f = c.split(":")
        run(int(f[0]), int(f[1]), f[2], threads=int(f[3]) or None, blocks=int(f[4]) if len(f) > 4 else None, lst=int(f[5]) if len(f) > 5 else 0)
