"""Dev-only: quick timing of the planner on the GPU (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_lanemerging_b200 import synthetic, _lib
from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device

def run(H, B, mode, reps=3):
    p = _lib.default_params()
    p.future_t, p.future_s = synthetic.horizon_settings(H)
    eng = MpcEngine(p, 0, max_batch=B)
    D = states_to_device(synthetic.make_states(B, "moderate", seed=0), "cuda:0")
    out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode=mode)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for _ in range(reps):
        ev[0].record()
        eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode=mode, out=out)
        ev[1].record(); torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    c = eng.counters()
    print(f"H={H} B={B} mode={mode}: {best:.3f} ms -> {B/best*1e3:.0f} gap-evals/s  threads={os.environ.get('MPC_THREADS','default')} counters={c} reached_full={(out['reached_t']==eng.num_t-1).float().mean().item():.3f}", flush=True)
    eng.close()

if __name__ == "__main__":
    for H, B in ((17, 4096), (50, 4096)):
        for mode in ("fast", "exact"):
            run(H, B, mode)
