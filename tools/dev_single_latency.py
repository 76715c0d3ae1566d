"""Dev-only: latency of the unbounded 64-bit pass on single problems (the tail of a step is the slowest hand-over problem), per launch shape.
    python tools/dev_single_latency.py H first_episode"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_lanemerging_b200 import synthetic, _lib
from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device

H, first = int(sys.argv[1]), int(sys.argv[2])
p = _lib.default_params(); p.future_t, p.future_s = synthetic.horizon_settings(H)
S = synthetic.make_states(4096, "moderate", seed=0, first_episode=first)
D = states_to_device(S, "cuda:0")
eng = MpcEngine(p, 0, max_batch=4096)
out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode="fast")
cost, reached = out["cost"].cpu().numpy(), out["reached_t"].cpu().numpy()
eng.close()
flag = np.nonzero((cost >= 2.0e6) | (reached < H))[0]
print("flagged-like problems:", len(flag))
res = {}
for name, env in (("2x512", {"MPC_FAST32": "0", "MPC_FAST_BOUND": "0"}), ("1x1024", {"MPC_FAST32": "0", "MPC_FAST_BOUND": "0", "MPC_FAST_BLOCKS": "32", "MPC_FAST_THREADS": "1024"})):
    for k in ("MPC_FAST32", "MPC_FAST_BOUND", "MPC_FAST_BLOCKS", "MPC_FAST_THREADS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    e1 = MpcEngine(p, 0, max_batch=1)
    e1.set_timing(True)
    ms = []
    for b in flag:
        a = tuple(D[k][b:b + 1].contiguous() for k in ("ego", "cars_x", "cars_v", "cars_a", "n_cars"))
        best = 1e9
        for _ in range(2):
            e1.plan(*a, mode="fast")
            best = min(best, e1.last_kernel_ms()[1])
        ms.append(best)
    e1.close()
    res[name] = np.array(ms)
    print(name, "mean %.3f ms  p90 %.3f  max %.3f" % (res[name].mean(), np.percentile(res[name], 90), res[name].max()))
o = np.argsort(-res["2x512"])[:10]
print("ten slowest (2x512 alone on the GPU -> 1x1024):", [(round(float(res["2x512"][i]), 3), round(float(res["1x1024"][i]), 3)) for i in o])
