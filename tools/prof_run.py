"""Dev-only: one plan() call per configuration, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_mpc_lanemerging_b200 import synthetic, _lib
from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
H = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
p = _lib.default_params(); p.future_t, p.future_s = synthetic.horizon_settings(H)
eng = MpcEngine(p, 0, max_batch=B)
D = states_to_device(synthetic.make_states(B, os.environ.get("PROF_TRAFFIC", "moderate"), seed=0, first_episode=int(os.environ.get("PROF_FIRST", "0"))), "cuda:0")
if mode in ("dense32", "dense64"):            # K2: the DP kernel on dense grids resident in HBM (mpc_solve_dense)
    g = eng.build_grid(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], dist_dtype=torch.float32 if mode == "dense32" else torch.float64)
    v0, a0 = D["ego"][:, 2].contiguous(), D["ego"][:, 3].contiguous()
    for _ in range(3):
        out = eng.solve_dense(g["obstacles"], g["distances"], g["start_s"], g["delta_s"], g["num_s"], v0, a0, mode="fast")
    torch.cuda.synchronize()
    print("done", eng.counters())
    sys.exit(0)
for _ in range(3):
    if mode == "grid":
        out = eng.build_grid(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], dist_dtype=torch.float32)
    else:
        out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode=mode)
torch.cuda.synchronize()
print("done", eng.counters())
