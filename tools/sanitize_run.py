"""Dev-only: small workloads through every kernel, to be run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_mpc_lanemerging_b200 import synthetic, _lib
from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
H = int(sys.argv[1]) if len(sys.argv) > 1 else 17
B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
p = _lib.default_params(); p.future_t, p.future_s = synthetic.horizon_settings(H)
eng = MpcEngine(p, 0, max_batch=B)
D = states_to_device(synthetic.make_states(B, "moderate", seed=2, kind="mixed"), "cuda:0")
a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
for mode in ("fast", "exact"):
    out = eng.plan(*a, mode=mode)
fine = eng.finer_fit(out["s_seq"], out["reached_t"], D["ego"])
g32 = eng.build_grid(*a, dist_dtype=torch.float32)                # segment rasteriser, fp32 distances; K2 (fast32 on dense grids) on them
eng.solve_dense(g32["obstacles"], g32["distances"], g32["start_s"], g32["delta_s"], g32["num_s"], D["ego"][:, 2].contiguous(), D["ego"][:, 3].contiguous(), mode="fast")
os.environ["MPC_RASTER_ROWS"] = "0"                                  # the 8-cells-per-thread rasteriser
eng.build_grid(*a)
del os.environ["MPC_RASTER_ROWS"]
g = eng.build_grid(*a)
eng.solve_dense(g["obstacles"], g["distances"], g["start_s"], g["delta_s"], g["num_s"], D["ego"][:, 2].contiguous(), D["ego"][:, 3].contiguous(), mode="fast")
eng.predict_step_with_ego(*a, D["ego"][:, 2].contiguous(), 0.2, 5.1)
eng.state_vector(*a)
eng.selftest_search(*a)
torch.cuda.synchronize()
print("sanitize_run done", eng.counters())
