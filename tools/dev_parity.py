"""Dev-only: compare the fast kernel with the CPU model problem by problem and print the mismatches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import helpers
from oracle import cpu_oracle as O
from rl_mpc_lanemerging_b200 import synthetic
from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device

H = int(sys.argv[1]); B = int(sys.argv[2]); traffic = sys.argv[3]; kind = sys.argv[4]; seed = int(sys.argv[5]) if len(sys.argv) > 5 else 12
op = O.horizon_params(H)
eng = MpcEngine(helpers.mpc_params_from_oracle(op), device=0, max_batch=B)
S = synthetic.make_states(B, traffic, seed=seed, kind=kind)
D = states_to_device(S, "cuda:0")
out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode="fast")
torch.cuda.synchronize()
idx, cost, rt = out["idx"].cpu().numpy(), out["cost"].cpu().numpy(), out["reached_t"].cpu().numpy()
bad = 0
for b in range(B):
    st = helpers.oracle_state(O, S, b)
    ob, di, sv = O.build_grid(op, st)
    m = O.solve_fast_model(op, ob, di, sv, op.t_disc, st.ego_v, st.ego_a)
    if m["reached_t"] != rt[b] or m["cost"] != cost[b] or not np.array_equal(m["idx"], idx[b]):
        bad += 1
        first = int(np.argmax(m["idx"] != idx[b])) if not np.array_equal(m["idx"], idx[b]) else -1
        print(f"b={b}: model reached {m['reached_t']} cost {m['cost']:.6f} | gpu reached {rt[b]} cost {cost[b]:.6f} | first idx diff at t={first}"
              f" model {m['idx'][max(first,0):first+3]} gpu {idx[b][max(first,0):first+3]} ego={S['ego'][b]} n={S['n_cars'][b]}")
print(f"{bad} of {B} differ; counters {eng.counters()}")
