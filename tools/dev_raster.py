#!/usr/bin/env python
"""Dev A/B of the two row rasterisers: kernel time of mpc_build_grid for fp32 / fp64 distances + a hash of the grids.
    MPC_RASTER_ROWS=0|1 python tools/dev_raster.py [H] [episodes]"""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_mpc_lanemerging_b200 import _lib, synthetic  # noqa: E402
from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
p = _lib.default_params(); p.future_t, p.future_s = synthetic.horizon_settings(H)
eng = MpcEngine(p, 0, max_batch=B)
D = states_to_device(synthetic.make_states(B, "moderate", seed=0), "cuda:0")
for dt, cell in ((torch.float32, 5), (torch.float64, 9)):
    eng.set_timing(True)
    ms = []
    for i in range(6):
        g = eng.build_grid(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], dist_dtype=dt)
        if i:
            ms.append(eng.last_kernel_ms()[1])
    eng.set_timing(False)
    h = hashlib.sha1(g["obstacles"].cpu().numpy().tobytes() + g["distances"].cpu().numpy().tobytes()).hexdigest()[:16]
    nbytes = B * eng.num_t * eng.num_s_stride * cell
    print(f"ROWS={os.environ.get('MPC_RASTER_ROWS')} H={H} B={B} {dt}: {min(ms):.4f} ms  {nbytes / min(ms) / 1e6:.0f} GB/s  hash {h}")
