"""Writes profiles/r02_work_counts.json: the work one gap-evaluation is, per (horizon, traffic) of the bench's sweep.

  model   nodes finalised and offers made (edges relaxed) by the bounded first attempt + the unbounded retry of the problems
          that need one, counted by the CPU model of the kernels (oracle/mpc_oracle.c) on N states of the bench's generator
  ncu     executed warp instructions of one launch of the dominant kernel, read from an ncu report (tools/ncu_sass.py total)

    python tools/work_counts.py [--states 200] [--ncu H:report.ncu-rep:episodes ...]

bench.py reads the file (it never runs the model): `compute` = nodes/s, offers/s, warp instructions per gap-evaluation and the
fraction of the SMs' issue slots the kernel uses."""
import argparse
import json
import os
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "profiles", "r02_work_counts.json")


def _one(job):
    H, traffic, i, frac, bound = job
    from oracle import cpu_oracle as O
    from rl_mpc_lanemerging_b200 import synthetic
    p = O.horizon_params(H)
    S = synthetic.make_states(1, traffic, seed=0, first_episode=i)
    n = S["n_cars"][0]
    st = O.make_state((S["ego"][0, 0], S["ego"][0, 1]), S["ego"][0, 2], S["ego"][0, 3], S["cars_x"][0, :n], S["cars_v"][0, :n], S["cars_a"][0, :n])
    ob, di, sv = O.build_grid(p, st)
    a = (p, ob, di, sv, p.t_disc, st.ego_v, st.ego_a)
    r = O.solve_fast_model_q(*a, frac, 0, bound)
    nodes, pushes, retried = r["nodes"], r["pushes"], 0
    if r["reached_t"] < H:
        r2 = O.solve_fast_model_ex(*a, prune_cost=0.0)
        nodes += r2["nodes"]; pushes += r2["pushes"]; retried = 1
    return nodes, pushes, retried


def frac_and_bound(H):
    """DevParams::f32_frac / f32_bound through the emulated library's derive_params (tests/emu)."""
    import ctypes as C
    from oracle import cpu_oracle as O
    from rl_mpc_lanemerging_b200._lib import PARAM_FIELDS, MpcParams
    from tests.emu import build_emu
    emu = C.CDLL(build_emu.build())
    op, p = O.horizon_params(H), MpcParams()
    for n in PARAM_FIELDS:
        setattr(p, n, getattr(op, n))
    info = (C.c_int64 * 3)()
    z = np.zeros(1)
    emu.emu_plan32(C.byref(p), 0, 32, None, None, None, None, 128, 0, 128, 0, None, None, None, None, None, None, None, info)
    return int(info[0]), int(info[1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=200)
    ap.add_argument("--ncu", nargs="*", default=[])
    a = ap.parse_args()
    res = json.load(open(OUT)) if os.path.exists(OUT) else {"model": {}, "ncu": {}}
    with ProcessPoolExecutor(8) as ex:
        for H in ((17, 25, 50, 100) if a.states > 0 else ()):             # --states 0: keep the model counts, refresh the ncu part only
            frac, bound = frac_and_bound(H)
            n = a.states if H <= 50 else max(a.states // 4, 16)
            for traffic in ("low", "medium", "default", "moderate", "fast"):
                R = np.array(list(ex.map(_one, [(H, traffic, i, frac, bound) for i in range(n)], chunksize=4)), float)
                res["model"][f"{H}:{traffic}"] = {"nodes": float(R[:, 0].mean()), "pushes": float(R[:, 1].mean()), "retried_fraction": float(R[:, 2].mean()),
                                                  "states": n, "label_frac_bits": frac}
                print(H, traffic, res["model"][f"{H}:{traffic}"], flush=True)
    for spec in a.ncu:
        H, rep, episodes = spec.split(":")
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass.py"), rep, "--min", "2.0"], capture_output=True, text=True).stdout
        kernel = out.splitlines()[0][2:].split("(")[0].split()[-1]
        total = int(out.splitlines()[1].split()[3].rstrip(","))
        res["ncu"][H] = {"kernel": kernel, "warp_instructions": total, "episodes": int(episodes), "profile": os.path.basename(rep)}
        print("ncu", H, res["ncu"][H])
    json.dump(res, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
