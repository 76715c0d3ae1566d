#!/usr/bin/env python
"""CPU study behind DESIGN.md §8 "Fewer nodes": what the cost hint (coarse probe) and the reachability heuristic save.

Runs the C model of the fast kernel (oracle/mpc_oracle.c, oracle/bound_model.py) on the bench workload and prints, per
traffic type: nodes expanded per problem by today's kernel (zone bound + unbounded retry), with the probe hint, and with
hint + heuristic -- every retry of the kernel's ladder included -- plus the widest layer span (what the label ring must
hold: decides how many blocks fit an SM).

    python tools/bound_study.py [N=150] [H=50] [traffic ...]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bound_model as BM          # noqa: E402
from oracle import cpu_oracle as O            # noqa: E402
from rl_mpc_lanemerging_b200 import synthetic as S   # noqa: E402


def study(N, H, traffic, margin=1.1, mult=(20, 3)):
    p = O.horizon_params(H)
    q = BM.probe_params(p, *mult)
    zb = BM.zone_bound(p)
    st = S.make_states(N, traffic, seed=0)
    L = O.lib()
    base = 0
    tot = {"hint": 0, "hint+h": 0}
    spans = {"zone": [], "hint": [], "hint+h": []}
    first_ok, tried, probe_nodes = 0, 0, 0
    for i in range(N):
        e, n = st["ego"][i], st["n_cars"][i]
        s = O.make_state((e[0], e[1]), e[2], e[3], st["cars_x"][i, :n], st["cars_v"][i, :n], st["cars_a"][i, :n])
        ob, di, sv = O.build_grid(p, s)
        unb = lambda: O.solve_fast_model_ex(p, ob, di, sv, p.t_disc, e[2], e[3], prune_cost=0)   # noqa: E731
        rz = O.solve_fast_model_ex(p, ob, di, sv, p.t_disc, e[2], e[3], prune_cost=zb)
        spans["zone"].append(L.orc_last_max_span())
        nb = rz["nodes"] + (unb()["nodes"] if rz["reached_t"] < H else 0)
        base += nb
        hint, pn = BM.probe_hint(p, q, s, margin)
        probe_nodes += pn
        if hint is None or hint >= zb:               # no usable hint / zone-crossing plan: the kernel's standard attempts
            for k in tot:
                tot[k] += nb
            continue
        tried += 1
        h = BM.heuristic_table(p, ob, di)
        r = O.solve_fast_model_ex(p, ob, di, sv, p.t_disc, e[2], e[3], prune_cost=hint)
        spans["hint"].append(L.orc_last_max_span())
        tot["hint"] += r["nodes"] + (0 if r["reached_t"] == H else nb)
        r = BM.solve_with_heuristic(p, ob, di, sv, e[2], e[3], hint, h)
        spans["hint+h"].append(L.orc_last_max_span())
        tot["hint+h"] += r["nodes"]
        if r["reached_t"] == H:
            first_ok += 1
        else:                                        # second rung: zone bound (the heuristic still applies), then unbounded
            r2 = BM.solve_with_heuristic(p, ob, di, sv, e[2], e[3], zb - 1, h)
            tot["hint+h"] += r2["nodes"] + (0 if r2["reached_t"] == H else unb()["nodes"])
    print(f"H={H} {traffic}: nodes/problem today {base / N:.0f}; " +
          "; ".join(f"{k} {v / N:.0f} ({v / base:.2f})" for k, v in tot.items()) +
          f"; probe {probe_nodes / N:.0f} coarse nodes; first attempt ok {first_ok}/{tried}")
    for k, v in spans.items():
        v = np.array(v)
        if len(v):
            print(f"   widest span [{k}]: median {int(np.median(v))}, max {v.max()} cells")


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    for tr in (sys.argv[3:] or ["moderate", "fast", "default", "low"]):
        study(N, H, tr)
