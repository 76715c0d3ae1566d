"""CPU models of the fast kernel's cost bounds -- TEST INFRASTRUCTURE ONLY (see oracle/mpc_oracle.h).

Nothing here is part of the reference: the reference's Dijkstra (st_cy.pyx:315-399) has no bound.  These
functions exist so that tests/ can check, on CPU and over many states, the two claims the CUDA fast
kernel's pruning rests on (rl_mpc_lanemerging_b200/csrc/mpc_fast.cu):

1. COST BOUND.  Dropping every node whose label exceeds U leaves all nodes with label <= U exactly as the
   unbounded pass computes them; if the bounded pass reaches the horizon its answer IS the unbounded
   answer.  `solve_with_ladder` mirrors the kernel's retry ladder (hint -> standard zone bound -> none).

2. REACHABILITY HEURISTIC (hinted solves: mpc_reach.cu + the HINT instances of the fast kernel).  h(t, k) = a lower bound of
   the cost still to pay from cell k of layer t that depends on the CELL ONLY (not on the node's
   history) and is consistent (h(p) <= edge(p -> c) + h(c) for every edge).  Dropping the nodes with
   label + h > U is then exact in the same sense as 1. (proof: DESIGN.md).  The bound used is the
   speed term alone: the ego cannot jump a blocked interval (band + penalty zones are wider than one
   step), so from cell k it advances at most D = cap - k cells in the n remaining steps, cap = the
   highest final cell reachable from k's 64-cell bucket; by convexity of the speed cost
   V the remaining cost is >= n * V(D / n), evaluated on the piecewise-linear interpolation of the
   fixed-point table in exact integer arithmetic: h = (n - r) * V[q] + r * V[q + 1], q = D div n, r = D mod n.
"""
from __future__ import annotations

import numpy as np

from . import cpu_oracle as O

FX_ONE = 262144.0
H_INF = np.uint64(1) << np.uint64(50)


def zone_bound(p) -> float:
    """The standard first bound of the fast kernel: one step inside a penalty zone (st_cy.pyx:34-38)."""
    return p.d_weight * 1000000.0 / max(p.min_allowed_distance, 1.0)


HINT_RETRY = 1.36       # middle rung of the hinted ladder (mpc_handle::hint_retry)


def solve_with_ladder(p, ob, di, sv, v0, a0, hint=None, retry=HINT_RETRY):
    """The kernel's attempts for one problem: [hint -> retry * hint] -> zone bound -> unbounded.  A hint at or above
    the zone bound skips the zone-bounded attempts.  Returns (result, nodes expanded over all attempts, attempts)."""
    T = ob.shape[0]
    zb = zone_bound(p)
    ladder = []
    if hint is not None and 0.0 < hint < 1e9:
        ladder.append(hint)
        if hint < zb:
            if retry > 1.0 and hint * retry < zb:
                ladder.append(hint * retry)
            ladder.append(zb)
    else:
        ladder.append(zb)
    ladder.append(0.0)          # 0 = no bound
    nodes = 0
    for n_try, U in enumerate(ladder, 1):
        r = O.solve_fast_model_ex(p, ob, di, sv, p.t_disc, v0, a0, prune_cost=U)
        nodes += r["nodes"]
        if r["reached_t"] == T - 1 or U == 0.0:
            return r, nodes, n_try
    raise AssertionError("unreachable")


# ---- reachability heuristic ----------------------------------------------------------------------------
def speed_table(p) -> np.ndarray:
    """The fast kernel's fixed-point speed-cost table (DevParams::vtab)."""
    v = np.arange(256, dtype=np.float64) * p.s_disc / p.t_disc
    return np.rint(np.minimum(p.v_weight * (v - p.desired_speed) ** 2, 16000.0) * FX_ONE).astype(np.int64)


H_BUCKET = 64        # cells per bucket (MPC_BUCKET_SHIFT of the kernel's per-layer lookup tables)


def heuristic_table(p, ob, di, bucket: int = H_BUCKET, caps_out=None) -> np.ndarray:
    """h in label units for every cell, u64 [T, S]; blocked and dead-end cells get H_INF.

    cap is kept per BUCKET of `bucket` cells (what a kernel would stage per layer: one u16 per bucket):
    cap[T-1][b] = highest free cell of bucket b;  cap[t][b] = max cap[t+1][b'] over the buckets b' = b .. b + reach
    that a cell of b can step into (reach = (bucket - 1 + vmax_c) div bucket; -1 when b has no free cell).  The
    bucket granularity only makes cap optimistic (h smaller), never inconsistent: every edge k -> k' has
    cap[t][bucket(k)] >= cap[t+1][bucket(k')] by construction (with h = 0 on the last layer that is all the
    exactness proof needs).  It is sharp when blocked intervals are wider than `reach` buckets, so that the ego is
    not credited with a jump over a car (400 cells against 256 at the published Settings)."""
    T, S = ob.shape
    blocked = (ob != 0) | (di < p.min_allowed_distance)
    vmax_c = int(np.floor(p.max_speed * p.t_disc / p.s_disc + 1e-9))
    V = speed_table(p)
    vstar = int(np.argmin(V[:vmax_c + 1]))
    Vm = V[np.minimum(np.arange(258), vstar)]
    nb = (S + bucket - 1) // bucket
    reach = (bucket - 1 + vmax_c) // bucket
    cells = np.arange(S)

    def free_hi(row):                                   # highest free cell per bucket, -1 = none
        out = np.full(nb, -1, np.int64)
        idx = np.flatnonzero(~row)
        if len(idx):
            np.maximum.at(out, idx // bucket, idx)
        return out

    h = np.full((T, S), H_INF, dtype=np.uint64)
    cap = free_hi(blocked[T - 1])
    h[T - 1][~blocked[T - 1]] = 0
    if caps_out is not None:
        caps_out[T - 1] = cap
    for t in range(T - 2, -1, -1):
        n = T - 1 - t
        row = blocked[t] if t > 0 else np.zeros(S, bool)        # the start cell is never tested (st_cy.pyx:383)
        new = np.full(nb, -1, np.int64)
        for j in range(reach + 1):
            new[:nb - j] = np.maximum(new[:nb - j], cap[j:])
        new = np.where(free_hi(row) >= 0, new, -1)
        c = new[cells // bucket]
        ok = (~row) & (c >= 0)
        D = np.clip(c - cells, 0, vstar * n)
        q, r = D // n, D % n
        h[t] = np.where(ok, ((n - r) * Vm[q] + r * Vm[q + 1]).astype(np.uint64), H_INF)
        cap = new
        if caps_out is not None:
            caps_out[t] = cap
    return h


def blocked_intervals(blocked_row: np.ndarray):
    """Maximal runs of blocked cells as [x, y) -- what LayerDesc::blk holds for a layer."""
    edges = np.flatnonzero(np.diff(np.concatenate(([0], blocked_row.view(np.int8), [0]))))
    return [(int(edges[i]), int(edges[i + 1])) for i in range(0, len(edges), 2)]


def bucket_caps_like_kernel(p, ob, di) -> np.ndarray:
    """The loops of reach_caps_kernel (rl_mpc_lanemerging_b200/csrc/mpc_reach.cu) restated step by step on the
    blocked INTERVALS: i32 [T, NB], -1 = no path.  tests/ compares it with the mask-based caps of heuristic_table."""
    T, S = ob.shape
    blocked = (ob != 0) | (di < p.min_allowed_distance)
    vmax_c = int(np.floor(p.max_speed * p.t_disc / p.s_disc + 1e-9))
    NB, reach = (S + 63) >> 6, (63 + vmax_c) >> 6
    caps = np.full((T, NB), -1, np.int64)
    cap = np.full(NB + 8, -1, np.int64)
    for t in range(T - 1, -1, -1):
        blk = [] if t == 0 else blocked_intervals(blocked[t])
        nv = np.full(NB, -1, np.int64)
        for j in range(NB):
            c = min((j << 6) + 63, S - 1)
            for x, y in reversed(blk):
                if x <= c < y:
                    c = x - 1
            if c >= (j << 6):
                nv[j] = c if t == T - 1 else max(cap[j + r] for r in range(reach + 1))
        cap[:NB] = nv
        caps[t] = nv
    return caps


def solve_with_heuristic(p, ob, di, sv, v0, a0, U, h, U_fx=None):
    """Fast-kernel model that drops a node when label + h > U (U in cost units or, U_fx, in label units; h from
    heuristic_table)."""
    import ctypes as C
    T, S = ob.shape
    idx = np.zeros(T, np.int32); seq = np.zeros(T, np.float64); cost = C.c_double(); counts = (C.c_int64 * 2)()
    ob = np.ascontiguousarray(ob, np.uint8); di = np.ascontiguousarray(di, np.float64); sv = np.ascontiguousarray(sv, np.float64)
    h = np.ascontiguousarray(h, np.uint64)
    dp = C.POINTER(C.c_double)
    r = O.lib().orc_solve_fast_model_h(C.byref(p), T, S, ob.ctypes.data_as(C.c_void_p), di.ctypes.data_as(dp),
                                       sv.ctypes.data_as(dp), C.c_double(p.t_disc), C.c_double(v0), C.c_double(a0),
                                       C.c_uint64(int(U_fx) if U_fx is not None else int(round(U * FX_ONE))), h.ctypes.data_as(C.POINTER(C.c_uint64)), None,
                                       idx.ctypes.data_as(C.POINTER(C.c_int)), seq.ctypes.data_as(dp), C.byref(cost), counts)
    return dict(reached_t=r, idx=idx, s_seq=seq, cost=cost.value, nodes=int(counts[0]), pushes=int(counts[1]))


# ---- coarse probe ----------------------------------------------------------------------------------------
def probe_params(p, s_mult=20, t_mult=3):
    """Settings of the probe grid (MpcEngine.make_probe): same Settings, coarser discretisation."""
    q = O.Params()
    for name, _ in O.Params._fields_:
        setattr(q, name, getattr(p, name))
    q.s_disc, q.t_disc = p.s_disc * s_mult, p.t_disc * t_mult
    return q


def probe_hint(p, q, state, margin=1.1):
    """margin * (num_t-1)/(probe num_t-1) * cost of the probe plan (mpc_plan_probed); None when the probe plan is
    incomplete.  Also returns the probe's node count."""
    ob, di, sv = O.build_grid(q, state)
    r = O.solve_fast_model_ex(q, ob, di, sv, q.t_disc, state.ego_v, state.ego_a, prune_cost=0.0)
    Tq = ob.shape[0]
    Tp = O.grid_dims(p, O.get_ego_s(state.ego_x, state.ego_y))[0]
    if r["reached_t"] != Tq - 1:
        return None, r["nodes"]
    return margin * r["cost"] * (Tp - 1) / (Tq - 1), r["nodes"]
