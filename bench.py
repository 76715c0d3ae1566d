#!/usr/bin/env python
"""Headline benchmark: MPC gap-evaluations / second (BASELINE.json metric) on synthetic traffic.

A "step" is one pass of the hot path -- traffic prediction over all layers + S-T obstacle/distance
evaluation + jerk-limited DP solve + back-track + crash test (one `st.get_appropriate_base_st_path_and_obstacles`
equivalent per episode) -- over one batch of synthetic episodes per GPU.

Default workload = BASELINE.json configs[1]: 4096 parallel episodes, moderate traffic, horizon 50
(51 x 9001 cells), 1 x B200.  `--horizon 17` gives the reference's published grid (18 x 3001).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RESULT_OUT = sys.stdout
METRIC = "mpc_gap_evals_per_sec"
UNIT = "gap-evals/s"


def b_alg(num_t: int, num_s: int) -> int:
    """Algorithmic bytes per gap-evaluation: the dense S-T grid the reference's solver
    boundary consumes, each cell once (1 B mask + 4 B fp32 distance), plus the fp32 path out."""
    return num_t * num_s * 5 + num_t * 4


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(H: int, episodes: int):
    """DRAM bytes (read + write) of the dominant kernel per launch, from the committed ncu --set full capture
    (profiles/*_traffic.json, written by tools/summarise_profiles.py), scaled to this run's batch."""
    import glob
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            rec = json.load(open(f)).get(str(H))
            if rec:
                return (rec["dram_bytes_read"] + rec["dram_bytes_write"]) * episodes / rec["episodes"]
        except Exception:
            pass
    return None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML every few milliseconds
    (the timed region of this benchmark lasts ~0.1-0.3 s, shorter than nvidia-smi's own sampling loop)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.stop_flag, self.err = index, [], set(), False, None
        self.max_mhz, self.t_begin = None, None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.stop_flag:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.t_begin is not None:
                    for name, bit in self.BAD.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.004)
        except Exception as e:          # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def begin(self):                      # call at the start of the timed region
        self.t_begin = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        self.stop_flag = True
        self.thread.join(timeout=2)
        self.samples = [c for (t, c) in self.samples if self.t_begin is not None and self.t_begin <= t <= t_end]
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "samples": 0, "reasons": [self.err or "no samples"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": sorted(self.reasons)}


def make_params(H: int):
    from rl_mpc_lanemerging_b200 import _lib, synthetic
    p = _lib.default_params()
    p.future_t, p.future_s = synthetic.horizon_settings(H)
    return p


# --------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_port_rate(H, traffic, seed, n, nthreads):
    """The C restatement (oracle/mpc_oracle.c, layered DP) on n states of the same workload."""
    from oracle import cpu_oracle as O
    from rl_mpc_lanemerging_b200 import synthetic
    S = synthetic.make_states(n, traffic, seed=seed)
    p = O.horizon_params(H)
    t0 = time.perf_counter()
    O.plan_batch(p, S["ego"], S["cars_x"], S["cars_v"], S["cars_a"], S["n_cars"], H + 1, layered=True, nthreads=nthreads)
    return n / (time.perf_counter() - t0)


_W = {}


def _ref_worker_init(H):
    from oracle import cpu_oracle as O, ref_harness
    _W["O"] = O
    _W["p"] = O.horizon_params(H)
    _W["st_cy"] = ref_harness.load_st_cy()
    _W["t"] = np.arange(H + 1) * 0.3


def _ref_worker(job):
    """One gap-evaluation the reference's way: grid (C restatement of st.py:25-70, the Python original
    cannot travel to the GPU box) + the reference's own compiled st_cy.solve_s_t_path_fast."""
    O, p, st_cy = _W["O"], _W["p"], _W["st_cy"]
    ego, xs, vs, acs = job
    st = O.make_state((ego[0], ego[1]), ego[2], ego[3], xs, vs, acs)
    ob, di, sv = O.build_grid(p, st)
    if st_cy is not None:
        seq = st_cy.solve_s_t_path_fast(ob.view(np.bool_), sv, _W["t"], ego[2], ego[3], di, p.d_weight, p.v_weight, p.a_weight,
                                        p.j_weight, p.desired_speed, p.max_speed, p.a_min, p.a_max, p.j_min, p.j_max,
                                        p.min_allowed_distance)
    else:
        seq = O.solve(p, ob, di, sv, p.t_disc, ego[2], ego[3], layered=False)["s_seq"]
    return float(seq[-1])


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import ref_harness
    from rl_mpc_lanemerging_b200 import synthetic
    H, cores = args.horizon, os.cpu_count() or 1
    have_ref = ref_harness.load_st_cy() is not None
    per_eval = {17: 0.009, 25: 0.03, 50: 0.14, 100: 0.45}.get(H, 0.14)       # seconds per gap-eval per core
    n = max(cores, min(4096, int(round(cores * 1.0 / per_eval))))          # ~1 s of wall clock per step
    S = synthetic.make_states(n, args.traffic, seed=args.seed)
    jobs = [(S["ego"][b], S["cars_x"][b, :S["n_cars"][b]], S["cars_v"][b, :S["n_cars"][b]], S["cars_a"][b, :S["n_cars"][b]])
            for b in range(n)]
    with mp.get_context("fork").Pool(cores, initializer=_ref_worker_init, initargs=(H,)) as pool:
        chunk = max(1, n // (cores * 4))
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_worker, jobs[:cores * 2], chunksize=1)
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, jobs, chunksize=chunk)
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = n / (ms * 1e-3)
    num_t, num_s = H + 1, int(round(30.0 * 0.3 * H / 0.05)) + 1 if H != 17 else 3001
    kind = "reference" if have_ref else "port"
    sample = (f"{n} of the workload's episodes per step; grid build by the C restatement of st.py:25-70, solve by "
              + ("the reference's compiled st_cy.solve_s_t_path_fast (oracle/_ref)" if have_ref else "the C Dijkstra port")
              + f"; {cores} worker processes")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"batched MPC gap-evaluation, {args.traffic} traffic, horizon={H} ({num_t}x{num_s} cells)",
                       "horizon": H, "traffic": args.traffic, "episodes_per_step": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=RESULT_OUT, flush=True)



def work_counts(H: int, traffic: str):
    """Nodes finalised / offers made per gap-evaluation (CPU model of the kernel on 200 states of the same generator) and executed
    warp instructions per launch of the dominant kernel (ncu), from the committed profiles/r02_work_counts.json
    (tools/work_counts.py writes it; the bench never runs the model itself)."""
    try:
        W = json.load(open(os.path.join(ROOT, "profiles", "r02_work_counts.json")))
        rec = W["model"].get(f"{H}:{traffic}")
        ncu = W.get("ncu", {}).get(str(H))
        return rec, ncu
    except Exception:
        return None, None


def time_plan(eng, D, out, steps, flush, torch, mode="fast"):
    """(mean step ms by CUDA events, mean kernel ms [predictor, first DP launch, later DP launches]) over `steps` plans, L2 flushed between."""
    a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    for _ in range(3):
        eng.plan(*a, mode=mode, out=out); flush.zero_()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for e0, e1 in ev:
        e0.record(); eng.plan(*a, mode=mode, out=out); e1.record(); flush.zero_()
    torch.cuda.synchronize()
    eng.set_timing(True)
    km = []
    for _ in range(2):
        eng.plan(*a, mode=mode, out=out)
        km.append(eng.last_kernel_ms()); flush.zero_()
    eng.set_timing(False)
    return sum(e0.elapsed_time(e1) for e0, e1 in ev) / steps, [float(np.mean([k[i] for k in km])) for i in range(3)]


def run_sweep(args, local, world, rank, first, flush, torch, sharding):
    """BASELINE.json configs[4]: traffic density x horizon, same batch per GPU, device-resident inputs, CUDA events, max over ranks."""
    from rl_mpc_lanemerging_b200 import synthetic
    from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device
    dev, B = f"cuda:{local}", args.batch
    peak = peak_hbm()[0]
    rows, meta = [], []
    for H in (17, 25, 50, 100):
        for traffic in ("low", "medium", "default", "moderate", "fast"):
            # one handle per point: a handle fits the ring of its first launch shape to the traffic it sees in its first calls
            eng = MpcEngine(make_params(H), device=local, max_batch=B)
            D = states_to_device(synthetic.make_states(B, traffic, seed=args.seed, first_episode=first), dev)
            out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode=args.mode)
            ms, km = time_plan(eng, D, out, args.sweep_steps, flush, torch, args.mode)
            info, c = eng.fast32_info(), eng.counters()
            rows.append([ms] + km)
            meta.append((H, traffic, eng.num_t, eng.num_s_max - 1, info["handed_on"], info["first_shape_handed_on"], c["fallback_problems"],
                         float((out["reached_t"] == eng.num_t - 1).float().mean())))
            eng.close()
            del out, D
    red = sharding.reduce_max([x for r in rows for x in r], dev)
    res = {}
    for i, (H, traffic, T, num_s, handed, handedA, back, full) in enumerate(meta):
        ms, pred, dp, fb = (float(x) for x in red[4 * i:4 * i + 4])
        res[f"H{H}:{traffic}"] = {"gap_evals_per_s": world * B / (ms * 1e-3), "ms": ms, "predict_ms": pred, "dp_ms": dp, "dp_later_ms": fb,
                                  "roofline_frac": b_alg(T, num_s) * B / ((dp + fb) * 1e-3) / 1e9 / peak,
                                  "handed_to_64bit_kernel": handed, "outgrew_first_ring": handedA - handed if handedA >= handed else 0,
                                  "exact_kernel_problems": back, "full_horizon_fraction": full}
    res["note"] = (f"{B} episodes per GPU, {args.sweep_steps} timed steps per point; roofline_frac = dense-grid-equivalent bytes / all DP launches "
                   "of the step / measured HBM peak; rank 0's counters")
    return res

# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from rl_mpc_lanemerging_b200 import sharding, synthetic
    from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    bound_cores, bound_how = None, "single rank"
    if world > 1:
        # one process per GPU: keep the launch thread and the pinned buffers on the GPU's own NUMA node
        bound_cores, bound_how = sharding.bind_to_local_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        dist.init_process_group("nccl", device_id=torch.device(dev))
    H, B, K, W = args.horizon, args.batch, args.steps, max(args.warmup, 3)
    eng = MpcEngine(make_params(H), device=local, max_batch=B)
    # episodes are sharded by global id: rank r owns episodes [r*B, (r+1)*B) (replicas, no exchange on the path)
    first, _last = sharding.shard_range(world * B, rank, world)
    S = synthetic.make_states(B, args.traffic, seed=args.seed, first_episode=first)
    D = states_to_device(S, dev)
    out = eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode=args.mode)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)           # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.plan(D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"], mode=args.mode, out=out)

    sampler = ClockSampler(local)
    sampler.start()                       # (every rank: its GPU's clocks explain its step time) NVML comes up during the warm-up; only samples inside the timed region are kept
    for _ in range(W):
        step(); flush.zero_()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    sampler.begin()
    t_wall = time.perf_counter()
    for i in range(K):                                                       # the host only enqueues: no synchronisation inside the timed region
        ev[i][0].record(); step(); ev[i][1].record()
        flush.zero_()                                                        # L2 flush between timed iterations (outside the events)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks = sampler.stop()
    step_ms = sum(a.elapsed_time(b_) for a, b_ in ev)
    # per-kernel times (events recorded inside the library; reading them synchronises, hence a separate, untimed pass)
    eng.set_timing(True)
    dp_ms, pred_ms, fb_ms = [], [], []
    for i in range(min(K, 5)):
        step()
        a, b_, c = eng.last_kernel_ms()
        pred_ms.append(a); dp_ms.append(b_); fb_ms.append(c)
        flush.zero_()
    counters = eng.counters()
    f32_info = eng.fast32_info()
    eng.set_timing(False)

    # ---- end-to-end through the host-buffer API: H2D of the step's states + D2H of its results, every step ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Sp = {k: torch.from_numpy(np.ascontiguousarray(S[k])).pin_memory() for k in ("ego", "cars_x", "cars_v", "cars_a", "n_cars")}   # the step's inputs: pinned host memory
    for _ in range(2):
        eng.plan_host(Sp["ego"], Sp["cars_x"], Sp["cars_v"], Sp["cars_a"], Sp["n_cars"], mode=args.mode)
    barrier()
    e2e_ms = 0.0
    for _ in range(K):
        t0 = time.perf_counter()
        e0.record()
        r = eng.plan_host(Sp["ego"], Sp["cars_x"], Sp["cars_v"], Sp["cars_a"], Sp["n_cars"], mode=args.mode)
        e1.record(); e1.synchronize()
        e2e_ms += max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
        flush.zero_()
    barrier()
    T = eng.num_t
    h2d = B * (4 * 8 + 2 * eng.nmax * 8 + 4)
    d2h = B * (T * 4 + T * 8 + 8 + 4 + 1 + 8 + 8)
    full = float((r["reached_t"] == T - 1).mean())

    # per-rank view (multi-GPU runs): the slowest rank defines the step; its clocks and kernel times say why
    per_rank = sharding.gather([step_ms / K, e2e_ms / K, float(np.mean(pred_ms)), float(np.mean(dp_ms)), float(np.mean(fb_ms)),
                                float((clocks or {}).get("sm_mhz") or 0.0)], dev) if world > 1 else None
    step_ms, e2e_ms = sharding.reduce_max([step_ms, e2e_ms], dev)        # the slowest rank defines the step
    # ---- the one HBM-bound kernel of the path: the dense S-T rasteriser behind mpc_build_grid (st.py:25-70 output) ----
    grid_res = None
    try:
        Bg = min(B, 256)
        eng.set_timing(True)
        sub = [D[k][:Bg].contiguous() for k in ("ego", "cars_x", "cars_v", "cars_a", "n_cars")]
        ras = {}
        for name, dt in (("fp32", torch.float32), ("fp64", torch.float64)):
            ras[name] = []
            for i in range(5):
                g_ = eng.build_grid(*sub, dist_dtype=dt)
                if i:
                    ras[name].append(eng.last_kernel_ms()[1])
                del g_
        eng.set_timing(False)
        gbytes = Bg * T * eng.num_s_stride * 5
        gbs = gbytes / (min(ras["fp32"]) * 1e-3) / 1e9
        gbs64 = Bg * T * eng.num_s_stride * 9 / (min(ras["fp64"]) * 1e-3) / 1e9
        grid_res = {"kernel": "rasterise_rows_kernel<float>", "episodes": Bg, "bytes_written": gbytes, "ms": min(ras["fp32"]), "achieved_gbs": gbs,
                    "frac_of_hbm_peak": gbs / peak_hbm()[0],
                    "fp64_distances": {"kernel": "rasterise_rows_kernel<double>", "bytes_written": Bg * T * eng.num_s_stride * 9,
                                       "ms": min(ras["fp64"]), "achieved_gbs": gbs64, "frac_of_hbm_peak": gbs64 / peak_hbm()[0]},
                    "note": "mask u8 + fp32 (fp64) distance per cell, written once (the dense grid the "
                    "reference's solver consumes); used only by the drop-in API, the fused planner never materialises it"}
    except Exception as e:              # noqa: BLE001
        grid_res = {"error": repr(e)}
    # ---- K2, the st_cy drop-in itself: the DP kernel fed by DENSE grids resident in HBM (mpc_solve_dense, st_cy.pyx:315) ----
    dense_res = None
    if rank == 0 and args.mode == "fast":
        try:
            Bg = min(B, 4096)
            sub = [D[k][:Bg].contiguous() for k in ("ego", "cars_x", "cars_v", "cars_a", "n_cars")]
            v0, a0 = sub[0][:, 2].contiguous(), sub[0][:, 3].contiguous()
            ref = eng.plan(*sub, mode="fast")
            ref = {k: v.clone() for k, v in ref.items()}
            dense_res = {}
            for name, dt, cell in (("fp32_distances", torch.float32, 5), ("fp64_distances", torch.float64, 9)):
                g_ = eng.build_grid(*sub, dist_dtype=dt)
                eng.set_timing(True)
                dms = []
                for i in range(4):
                    r_ = eng.solve_dense(g_["obstacles"], g_["distances"], g_["start_s"], g_["delta_s"], g_["num_s"], v0, a0, mode="fast")
                    if i:
                        km_ = eng.last_kernel_ms()
                        dms.append(km_[1] + km_[2])                # first DP launch + the launches on what it hands over
                    flush.zero_()
                eng.set_timing(False)
                dbytes = (T * (eng.num_s_max - 1) * cell + T * 4) * Bg
                dense_res[name] = {"ms": min(dms), "gap_evals_per_s": Bg / (min(dms) * 1e-3), "algorithmic_gbs": dbytes / (min(dms) * 1e-3) / 1e9,
                                   "frac_of_hbm_peak": dbytes / (min(dms) * 1e-3) / 1e9 / peak_hbm()[0],
                                   "same_sequences_as_fused": bool(torch.equal(r_["idx"], ref["idx"]))}
                del g_, r_
            dense_res["episodes"] = Bg
            tr = measured_traffic("dense32_50", Bg) if H == 50 else None
            if tr:                              # ncu --set full of the same kernel (profiles/r02_dense_fast32_f32_h50.txt)
                alg = (T * (eng.num_s_max - 1) * 5 + T * 4) * Bg
                dense_res["fp32_distances"].update({"dram_bytes_per_launch": tr, "algorithmic_bytes_per_launch": alg, "dram_over_algorithmic": tr / alg,
                                                    "note": "the DP reads only the span its frontier can reach (mask byte + distance of every cell of that span for the "
                                                            "blocked bits, the distance again for every winner): dram_over_algorithmic is that reachable share "
                                                            "of the dense grid; the kernel is issue bound like the fused one, not HBM bound"})
        except Exception as e:          # noqa: BLE001
            dense_res = {"error": repr(e)}
    sweep_res = None
    if not args.no_sweep and args.mode == "fast":
        try:
            eng.close()
            sweep_res = run_sweep(args, local, world, rank, first, flush, torch, sharding)
        except Exception as e:          # noqa: BLE001
            sweep_res = {"error": repr(e)}
    env_res = None
    if args.env_ticks > 0:              # every rank runs its own environments; the job rate is the sum over ranks
        try:
            eng.close()
            rate, take = env_steps_per_sec(local, world, args.env_envs, args.env_ticks, args.seed + rank)
            r = sharding.reduce_sum([rate, take], dev)
            env_res = {"value": float(r[0]), "unit": "env-steps/s", "envs_per_gpu": args.env_envs, "ticks": args.env_ticks,
                       "controller": "RL proposes + MPC vetoes (combined_moderate_1 semantics), H=17 grid, fast mode",
                       "planner_takeover_fraction": float(r[1]) / world, "published_takeover_fraction": 0.037,
                       "policy": "published actor pretrained_models/ddpg_moderate1_extended (tests/golden/policy_moderate1.npz)",
                       "world_model": "merge_gym.MergeEnv, Krauss car-following traffic (SUMO-free); published fraction: saved_data.csv:46 (SUMO)"}
            try:                        # the same tick as ONE CUDA-graph launch (bit-identical trajectories: tests/test_graphed_tick_gpu.py)
                grate = env_steps_graphed(local, args.env_envs, args.env_ticks, args.seed + rank)
                gr = sharding.reduce_sum([grate], dev)
                env_res["cuda_graph_tick"] = {"value": float(gr[0]), "unit": "env-steps/s", "launches_per_tick": 1,
                                              "switches": "FUSED_ENV_STEP + SYNC_FREE_TAKEOVER (no host sync in the tick)"}
            except Exception as e:      # noqa: BLE001
                env_res["cuda_graph_tick"] = {"error": repr(e)}
        except Exception as e:          # noqa: BLE001  -- the secondary figure must never break the headline line
            env_res = {"error": repr(e)}
    train_res = None
    if args.train_ticks > 0:            # DDPG training: envs sharded over the ranks, one NCCL all-reduce of the flat gradient per step
        try:
            frames, gsteps, nbytes = train_steps_per_sec(local, world, args.env_envs, args.train_ticks, args.seed + rank)
            r = sharding.reduce_sum([frames], dev)
            g = sharding.reduce_max([gsteps], dev)
            train_res = {"value": float(r[0]), "unit": "env-frames/s", "grad_steps_per_s": float(g[0]), "envs_per_gpu": args.env_envs,
                         "ticks": args.train_ticks, "minibatch_per_gpu": 4096, "allreduce_bytes_per_step": nbytes if world > 1 else 0,
                         "collective": "NCCL all-reduce of ONE flat fp32 gradient (actor + critic)" if world > 1 else "none (1 rank)",
                         "config": "train_medium_1.json semantics (DDPG, traffic 7 m/s / 1.8 s); library hyper-parameters unpinned; THROUGHPUT "
                                   "setting (one gradient step of 4096 samples per tick of envs_per_gpu frames)"}
            if rank == 0 and world == 1:            # the library defaults = the reference's update-to-data ratio (what the learning curve used)
                try:
                    train_res["reference_update_ratio"] = train_reference_ratio(local, args.seed)
                except Exception as e:  # noqa: BLE001
                    train_res["reference_update_ratio"] = {"error": repr(e)}
        except Exception as e:          # noqa: BLE001
            train_res = {"error": repr(e)}
    if rank == 0:
        num_s = eng.num_s_max - 1
        ms_per_step = step_ms / K
        value = world * B / (ms_per_step * 1e-3)
        dp, fb, pred = float(np.mean(dp_ms)), float(np.mean(fb_ms)), float(np.mean(pred_ms))
        peak, peak_src = peak_hbm()
        f32 = f32_info
        solved = B - (f32["handed_on"] if f32["in_use"] else 0)          # problems the dominant kernel finished itself
        bytes_per_launch = b_alg(T, num_s) * solved
        achieved = bytes_per_launch / (dp * 1e-3) / 1e9
        kernel = ("fast32_kernel" if f32["in_use"] else "fast_pull_kernel") if args.mode == "fast" else "exact_push_kernel"
        wc, ncu = work_counts(H, args.traffic)
        compute = None
        if wc is not None and args.mode == "fast":
            clk = (clocks or {}).get("sm_mhz") or 1965.0
            compute = {"nodes_per_gap_eval": wc["nodes"], "offers_per_gap_eval": wc["pushes"],
                       "nodes_per_s": value * wc["nodes"], "offers_per_s": value * wc["pushes"],
                       "source": f"CPU model of the kernel on {wc['states']} states of this workload (profiles/r02_work_counts.json)"}
            if ncu:
                wi = ncu["warp_instructions"] * B / ncu["episodes"]
                compute.update({"warp_instructions_per_gap_eval": wi / B, "issue_slot_frac": wi / (dp * 1e-3 * 148 * 4 * clk * 1e6),
                                "issue_slots": f"148 SMs x 4 schedulers x {clk:.0f} MHz; warp instructions of {ncu['kernel']} from ncu ({ncu['profile']})"})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32" if (args.mode == "fast" and f32["in_use"]) else ("i64" if args.mode == "fast" else "f64"), "data": "synthetic",
                "config": {"workload": f"batched MPC gap-evaluation: {B} parallel episodes per GPU, {args.traffic} traffic, "
                                       f"horizon={H} ({T}x{num_s} cells)", "horizon": H, "traffic": args.traffic,
                           "episodes_per_gpu": B, "mode": args.mode, "parallelism": f"replicas x{world} (episodes sharded, no collective)",
                           "arithmetic": (f"integer-cell kinematics; first attempt: 32-bit fixed-point labels (2^-{f32['frac_bits']}) min-combined with the native "
                                          "32-bit shared-memory atomic; hand-overs: 48-bit labels (2^-18); fp64 obstacle / threshold tests")
                           if args.mode == "fast" else "fp64, reference operation order",
                           "l2": "256 MiB buffer written between timed iterations (outside the timed events)",
                           "inputs": "resident in HBM (fp64 SoA state)",
                           "cpu_affinity": (f"rank 0 bound to {len(bound_cores)} cores ({bound_how})" if bound_cores
                                            else f"not bound ({bound_how})")},
                "wall_ms_per_step_incl_flush": wall_ms / K,
                "kernel_ms": {"predict_layers": pred, "dp": dp, "dp_later_launches": fb},
                "full_horizon_fraction": full, "handed_to_64bit_kernel": f32["handed_on"], "exact_kernel_problems": counters["fallback_problems"],
                "gpu_launches": counters["kernels_launched"] * K,
                "e2e": {"value": world * B / (e2e_ms / K * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "MpcEngine.plan_host -> mpc_plan_host (pinned host buffers, copies inside the timed region)"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": measured_traffic(H, B), "kernel": kernel, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": bytes_per_launch,
                             "stage_frac": b_alg(T, num_s) * B / ((dp + fb) * 1e-3) / 1e9 / peak,
                             "note": "achieved = dense-grid-equivalent bytes (num_t*num_s*5 + num_t*4 per gap-eval, SURVEY 8d) of the problems the kernel "
                                     "finished / its launch time; stage_frac = all problems / all DP launches; the fused kernel never materialises the "
                                     "grid (`traffic` = its DRAM bytes, ncu): it is bound by instruction issue, see `compute`"},
                "clocks": clocks}
        if per_rank is not None:
            line["per_rank"] = {"columns": ["ms_per_step", "e2e_ms_per_step", "predict_ms", "dp_ms", "dp_later_ms", "sm_mhz_median"],
                                "rows": [[round(x, 4) for x in r] for r in per_rank]}
        if compute is not None:
            line["compute"] = compute
        if dense_res is not None:
            line["dense_solve"] = dense_res
        if sweep_res is not None:
            line["sweep"] = sweep_res
        if grid_res is not None:
            line["grid_build"] = grid_res
        if env_res is not None:
            line["env_steps"] = env_res
        if train_res is not None:
            line["train"] = train_res
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            probe = cpu_port_rate(H, args.traffic, args.seed, cores * 2, cores)
            n = int(min(B, max(cores * 2, probe * 15.0)))                     # ~15 s of CPU work
            rate = cpu_port_rate(H, args.traffic, args.seed, n, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {n} episodes of the same workload, C restatement (layered DP, fp64), {cores} threads"}
        # the last ~2 KB of the line: the secondary figures in compact form (a record that keeps only the tail of stdout still has them)
        def _r(x, n=3):
            return None if x is None else round(float(x), n)
        summ = {"gap_evals_per_s": _r(value, 0), "e2e_gap_evals_per_s": _r(line["e2e"]["value"], 0), "roofline_frac": _r(line["roofline"]["frac"]),
                "issue_slot_frac": _r((line.get("compute") or {}).get("issue_slot_frac")),
                "kernel_ms": [_r(pred), _r(dp), _r(fb)]}
        if isinstance(grid_res, dict) and "frac_of_hbm_peak" in grid_res:
            summ["rasteriser_frac_of_hbm_peak"] = {"fp32": _r(grid_res["frac_of_hbm_peak"]), "fp64": _r(grid_res["fp64_distances"]["frac_of_hbm_peak"])}
        if isinstance(dense_res, dict) and "fp32_distances" in dense_res:
            summ["k2_dense_solve"] = {"episodes": dense_res.get("episodes"), "gap_evals_per_s_fp32": _r(dense_res["fp32_distances"]["gap_evals_per_s"], 0),
                                      "gap_evals_per_s_fp64": _r(dense_res["fp64_distances"]["gap_evals_per_s"], 0),
                                      "dram_over_algorithmic": _r(dense_res["fp32_distances"].get("dram_over_algorithmic")),
                                      "same_sequences_as_fused": dense_res["fp32_distances"]["same_sequences_as_fused"] and dense_res["fp64_distances"]["same_sequences_as_fused"]}
        if isinstance(sweep_res, dict) and "error" not in sweep_res:
            summ["sweep_M_gap_evals_per_s"] = {k: _r(v["gap_evals_per_s"] / 1e6) for k, v in sweep_res.items() if isinstance(v, dict) and "gap_evals_per_s" in v}
        if isinstance(env_res, dict) and "value" in env_res:
            summ["env_steps_per_s"] = {"eager": _r(env_res["value"], 0), "cuda_graph_tick": _r((env_res.get("cuda_graph_tick") or {}).get("value"), 0),
                                       "takeover_fraction": _r(env_res.get("planner_takeover_fraction"))}
        if isinstance(train_res, dict) and "value" in train_res:
            summ["train_env_frames_per_s"] = _r(train_res["value"], 0)
        if per_rank is not None:
            summ["per_rank_ms_per_step"] = [_r(r[0]) for r in per_rank]
        if "cpu_baseline" in line:
            summ["cpu_port_gap_evals_per_s"] = _r(line["cpu_baseline"]["value"], 0)
        line["summary"] = summ
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def env_steps_per_sec(local, world, n_envs, ticks, seed, extra=None):
    """Secondary figure of BASELINE.json's metric: closed-loop env-steps/s of the combined controller
    (configs/combined_moderate_1.json semantics: 5 policy forwards + 5 predictor steps + 1 gap-evaluation per tick,
    + a second gap-evaluation for the episodes the planner takes over; reference dqn.py:117-200) on the published
    solver grid (H=17).  World model = merge_gym.MergeEnv (SUMO-free); policy = the PUBLISHED actor
    pretrained_models/ddpg_moderate1_extended/policy.pt (committed as tests/golden/policy_moderate1.npz).  Returns (env-steps/s over all ranks' envs of this rank, takeover fraction)."""
    import torch
    from rl_mpc_lanemerging_b200 import ddpg, merge_gym, st
    from rl_mpc_lanemerging_b200.config import Settings
    Settings.reset()
    Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.BASE_TRAFFIC_INTERVAL = 20, 11.0, 1.2      # combined_moderate_1.json
    Settings.TEST_ST_STRICTLY_BETTER, Settings.CUDA_DEVICE, Settings.ALT_J_WEIGHT = False, local, 0.1
    for k, v in (extra or {}).items():          # (hint leg: FUSED_ENV_STEP / PLAN_COST_HINTS)
        setattr(Settings, k, v)
    st.refresh_engine()
    env = merge_gym.MergeEnv(n_envs, seed=seed)
    agent = ddpg.DDPGAgent.load_npz(os.path.join(ROOT, "tests", "golden", "policy_moderate1.npz"), device=f"cuda:{local}")
    env.reset()
    take = torch.zeros((), dtype=torch.float32, device=f"cuda:{local}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # warm-up: every branch of the controller must have run once (lazy CUDA module loading, cuBLAS heuristics, allocator):
    # the planner take-over is rare (~1 % of the episodes per tick) and would otherwise first happen inside the timed region
    from rl_mpc_lanemerging_b200.prediction import BatchedState
    st.do_st_control(BatchedState(*(t[:64].contiguous() for t in env.state.args())))
    # (24 warm ticks: by then the episodes have reached the merge area and take-overs of realistic size have happened; on multi-rank
    # runs a first-time stall inside the timed region was observed to cost seconds -- profiles/r01_scaling.json)
    for phase, n in (("warm", 24), ("timed", ticks)):
        if phase == "timed":
            torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            speed, takeover = agent.do_combined_control(env.state)
            jerk = ((speed - env.state.ego[:, 2]) / Settings.TICK_LENGTH - env.state.ego[:, 3]) / Settings.TICK_LENGTH
            _obs, _r, done, _info = env.step(jerk)
            agent.reset_time(done)
            if phase == "timed":
                take += takeover.float().mean()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st.refresh_engine()
    Settings.reset()
    return n_envs * ticks / (ms * 1e-3), float(take) / ticks


def env_steps_graphed(local, n_envs, ticks, seed):
    """env_steps_per_sec's tick captured once in a CUDA graph and replayed (rl_mpc_lanemerging_b200/graphed_tick.py)."""
    import torch
    from rl_mpc_lanemerging_b200 import ddpg, merge_gym, st
    from rl_mpc_lanemerging_b200.config import Settings
    from rl_mpc_lanemerging_b200.graphed_tick import GraphedTick
    Settings.reset()
    Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.BASE_TRAFFIC_INTERVAL = 20, 11.0, 1.2      # combined_moderate_1.json
    Settings.TEST_ST_STRICTLY_BETTER, Settings.CUDA_DEVICE, Settings.ALT_J_WEIGHT = False, local, 0.1
    Settings.FUSED_ENV_STEP = Settings.SYNC_FREE_TAKEOVER = True
    st.refresh_engine()
    try:
        env = merge_gym.MergeEnv(n_envs, seed=seed)
        agent = ddpg.DDPGAgent.load_npz(os.path.join(ROOT, "tests", "golden", "policy_moderate1.npz"), device=f"cuda:{local}")
        env.reset()
        gt = GraphedTick(env, agent).capture(warmup_ticks=32)
        gt.replay(4)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        gt.replay(ticks)
        e1.record(); torch.cuda.synchronize()
        return n_envs * ticks / (e0.elapsed_time(e1) * 1e-3)
    finally:
        Settings.reset()
        st.refresh_engine()


def train_reference_ratio(local, seed, ticks=12):
    """The trainer at the library defaults (config.py: TRAIN_NUM_ENVS 256, TRAIN_MINIBATCH 1024, TRAIN_SAMPLES_PER_FRAME 100): the
    reference's update-to-data ratio -- one gradient step on 100 replayed samples per environment frame (ddpg.py:46-81 with the
    preset's update_frequency 1) -- kept with larger minibatches: 25 gradient steps of 1024 samples per tick of 256 frames.  This is
    the configuration tools/learning_curve.py trains with; returns env-frames/s and grad steps/s of this rank."""
    import torch
    from rl_mpc_lanemerging_b200 import merge_gym, st, trainer
    from rl_mpc_lanemerging_b200.config import Settings
    Settings.reset()
    Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.BASE_TRAFFIC_INTERVAL, Settings.CUDA_DEVICE = 20, 7.0, 1.8, local   # train_medium_1.json
    st.refresh_engine()
    try:
        n_envs, mb = int(Settings.TRAIN_NUM_ENVS), int(Settings.TRAIN_MINIBATCH)
        upt = max(1, round(n_envs * float(Settings.TRAIN_SAMPLES_PER_FRAME) / mb))
        env = merge_gym.MergeEnv(n_envs, seed=seed)
        tr = trainer.DDPGTrainer(env, device=f"cuda:{local}", lr=2e-4, seed=0, minibatch_size=mb, updates_per_tick=upt,
                                 replay_start_size=n_envs * 4, replay_buffer_size=65536)
        tr.train(n_envs * 8)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0 = tr.grad_steps
        torch.cuda.synchronize(); e0.record()
        tr.train(n_envs * ticks)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        return {"env_frames_per_s": n_envs * ticks / (ms * 1e-3), "grad_steps_per_s": (tr.grad_steps - g0) / (ms * 1e-3), "envs": n_envs,
                "minibatch": mb, "updates_per_tick": upt, "samples_per_frame": float(Settings.TRAIN_SAMPLES_PER_FRAME)}
    finally:
        st.refresh_engine()
        Settings.reset()


def train_steps_per_sec(local, world, n_envs, ticks, seed):
    """BASELINE.json configs[3] (train_medium_1.json): DDPG training, environments sharded over the ranks, ONE NCCL all-reduce
    of the flat actor+critic gradient per step.  Returns this rank's env-frames/s incl. the gradient steps, grad steps/s
    and the all-reduce payload."""
    import torch
    from rl_mpc_lanemerging_b200 import merge_gym, st, trainer
    from rl_mpc_lanemerging_b200.config import Settings
    Settings.reset()
    Settings.CRASH_MIN_S, Settings.OTHER_CAR_SPEED, Settings.BASE_TRAFFIC_INTERVAL, Settings.CUDA_DEVICE = 20, 7.0, 1.8, local   # train_medium_1.json
    st.refresh_engine()
    env = merge_gym.MergeEnv(n_envs, seed=seed)
    tr = trainer.DDPGTrainer(env, device=f"cuda:{local}", lr=2e-4, seed=0, minibatch_size=4096, replay_start_size=n_envs,
                             replay_buffer_size=max(4 * n_envs, 65536))
    tr.train(n_envs * 3)                                          # warm-up: fills the replay ring, first updates
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0 = tr.grad_steps
    torch.cuda.synchronize(); e0.record()
    tr.train(n_envs * ticks)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = (n_envs * ticks / (ms * 1e-3), (tr.grad_steps - g0) / (ms * 1e-3), tr.allreduce_bytes)
    st.refresh_engine()
    Settings.reset()
    return out


def _claim_stdout():
    """The contract is ONE JSON line on stdout: libraries (NCCL prints its version banner) write to fd 1 directly, so fd 1 is
    pointed at stderr for the whole run and the result line goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    global RESULT_OUT
    RESULT_OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--horizon", type=int, default=50)
    ap.add_argument("--batch", type=int, default=4096, help="episodes per GPU")
    ap.add_argument("--traffic", default="moderate")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--env-ticks", type=int, default=40, help="ticks of the closed-loop env-steps/s measurement (0 = skip)")
    ap.add_argument("--env-envs", type=int, default=8192, help="environments per GPU for it (BASELINE configs[2])")
    ap.add_argument("--train-ticks", type=int, default=10, help="ticks of the DDPG training throughput figure (0 = skip)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the traffic x horizon sweep (BASELINE.json configs[4])")
    ap.add_argument("--sweep-steps", type=int, default=5)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
