"""Dev-only: the 32-bit-key kernel on the GPU -- results against the 64-bit kernel on the same states, then a sweep of launch shapes.
    python tools/dev_sweep32.py H B [blocks:threads ...]        (blocks = MPC_F32_BLOCKS / 32)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_mpc_lanemerging_b200 import synthetic, _lib
from rl_mpc_lanemerging_b200.engine import MpcEngine, states_to_device


def run(H, B, env, reps=4, traffic="moderate", ref=None):
    for k in ("MPC_FAST32", "MPC_F32_BLOCKS", "MPC_F32_THREADS", "MPC_FAST_BLOCKS", "MPC_FAST_THREADS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    p = _lib.default_params()
    p.future_t, p.future_s = synthetic.horizon_settings(H)
    eng = MpcEngine(p, 0, max_batch=B)
    D = states_to_device(synthetic.make_states(B, traffic, seed=0, first_episode=int(os.environ.get("SWEEP_FIRST", "0"))), "cuda:0")
    a = (D["ego"], D["cars_x"], D["cars_v"], D["cars_a"], D["n_cars"])
    out = eng.plan(*a, mode="fast")
    torch.cuda.synchronize()
    eng.set_timing(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best, km = 1e9, None
    for _ in range(reps):
        ev[0].record(); eng.plan(*a, mode="fast", out=out); ev[1].record(); torch.cuda.synchronize()
        t = ev[0].elapsed_time(ev[1])
        if t < best: best, km = t, eng.last_kernel_ms()
    info, c = eng.fast32_info(), eng.counters()
    res = {k: v.cpu().numpy().copy() for k, v in out.items()}
    msg = ""
    if ref is not None:
        same = (res["idx"] == ref["idx"]).all(1)
        rel = np.abs(res["cost"] - ref["cost"]) / np.maximum(ref["cost"], 1e-9)
        msg = f" | vs 64-bit kernel: idx identical {same.mean():.4f}, cost rel max {rel.max():.2e}, reached equal {np.array_equal(res['reached_t'], ref['reached_t'])}, crash equal {np.array_equal(res['crash'], ref['crash'])}"
    print(f"H={H} B={B} {traffic} {env}: {best:.3f} ms -> {B / best * 1e3:.0f} gap-evals/s  (pred, dp, fb) = {tuple(round(x, 3) for x in km)}  f32={info} 64-bit hand-backs={c['fallback_problems']} counters={eng.raw_counters() if hasattr(eng, 'raw_counters') else ''}{msg}", flush=True)
    eng.close()
    return res


if __name__ == "__main__":
    H, B = int(sys.argv[1]), int(sys.argv[2])
    traffic = os.environ.get("SWEEP_TRAFFIC", "moderate")
    ref = run(H, B, {"MPC_FAST32": "0"}, traffic=traffic)
    run(H, B, {}, ref=ref, traffic=traffic)
    for cfg in sys.argv[3:]:
        if cfg.startswith("fb"):                      # fbN:T -> launch shape of the 64-bit kernel (the hand-over launches)
            nb, th = cfg[2:].split(":")
            run(H, B, {"MPC_FAST_BLOCKS": str(32 * int(nb)), "MPC_FAST_THREADS": th}, ref=ref, traffic=traffic)
            continue
        nb, th = cfg.split(":")
        run(H, B, {"MPC_F32_BLOCKS": str(32 * int(nb)), "MPC_F32_THREADS": th}, ref=ref, traffic=traffic)
