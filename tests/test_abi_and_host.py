"""-m "not gpu": the C-ABI library loads and exports every symbol include/mpcb200.h declares; host-side
logic (parameter derivation, Settings loader, synthetic generator, fast-arithmetic model) -- no GPU compute."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mpcb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mpc_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from rl_mpc_lanemerging_b200 import _lib, build
    build.build_lib()
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mpcb200.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in _lib.SYMBOLS"
    assert lib.mpc_abi_version() == 1
    p = _lib.default_params()
    assert (p.s_disc, p.t_disc, p.future_s, p.future_t) == (0.05, 0.30, 150.0, 5.0)


def test_no_cpu_fallback_without_device():
    """Without a GPU the library refuses to create a handle (there is no CPU path to fall back to)."""
    from rl_mpc_lanemerging_b200 import _lib
    lib = _lib.load()
    if lib.mpc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    p = _lib.default_params()
    rc = lib.mpc_create(C.byref(p), 0, 16, 32, C.byref(h))
    assert rc == _lib.E_NODEVICE and b"no CPU fallback" in lib.mpc_last_error()


def test_settings_loads_reference_style_json(tmp_path):
    from rl_mpc_lanemerging_b200.config import Settings
    cfg = {"TASK": "ST", "OTHER_CAR_SPEED": 11.0, "CRASH_MIN_S": 20, "FUTURE_T": 5.0, "JERK_VALUES_DQN": {"0": -5, "1": 0, "2": 5},
           "SOME_UNKNOWN_KEY": 3}
    f = tmp_path / "c.json"
    f.write_text(json.dumps(cfg))
    Settings.reset()
    Settings.load_from_file(str(f))
    assert Settings.CRASH_MIN_S == 20 and Settings.OTHER_CAR_SPEED == 11.0 and Settings.SOME_UNKNOWN_KEY == 3
    assert Settings.JERK_VALUES_DQN == {0: -5, 1: 0, 2: 5}          # dict values get int keys (reference config.py:168-169)
    assert Settings.S_DISCRETIZATION == 0.05                          # untouched defaults survive
    Settings.reset()
    assert not hasattr(Settings, "SOME_UNKNOWN_KEY")


def test_synthetic_states_are_shard_invariant():
    """Episode e's state depends only on (seed, e): sharding over GPUs cannot change the workload."""
    from rl_mpc_lanemerging_b200 import synthetic
    A = synthetic.make_states(64, "moderate", seed=3)
    B = synthetic.make_states(16, "moderate", seed=3, first_episode=32)
    for k in A:
        assert np.array_equal(A[k][32:48], B[k])
    assert (A["n_cars"] <= synthetic.NMAX).all() and A["n_cars"].max() > 3
    x = A["cars_x"]
    for b in range(64):
        n = A["n_cars"][b]
        assert np.all(np.diff(x[b, :n]) < 0)                          # front -> back
        d = np.hypot(x[b, :n] - A["ego"][b, 0], -1.6 - A["ego"][b, 1])
        assert np.all(d < synthetic.SENSOR_RADIUS)


@pytest.mark.parametrize("H,n", [(17, 60), (50, 12)])
def test_fast_arithmetic_model_within_tolerance(oracle, H, n):
    """The fast kernel's arithmetic (integer kinematics, fp32 edge cost, fp64 labels), modelled on the CPU,
    stays inside the north_star tolerance; fp32 labels (the rejected design) are checked to be worse."""
    from rl_mpc_lanemerging_b200 import synthetic
    p = oracle.horizon_params(H)
    same64 = same32 = tot = 0
    for traffic, kind in (("moderate", "mixed"), ("fast", "mixed"), ("default", "onramp")):
        S = synthetic.make_states(n, traffic, seed=41, kind=kind)
        ref = helpers.oracle_plan_batch(oracle, p, S, H + 1)
        for b in range(n):
            st = helpers.oracle_state(oracle, S, b)
            ob, di, sv = oracle.build_grid(p, st)
            m = oracle.solve_fast_model(p, ob, di, sv, p.t_disc, st.ego_v, st.ego_a)
            m32 = oracle.solve_fast_model(p, ob, di, sv, p.t_disc, st.ego_v, st.ego_a, f32_labels=True)
            assert m["reached_t"] == ref["reached_t"][b]
            if ref["cost"][b] > 0:
                assert abs(m["cost"] - ref["cost"][b]) <= 1e-6 * ref["cost"][b]
            same64 += np.array_equal(m["idx"], ref["idx"][b]); same32 += np.array_equal(m32["idx"], ref["idx"][b]); tot += 1
    assert same64 >= tot - 1 and same64 >= same32


def test_reference_configs_load_unchanged(tmp_path, monkeypatch):
    """Every configs/*.json of the reference loads through Settings.load_from_file (reference config.py:161-170), yields a
    parameter snapshot for the library, and names a task the dispatcher knows (or one that is explicitly out of scope).
    Skipped where the reference tree is not mounted (the GPU box)."""
    import glob
    import os
    import pytest
    from rl_mpc_lanemerging_b200.config import Settings
    from rl_mpc_lanemerging_b200.engine import params_from_settings
    files = sorted(glob.glob("/root/reference/configs/*.json"))
    if not files:
        pytest.skip("reference tree not mounted")
    known = {"ST", "TRAIN_DDPG", "RESUME_DDPG", "EVALUATE_DDPG", "EVALUATE_COMBINED_DDPG", "EVALUATE_COMBINED_DQN"}
    tasks = set()
    for f in files:
        Settings.reset()
        Settings.load_from_file(f)
        p = params_from_settings(Settings)
        assert (p.s_disc, p.t_disc, p.future_s, p.future_t) == (0.05, 0.3, 150.0, 5.0)
        assert p.min_allowed_distance == 5 and p.d_weight == 10.0
        tasks.add(Settings.TASK)
    assert tasks <= known | {"TRAIN_DQN", "EVALUATE_DQN", "RESUME_DQN"}, tasks
    assert {"ST", "TRAIN_DDPG", "EVALUATE_COMBINED_DDPG"} <= tasks
    monkeypatch.chdir(tmp_path)
    Settings.reset(); Settings.LOG_DIR = "unit"
    d = Settings.setup_logging()
    assert os.path.exists(os.path.join(d, "settings.json")) and Settings.FULL_LOG_DIR == d
    Settings.reset()


def test_numa_binding_helper_never_raises(tmp_path, monkeypatch):
    """sharding.bind_to_local_cores: cpulist parsing, and a quiet None where the PCI topology cannot be read (this container has no
    GPU; a box without sysfs PCI entries behaves the same) -- affinity is an optimisation, never an error."""
    import os
    from rl_mpc_lanemerging_b200 import sharding
    assert sharding._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert sharding._parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    cores, why = sharding.bind_to_local_cores(0, 1, sysfs=str(tmp_path))
    assert cores is None and isinstance(why, str)
    assert os.sched_getaffinity(0) == before


def test_graphed_tick_refuses_a_tick_with_host_syncs():
    """GraphedTick only captures the tick that has no host synchronisation left (FUSED_ENV_STEP + SYNC_FREE_TAKEOVER) on an
    auto-resetting environment; anything else is refused before any CUDA call."""
    import types
    from rl_mpc_lanemerging_b200.config import Settings
    from rl_mpc_lanemerging_b200.graphed_tick import GraphedTick
    env = types.SimpleNamespace(B=4, device="cpu", auto_reset=True)
    Settings.reset()
    try:
        with pytest.raises(RuntimeError, match="FUSED_ENV_STEP"):
            GraphedTick(env, agent=None)
        Settings.FUSED_ENV_STEP = Settings.SYNC_FREE_TAKEOVER = True
        env.auto_reset = False
        with pytest.raises(RuntimeError, match="auto-resetting"):
            GraphedTick(env, agent=None)
        env.auto_reset = True
        gt = GraphedTick(env, agent=None)
        with pytest.raises(RuntimeError, match="capture"):
            gt.replay()
    finally:
        Settings.reset()
