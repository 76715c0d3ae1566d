"""Import the UNMODIFIED reference Python modules (THIS CONTAINER ONLY).

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py to generate the
committed golden vectors and by the optional ``-m "not gpu"`` cross-checks that skip
when /root/reference is absent.  /root/reference does not exist on the GPU box, so
nothing that runs there may call load().

The reference needs SUMO/traci, cvxopt and matplotlib at import time
(control.py:17-23, st.py:4-5,16-17, prediction.py:2); none is on the hot path, so we
insert empty stub modules for them, exactly as SURVEY.md Appendix A describes.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE = "/root/reference"
_loaded = None


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE, "st.py"))


def load(config: str | None = "configs/st_moderate.json"):
    """Returns a namespace with the reference modules: Settings, st, st_cy, prediction, control, dqn."""
    global _loaded
    if _loaded is not None:
        if config:
            _loaded.Settings.load_from_file(os.path.join(REFERENCE, config))
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    from . import build_ref

    so = build_ref.build(REFERENCE)
    assert so, "could not build oracle/_ref/st_cy"
    os.environ.setdefault("SUMO_HOME", "/nonexistent")
    for name in ("traci", "cvxopt", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["cvxopt"].solvers = types.SimpleNamespace(options={})
    sys.modules["cvxopt"].matrix = None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # the reference uses flat top-level module names (config, st, control ...): make sure
    # our own package (which mirrors those names INSIDE rl_mpc_lanemerging_b200) never shadows them
    for name in ("config", "st", "st_cy", "prediction", "control", "dqn", "stats", "rl", "sumo"):
        assert name not in sys.modules, f"top-level module {name} already imported"
    sys.path[:0] = [os.path.dirname(so), REFERENCE]
    try:
        from config import Settings  # type: ignore

        if config:
            Settings.load_from_file(os.path.join(REFERENCE, config))
        import st, st_cy, prediction, control, dqn  # type: ignore  # noqa: E401
    finally:
        sys.path.remove(REFERENCE)
        sys.path.remove(os.path.dirname(so))
    _loaded = types.SimpleNamespace(Settings=Settings, st=st, st_cy=st_cy, prediction=prediction,
                                    control=control, dqn=dqn)
    return _loaded


def load_st_cy():
    """The compiled reference solver alone (works on the GPU box from the prebuilt oracle/_ref)."""
    from . import build_ref
    import importlib.util

    so = build_ref.build(REFERENCE)
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("st_cy", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
