import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_oracle
    cpu_oracle.build()
    return cpu_oracle


@pytest.fixture
def emulated_engine(monkeypatch):
    """The Python host layer on the EMULATED library with CPU tensors (tests/emu/emu_engine.py): st.get_engine() hands out
    EmuBackedEngine instances for the duration of the test."""
    from rl_mpc_lanemerging_b200 import st
    from rl_mpc_lanemerging_b200.config import Settings
    from tests.emu.emu_engine import EmuBackedEngine
    st.refresh_engine()
    Settings.reset()
    from rl_mpc_lanemerging_b200 import st_cy
    monkeypatch.setattr(st, "MpcEngine", EmuBackedEngine)
    monkeypatch.setattr(st_cy, "MpcEngine", EmuBackedEngine)
    monkeypatch.setattr(st_cy, "_engines", {})
    yield EmuBackedEngine
    st.refresh_engine()
    Settings.reset()
