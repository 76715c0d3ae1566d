"""MpcEngine on the EMULATED library (tests/emu/libmpc_emu_api.so) with CPU tensors -- TEST INFRASTRUCTURE ONLY.

Lets the Python host layer (st, dqn, ddpg, merge_gym, control, trainer) run in a container without a GPU: every kernel is the
library's own source executed on fibers (cuda_emu.h), "device" tensors are CPU tensors.  The product class
(rl_mpc_lanemerging_b200.engine.MpcEngine) is unchanged and still refuses to start without a CUDA device; nothing in the
package imports this module.  Use the `emulated_engine` fixture of tests/conftest.py.
"""
import contextlib
import ctypes as C

import torch

from rl_mpc_lanemerging_b200 import _lib
from rl_mpc_lanemerging_b200.engine import MpcEngine
from tests.emu import emu_api


class EmuBackedEngine(MpcEngine):
    _pin_host = False

    def __init__(self, params=None, device=0, max_batch: int = 64, nmax: int = 32):   # noqa: D401 -- no super().__init__: no CUDA
        self.lib = emu_api.lib()
        self.device = torch.device("cpu")
        self.dev_index = 0
        if params is None:
            params = _lib.MpcParams()
            self.lib.mpc_default_params(C.byref(params))
        self.params = params
        self.max_batch, self.nmax = int(max_batch), int(nmax)
        h = C.c_void_p()
        emu_api.check(self.lib.mpc_create(C.byref(self.params), 0, self.max_batch, self.nmax, C.byref(h)))
        self.h = h
        nt, ns = C.c_int(), C.c_int()
        emu_api.check(self.lib.mpc_grid_dims(self.h, C.byref(nt), C.byref(ns)))
        self.num_t, self.num_s_max = nt.value, ns.value
        self.num_s_stride = int(self.lib.mpc_grid_stride(self.h))
        self._pinned = {}

    def _stream(self):
        return None

    @staticmethod
    def _is_dev(t) -> bool:
        return not t.is_cuda

    def _device_ctx(self):
        return contextlib.nullcontext()
