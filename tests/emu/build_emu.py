"""Builds tests/emu/libmpc_emu.so: the library's kernel sources compiled by g++ against the fiber shim (cuda_emu.h).
TEST INFRASTRUCTURE ONLY -- see cuda_emu.h for what the emulation does and does not show."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "rl_mpc_lanemerging_b200", "csrc")
LIB = os.path.join(HERE, "libmpc_emu.so")
DEPS = [os.path.join(HERE, f) for f in ("cuda_emu.h", "emu_harness.cpp")] + \
       [os.path.join(CSRC, f) for f in ("mpc_common.cuh", "mpc_solve_common.cuh", "mpc_derive.h", "mpc_predict.cu", "mpc_reach.cu",
                                        "mpc_fast.cu", "mpc_fast32.cuh")] + [os.path.join(ROOT, "include", "mpcb200.h")]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in DEPS):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-DMPC_HOST_EMU=1",
                               "-I", HERE, "-I", os.path.join(ROOT, "include"), "-Wno-unused-function",
                               os.path.join(HERE, "emu_harness.cpp"), "-o", LIB])
    return LIB


API_LIB = os.path.join(HERE, "libmpc_emu_api.so")
API_SOURCES = ["mpc_api.cu", "mpc_predict.cu", "mpc_solve.cu", "mpc_fast.cu", "mpc_qp.cu", "mpc_reach.cu"]


def build_api(force: bool = False) -> str:
    """The WHOLE library (C ABI included) compiled by g++ against the shim: every .cu its own translation unit, the CUDA
    runtime replaced by the stubs of cuda_emu.h ("device" memory = host memory)."""
    deps = [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "mpcb200.h")] + \
           [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if force or not os.path.exists(API_LIB) or os.path.getmtime(API_LIB) < max(os.path.getmtime(d) for d in deps):
        objs = []
        for f in API_SOURCES:
            o = os.path.join(HERE, "_" + f.replace(".cu", ".emu.o"))
            subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-c", "-DMPC_HOST_EMU=1",
                                   "-I", HERE, "-I", os.path.join(ROOT, "include"), "-Wno-unused-function",
                                   os.path.join(CSRC, f), "-o", o])
            objs.append(o)
        subprocess.check_call(["g++", "-shared", "-o", API_LIB] + objs)
        for o in objs:
            os.remove(o)
    return API_LIB


if __name__ == "__main__":
    print(build_api(force=True))
    print(build(force=True))
