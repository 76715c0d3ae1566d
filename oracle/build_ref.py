"""Recipe that compiles the REFERENCE's own native solver into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path may import this.

The reference ships exactly one native component, ``st_cy.pyx`` (Cython).  Its own
build recipe (reference ``setup.py:1-8``) is ``cythonize("st_cy.pyx", language_level=3)``
with the numpy include dir.  We do the same thing here, reading the ``.pyx`` where it
lies under ``/root/reference`` and writing every output (generated C, objects, the
``st_cy*.so`` extension module) under ``oracle/_ref/`` which is git-ignored (it still
travels to the GPU box with the gpurun snapshot).  No reference source is copied into
the repository history.

Usage:  python oracle/build_ref.py [--reference /root/reference]
Returns 0 and prints the path of the built module; returns 0 with a notice when the
reference tree is absent and a prebuilt module already exists (the GPU-box case).
"""
from __future__ import annotations

import argparse
import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def built_module() -> str | None:
    hits = sorted(glob.glob(os.path.join(OUT, "st_cy*.so")))
    return hits[0] if hits else None


def build(reference: str = "/root/reference", force: bool = False) -> str | None:
    os.makedirs(OUT, exist_ok=True)
    pyx = os.path.join(reference, "st_cy.pyx")
    have = built_module()
    if not os.path.exists(pyx):
        return have  # GPU box: only the prebuilt file is available
    if have and not force and os.path.getmtime(have) >= os.path.getmtime(pyx):
        return have
    import numpy

    c_file = os.path.join(OUT, "st_cy.c")
    # 1. Cython -> C  (reference setup.py:5 uses language_level 3, all other directives default)
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file])
    # 2. C -> extension module with the flags distutils would use (-O2, fPIC)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, "st_cy" + ext)
    cmd = [
        "gcc", "-O2", "-fPIC", "-shared", "-fwrapv", "-fno-strict-aliasing",
        "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
        "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
        c_file, "-o", so,
    ]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    path = build(a.reference, a.force)
    print(path if path else "oracle/_ref: reference tree absent and no prebuilt st_cy module")
