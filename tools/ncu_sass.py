#!/usr/bin/env python
"""Executed-instruction profile of one ncu capture at SASS level (no GUI): every instruction of the kernel that was executed at
least `--min` (fraction of the most-executed one) times, in address order, with warp-level count, active lanes and stall samples.

    python tools/ncu_sass.py gpurun_out/x.ncu-rep [--min 0.2] [--grep ATOMS]

Reading it: instructions of the per-node path share one count; dividing the kernel's total by it gives warp instructions per
32-cell chunk that holds nodes."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--min", type=float, default=0.2)
ap.add_argument("--grep", default=None)
a = ap.parse_args()
kernel, data, _ = sass_rows(a.rep)
total = sum(d["inst"] for d in data)
top = max(d["inst"] for d in data)
print(f"# {kernel}\n# warp instructions {total}, most executed {top}; listed: >= {a.min:.2f} of it")
hist = {}
for d in data:
    op = d["sass"].split()[0] if not d["sass"].startswith("@") else d["sass"].split()[1]
    hist[op.split(".")[0]] = hist.get(op.split(".")[0], 0) + d["inst"]
    if d["inst"] >= a.min * top and (a.grep is None or a.grep in d["sass"]):
        lanes = d["thr"] / d["inst"] if d["inst"] else 0
        print(f"{d['off']:6x} {d['inst'] / 1e6:9.2f}M  lanes {lanes:4.1f}  samples {d['samples']:6d}  {d['sass']}")
print("# by opcode (share of all warp instructions):")
for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:25]:
    print(f"#   {k:12s} {v / total * 100:5.1f} %")
