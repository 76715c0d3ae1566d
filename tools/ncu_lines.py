#!/usr/bin/env python
"""Attribute an ncu capture to CUDA source lines (no GUI).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--so rl_mpc_lanemerging_b200/libmpcb200.so] [--top 40]

ncu's `--page source --csv` lists per-SASS-instruction counters; nvdisasm -g gives the source line of
every SASS instruction of the same kernel in the shipped .so (compiled with -lineinfo).  The two are
joined on the instruction offset inside the kernel.
"""
from __future__ import annotations

import argparse
import csv
import os
import re
import subprocess
import tempfile
from collections import defaultdict


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernel = rows[0][1]
    hdr = rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        data.append(dict(addr=int(r[0], 16), sass=r[col["Source"]], inst=int(r[col["Instructions Executed"]]),
                         thr=int(r[col["Thread Instructions Executed"]]), samples=int(r[col["# Samples"]]),
                         stalls={h: int(r[i] or 0) for h, i in col.items() if h.startswith("stall_") and r[i] not in ("", "-")}))
    base = data[0]["addr"]
    for d in data:
        d["off"] = d["addr"] - base
    return kernel, data, hdr


def line_map(so, kernel_name):
    """offset -> (file, line) for the kernel whose demangled name matches."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    short = kernel_name.split("(")[0].split("<")[0].split()[-1]
    best = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line, m = None, None, {}
        want_tpl = kernel_name
        for ln in txt.splitlines():
            fm = re.match(r"\s*\.text\.(\S+):", ln)
            if fm:
                if cur_fn and m:
                    best[cur_fn] = m
                cur_fn, m = fm.group(1), {}
                continue
            lm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if lm:
                cur_line = (os.path.basename(lm.group(1)), int(lm.group(2)))
                continue
            im = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
            if im and cur_fn:
                m[int(im.group(1), 16)] = cur_line
        if cur_fn and m:
            best[cur_fn] = m
    # pick the mangled function by demangling
    cands = [fn for fn in best if short in fn]
    dem = {}
    for fn in cands:
        d = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip()
        dem[fn] = d
    norm = lambda s: re.sub(r"\s+|\(bool\)|\(int\)", "", s).replace("true", "1").replace("false", "0")
    target = norm(kernel_name)
    for fn, d in dem.items():
        if norm(d).startswith(target[:len(norm(d))]) or norm(d) == target:
            if norm(d).split("(")[0] == target.split("(")[0]:
                return best[fn]
    return best[cands[0]] if cands else {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--so", default="rl_mpc_lanemerging_b200/libmpcb200.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--src", default="rl_mpc_lanemerging_b200/csrc")
    a = ap.parse_args()
    kernel, data, hdr = sass_rows(a.rep)
    lm = line_map(a.so, kernel)
    agg = defaultdict(lambda: [0, 0, 0])
    for d in data:
        key = lm.get(d["off"]) or ("?", 0)
        agg[key][0] += d["inst"]; agg[key][1] += d["samples"]; agg[key][2] += d["thr"]
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"kernel: {kernel}\nwarp-instructions {ti}  samples {ts}  mapped offsets {sum(1 for d in data if d['off'] in lm)}/{len(data)}")
    src_cache = {}
    for (f, line), (inst, samp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        path = os.path.join(a.src, f)
        if path not in src_cache:
            src_cache[path] = open(path).read().splitlines() if os.path.exists(path) else []
        text = src_cache[path][line - 1].strip()[:100] if 0 < line <= len(src_cache[path]) else ""
        print(f"{100 * samp / ts:5.1f}% samples {100 * inst / ti:5.1f}% inst  lanes {thr / max(inst, 1):4.1f}  {f}:{line}  {text}")


if __name__ == "__main__":
    main()
