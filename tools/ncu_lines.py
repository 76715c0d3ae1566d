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
        def num(name):
            try:
                return int(r[col[name]])
            except Exception:
                return 0
        data.append(dict(addr=int(r[0], 16), sass=r[col["Source"]], inst=int(r[col["Instructions Executed"]]),
                         thr=int(r[col["Thread Instructions Executed"]]), samples=int(r[col["# Samples"]]),
                         wf=num("L1 Wavefronts Shared"), wf_ideal=num("L1 Wavefronts Shared Ideal"),
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


KEY_METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
               "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
               "sm__inst_executed.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
               "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
               "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
               "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
               "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    res = []
    for i, h in enumerate(hdr):
        if h in KEY_METRICS:
            res.append((h, vals[i], units[i]))
        elif "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                if float(vals[i]) >= 0.3:
                    res.append(("stall " + h.split("issue_stalled_")[1].split("_per")[0] + " (warps per issue)", vals[i], ""))
            except ValueError:
                pass
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--so", default="rl_mpc_lanemerging_b200/libmpcb200.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--src", default="rl_mpc_lanemerging_b200/csrc")
    a = ap.parse_args()
    for name, val, unit in raw_metrics(a.rep):
        print(f"  {name} = {val} {unit}")
    kernel, data, hdr = sass_rows(a.rep)
    lm = line_map(a.so, kernel)
    agg = defaultdict(lambda: [0, 0, 0, 0, 0])
    for d in data:
        key = lm.get(d["off"]) or ("?", 0)
        agg[key][0] += d["inst"]; agg[key][1] += d["samples"]; agg[key][2] += d["thr"]; agg[key][3] += d["wf"]; agg[key][4] += d["wf_ideal"]
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    twf = sum(v[3] for v in agg.values()); twfi = sum(v[4] for v in agg.values())
    print(f"kernel: {kernel}\nwarp-instructions {ti}  samples {ts}  shared wavefronts {twf} (ideal {twfi})  mapped offsets {sum(1 for d in data if d['off'] in lm)}/{len(data)}")
    src_cache = {}
    for (f, line), (inst, samp, thr, wf, wfi) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        path = os.path.join(a.src, f)
        if path not in src_cache:
            src_cache[path] = open(path).read().splitlines() if os.path.exists(path) else []
        text = src_cache[path][line - 1].strip()[:100] if 0 < line <= len(src_cache[path]) else ""
        print(f"{100 * samp / ts:5.1f}% samples {100 * inst / ti:5.1f}% inst  lanes {thr / max(inst, 1):4.1f}  smem-wf {100 * wf / max(twf, 1):4.1f}% (x{wf / max(wfi, 1):.1f})  {f}:{line}  {text}")


if __name__ == "__main__":
    main()
