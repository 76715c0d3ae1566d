#!/usr/bin/env python
"""Turn the raw ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarise_profiles.py r01a r01        # gpurun_out/r01a_*  ->  profiles/r01_*
"""
import csv
import os
import subprocess
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
os.makedirs("profiles", exist_ok=True)


def short(name):
    return name.split("(")[0].replace("void ", "").strip()


for H in (50, 17):
    f = f"gpurun_out/{src}_launches_h{H}.csv"
    if os.path.exists(f):
        rows = [r for r in csv.reader(open(f)) if len(r) > 14 and r[0].isdigit()]
        agg = defaultdict(lambda: [0, 0.0])
        order = []
        for r in rows:
            k = short(r[4])
            if k not in agg:
                order.append(k)
            v = float(r[14].replace(",", ""))
            unit = r[13]
            us = v / 1000.0 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1000.0 if unit in ("msecond", "ms") else v)
            agg[k][0] += 1; agg[k][1] += us
        tot = sum(v[1] for v in agg.values())
        with open(f"profiles/{dst}_launches_h{H}.txt", "w") as o:
            o.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none   python bench.py --steps 2 --warmup 3 --horizon {H} --no-cpu-baseline\n")
            o.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  (the torch fill/zero kernels are the L2 flush\n")
            o.write("# and buffer initialisation of bench.py, outside the timed region)\n")
            o.write(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>10s}\n")
            for k in sorted(order, key=lambda k: -agg[k][1]):
                n, us = agg[k]
                o.write(f"{k[:70]:70s} {n:8d} {us:12.1f} {100 * us / tot:6.1f}% {us / n:10.1f}\n")
        print("wrote", o.name)
    for kern, tag, regex in (("fast32", "fast32", "fast32"), ("fast", "fast_pull", "fast_pull"), ("fallback", "fallback_fast_pull", "fast_pull"),
                             ("predict", "predict_layers", "predict_layers"), ("rasterise", "rasterise", "rasterise"), ("rasterise_rows", "rasterise_rows", "rasterise_rows"), ("dense32", "dense_fast32_f32", "fast32")):
        rep = f"gpurun_out/{src}_{kern}_h{H}.ncu-rep"
        if not os.path.exists(rep):
            continue
        out = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "--so", f"gpurun_out/{src}_lib.so", "--top", "30"],
                             capture_output=True, text=True).stdout
        sass = subprocess.run([sys.executable, "tools/ncu_sass.py", rep, "--min", "2.0"], capture_output=True, text=True).stdout
        with open(f"profiles/{dst}_{tag}_h{H}.txt", "w") as o:
            o.write(f"# ncu --set full --clock-control none --import-source on -k regex:{regex} -c 1   python tools/prof_run.py {H} 4096 fast\n")
            o.write("# key raw metrics, per-source-line attribution (tools/ncu_lines.py), then executed instructions by opcode (tools/ncu_sass.py)\n")
            o.write(out)
            o.write(sass)
        print("wrote", o.name)
# DRAM traffic of the dominant kernel per launch, for bench.py's roofline.traffic
import json, re
traffic = {}
for H in (50, 17):
    rep = f"gpurun_out/{src}_fast32_h{H}.ncu-rep"
    kname = "fast32_kernel"
    if not os.path.exists(rep):
        rep, kname = f"gpurun_out/{src}_fast_h{H}.ncu-rep", "fast_pull_kernel"
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def get(name):
        i = hdr.index(name)
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(vals[i]) * mult
    traffic[str(H)] = {"episodes": 4096, "kernel": kname, "dram_bytes_read": get("dram__bytes_read.sum"),
                       "dram_bytes_write": get("dram__bytes_write.sum"), "source": f"profiles/{dst}_{'fast32' if kname == 'fast32_kernel' else 'fast_pull'}_h{H}.txt"}
rep = f"gpurun_out/{src}_dense32_h50.ncu-rep"
if os.path.exists(rep):                     # K2 on dense fp32 grids (tools/prof_run.py 50 1024 dense32)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def getd(name):
        i = hdr.index(name)
        return float(vals[i]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    traffic["dense32_50"] = {"episodes": 1024, "kernel": "fast32_kernel<FastDenseProv<float>>", "dram_bytes_read": getd("dram__bytes_read.sum"),
                             "dram_bytes_write": getd("dram__bytes_write.sum"), "source": f"profiles/{dst}_dense_fast32_f32_h50.txt"}
if traffic:
    json.dump(traffic, open(f"profiles/{dst}_traffic.json", "w"), indent=1)
    print("wrote", f"profiles/{dst}_traffic.json")
for H in (50, 17):
    f = f"gpurun_out/{src}_bench_h{H}.json"
    if os.path.exists(f) and os.path.getsize(f):
        open(f"profiles/{dst}_bench_h{H}.json", "w").write(open(f).read())
        print("copied", f)
