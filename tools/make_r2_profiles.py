#!/usr/bin/env python3
"""gpurun_out/ of tools/profile_r2.sh -> profiles/: bench lines, GPU test log, launch list + per-kernel summary, ncu --set full
summaries of K1 (+ k1_traffic.json, which bench.py reports as roofline.traffic) and of the kernels after K1, K1 SASS.
Runs in the build container (ncu -i reads the reports without a GPU)."""
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out") + "/", os.path.join(ROOT, "profiles") + "/"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_ms(unit, val):
    x = float(val.replace(",", ""))
    return x / 1e6 if unit.startswith("n") else x / 1e3 if unit.startswith("u") else x if unit.startswith("m") else x * 1e3


def to_bytes(unit, val):
    return float(val.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


for f in ("r2_bench_line.json", "r2_bench_line_reference.json", "r2_launches.csv"):
    shutil.copy(G + f, P + f)
with open(P + "r2_gpu_tests.txt", "w") as f:
    f.write("# python -m pytest tests -x -q -m gpu --durations=12 on a B200 (tools/profile_r2.sh), final build of round 2\n")
    f.write(open(G + "r2_pytest_gpu.log").read())
plain = json.loads(open(G + "r2_bench_line.json").read().strip().splitlines()[-1])
cells = plain["config"]["cells_per_step"]

# ---- K1
h, u, rows = raw(G + "r2_k1_full.ncu-rep")
v = rows[0]
d = {n: (u[i], v[i]) for i, n in enumerate(h)}
kname = d["Kernel Name"][1] if "Kernel Name" in d else "swg_score_kernel"
ms = to_ms(*d["gpu__time_duration.sum"])
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
with open(P + "r2_k1_ncu_summary.txt", "w") as f:
    f.write(f"# Round 2 (final build): ncu --set full --clock-control none of {kname} on the 18-fragment BENCH launch itself\n")
    f.write("# (python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-driver, -k regex:swg_score_kernel -s 1 -c 1; 1x B200; tools/profile_r2.sh)\n")
    f.write(f"# {cells:.4e} cells in {ms:.1f} ms under ncu ({plain['roofline']['k1_ms_per_launch']:.1f} ms in the plain run: CUDA events) = {cells / ms / 1e9:.2f} TCUPS\n")
    for i, n in enumerate(h):
        if n in keys or n.startswith("smsp__average_warps_issue_stalled") or ("inst_executed_pipe" in n and n.endswith(".sum")):
            f.write(f"{n:95s} {u[i]:16s} {v[i]}\n")
rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
json.dump({"source": "ncu --set full -k regex:swg_score_kernel -s 1 -c 1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-driver (tools/profile_r2.sh), B200, round 2 final build: the 18-fragment bench launch itself",
           "kernel": kname, "cells_per_launch": cells, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch": rd + wr, "gpu_time_ns_under_ncu": ms * 1e6,
           "note": "the column checkpoints for the origin pass (128 per query record x 8 B per library row, written once: 18 x 128 x 588,800 x 8 B = 10.9 GB) and the stripe-boundary ring (8 B per column per stripe with a stripe below: written once, read once); the algorithmic input of K1 (packed query + 8-byte row profiles) is ~5 MB"},
          open(P + "k1_traffic.json", "w"), indent=1)
print(f"K1 {kname}: {ms:.2f} ms under ncu, {(rd + wr) / 1e9:.2f} GB DRAM")

# ---- the kernels after K1
h, u, rows = raw(G + "r2_k2k3k4_full.ncu-rep")
ki = h.index("Kernel Name")
names = [r[ki][:34] for r in rows]
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
with open(P + "r2_k2_k3_k4_ncu_summary.txt", "w") as f:
    f.write("# Round 2 (final build): ncu --set full of the kernels after K1 on the bench launch (tools/profile_r2.sh; columns = captured launches in order;\n")
    f.write("# ncu serialises launches that overlap in a plain run: swg_score32_kernel (K2w) and the first swg_origin_block_kernel (common triples) run side by side there,\n")
    f.write("# the second swg_origin_block_kernel is the launch over the hot blocks of the 11 wrapped triples, started from K2w's int32 checkpoints)\n")
    f.write(f"{'Kernel Name':90s} {names}\n")
    for i, n in enumerate(h):
        if n in want or n.startswith("smsp__average_warps_issue_stalled"):
            f.write(f"{n + ' [' + u[i] + ']':90s} {[r[i] for r in rows]}\n")
for r in rows:
    print(f"  {r[ki][:44]:44s} {r[h.index('gpu__time_duration.sum')]} {u[h.index('gpu__time_duration.sum')]}")

# ---- launch list
rows = [r for r in csv.reader(open(G + "r2_launches.csv")) if len(r) > 10]
h = rows[0]
ki, vi, ui, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += to_ms(r[ui], r[vi])
step = sum(m for k, (n, m) in agg.items() if "dpx" not in k and "at::" not in k)
with open(P + "r2_launch_summary.txt", "w") as f:
    f.write("# Round 2 (final build): ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-driver` (1x B200), 3 steps + the DPX microbenchmark\n")
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400  -> profiles/r2_launches.csv (per-launch times are cold-cache and serialised: compare SHARES)\n")
    f.write("# dpx_* = peak microbenchmark, at::/elementwise = L2 flush (outside the step).  swg_score32_kernel (K2w) and the two swg_origin_block_kernel launches run\n# CONCURRENTLY in a normal run; ncu serialises them.\n")
    f.write(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s}  share of step\n")
    for k, (n, m) in sorted(agg.items(), key=lambda x: -x[1][1]):
        sh = "n/a" if ("dpx" in k or "at::" in k) else f"{100 * m / step:5.1f}%"
        f.write(f"{k[:70]:70s} {n:8d} {m:10.3f}  {sh:>12s}\n")
print(open(P + "r2_launch_summary.txt").read())

# ---- SASS of the K1 instantiation the bench runs
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "repeatmasker_b200", "librmgpu.so")], capture_output=True, text=True).stdout
out, keep = [], False
for ln in sass.splitlines():
    if "Function :" in ln:
        keep = "swg_score_kernelILi16ELb0ELin30ELin6ELin5E" in ln
    if keep and not ln.lstrip().startswith("/* 0x"):
        out.append(ln.split("/* 0x")[0].rstrip())
open(P + "r2_k1_swg_score_kernel.sass", "w").write("\n".join(out) + "\n")
print("SASS lines", len(out), {m: sum(m in l for l in out) for m in ("VIADDMNMX.S16x2", "VIADD.16x2", "PRMT", "SHFL.UP", "VIMNMX3", "IMAD")})
