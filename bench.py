#!/usr/bin/env python3
"""bench.py -- the hot path on BASELINE.json's configuration 2:
   1 Mbp synthetic sequence with implanted 10-30 %-diverged copies of a 500-family consensus library, both strands.

A "step" = one pass of the search (K0 streams/profile, K1 score, K2 origin, K3 traceback, K4 rows/counts, K5 '*' /
masklevel, records back) over a batch of SimpleBatcher fragments (SimpleBatcher.pm:642-720 arithmetic: 60 kbp, 2 kbp
overlap) against the whole library: Q*L*2 ~ 1.2e12 DP cells per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload config2|config3] [--scaling weak|strong] [--genome-bp B] [--families F]

N > 1 is launched by torchrun, one rank per GPU.  The units are SimpleBatcher fragments (independent), fragment b of ONE
genome goes to rank b mod N, the library is replicated, and every step ends with the host-side gather of all hit
records on rank 0 (a gloo group: no data-path collective on the GPUs; NCCL carries only the barrier and the
max-reduce of the timing).  Default: weak scaling (the genome grows with N: N x 1 Mbp, per-GPU work fixed).
--scaling strong fixes the genome (use with --workload config3: BASELINE configs[2], 46 Mbp x 1,500 families, -s
parameters, whose 794 fragments are the reference's own unit of parallelism).

--impl reference is the CPU arm: the UNMODIFIED reference driver from baseline/_ref,
`RepeatMasker -engine crossmatch -pa <host cores> -lib lib.fa genome.fa`, with this repository's restated crossmatch
(oracle/_build/cross_match; Phil Green's cross_match is not in the reference tree and cannot be built here) as its
engine, on a bounded sub-genome of the same workload; GCUPS = cells of its search stage / wall time of the whole run.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "GCUPS (1e9 Smith-Waterman-Gotoh cell updates/s, Q*L*2 per stage) searched vs restated crossmatch on host cores"
WORKLOADS = {
    # BASELINE configs[1] / configs[2]; gap recipes RepeatMasker:2146-2366 (default speed / -s differ in minmatch only,
    # which the exhaustive search ignores)
    "config2": dict(seed=1001, families=500, genome_bp=1_000_000, matrix="20p41g.matrix",
                    gap=dict(gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5, minscore=225, masklevel=101, minmatch=9, bandwidth=14),
                    text="BASELINE configs[1]: 1 Mbp synthetic genome (45% implanted 10-30%-diverged copies) x 500-family "
                         "synthetic consensus library, both strands, SimpleBatcher 60 kbp fragments / 2 kbp overlap"),
    "config3": dict(seed=1002, families=1500, genome_bp=46_000_000, matrix="20p41g.matrix",
                    gap=dict(gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5, minscore=225, masklevel=101, minmatch=8, bandwidth=14),
                    text="BASELINE configs[2]: 46 Mbp chr21-sized synthetic genome x 1,500-family library, -s parameters, "
                         "SimpleBatcher 60 kbp fragments / 2 kbp overlap"),
}


def make_genome(W, genome_bp, sseqs):
    """One sequence, generated 1 Mbp at a time (the first Mbp is the same for every size)."""
    from repeatmasker_b200 import synth

    parts, left, k = [], genome_bp, 0
    while left > 0:
        g, _ = synth.make_genome(W["seed"] + 7919 * k, min(left, 1_000_000), sseqs, coverage=0.45)
        parts.append(g)
        left -= 1_000_000
        k += 1
    return "".join(parts)


def workload(args, rank, world):
    """Library (replicated) and this rank's fragments of the one genome: fragment b -> rank b mod world."""
    from repeatmasker_b200 import sharding, synth

    W = WORKLOADS[args.workload]
    snames, sseqs = synth.make_library(W["seed"], n_families=args.families or W["families"])
    bp = args.genome_bp or W["genome_bp"]
    if args.scaling == "weak":
        bp *= world
    genome = make_genome(W, bp, sseqs)
    qn, qs = synth.fragment(genome, "synth0", frag=60000, overlap=2000)
    mine = sharding.my_fragments(len(qs), rank, world)
    return W, snames, sseqs, [qn[i] for i in mine], [qs[i] for i in mine], len(qs), bp


def matrix_file(tmpdir, name):
    import util

    return util.write_matrix(name, os.path.join(tmpdir, name))


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------- CPU arms
def oracle_sample(W, tmpdir, snames, sseqs, qn, qs, procs, stride):
    """The raw oracle (no driver) on a bounded sample: the first fragment against every `stride`-th family (family
    popularity falls along the library, so a constant stride keeps the hit density representative), dealt round-robin to
    `procs` oracle processes.  Returns (cells, seconds, description)."""
    import util
    from repeatmasker_b200 import synth

    qpath = os.path.join(tmpdir, "cpu_q.fa")
    synth.write_fasta(qpath, [qn[0]], [qs[0]])
    m = matrix_file(tmpdir, W["matrix"])
    picked = list(range(0, len(sseqs), stride))
    running, cells = [], 0.0
    for t in range(procs):
        sub = picked[t::procs]
        if not sub:
            continue
        sp = os.path.join(tmpdir, f"cpu_s{t}.fa")
        synth.write_fasta(sp, [snames[i] for i in sub], [sseqs[i] for i in sub])
        cells += 2.0 * len(qs[0]) * sum(len(sseqs[i]) for i in sub)
        running.append(util.oracle_argv(qpath, sp, m, **W["gap"]))
    t0 = time.perf_counter()
    ps = [subprocess.Popen(a, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for a in running]
    for p in ps:
        p.wait()
    dt = time.perf_counter() - t0
    desc = (f"fragment 1 ({len(qs[0])} bp) x {len(picked)} of {len(sseqs)} families (every {stride}th), both strands, "
            f"{len(ps)} oracle process(es): restated crossmatch of this repository (AVX2 score prefilter + scalar origin pass), no driver")
    return cells, dt, desc


def harness():
    h = os.path.join(ROOT, "baseline", "_ref")
    return h if os.path.exists(os.path.join(h, "env.sh")) and shutil.which("perl") else None


def driver_run(engine_name, workdir, pa, extra="", frag=0):
    """`RepeatMasker -engine <engine_name> -pa <pa> -lib lib.fa genome.fa` from baseline/_ref in workdir; wall seconds."""
    H = harness()
    if frag:
        extra += f" -frag {frag}"
    cmd = f"source {H}/env.sh && cd {workdir} && perl $RM_HOME/RepeatMasker -engine {engine_name} -lib lib.fa -gc 41 -pa {pa} {extra} genome.fa"
    t0 = time.perf_counter()
    res = subprocess.run(["bash", "-c", cmd], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if res.returncode:
        raise SystemExit("RepeatMasker failed: " + res.stdout[-1500:] + res.stderr[-1500:])
    return dt


def driver_inputs(W, workdir, snames, sseqs, genome):
    from repeatmasker_b200 import synth

    os.makedirs(workdir, exist_ok=True)
    synth.write_fasta(os.path.join(workdir, "lib.fa"), snames, sseqs)
    synth.write_fasta(os.path.join(workdir, "genome.fa"), ["synth0"], [genome], width=60)
    for f in os.listdir(workdir):
        if f.startswith("genome.fa.") or f.startswith("RM_"):
            p = os.path.join(workdir, f)
            shutil.rmtree(p) if os.path.isdir(p) else os.remove(p)


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from repeatmasker_b200 import synth

    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    cores = os.cpu_count() or 1
    W = WORKLOADS[args.workload]
    snames, sseqs = synth.make_library(W["seed"], n_families=args.families or W["families"])
    sub_bp = args.cpu_genome_bp
    genome = make_genome(W, sub_bp, sseqs)
    qn, qs = synth.fragment(genome, "synth0", frag=60000, overlap=2000)
    cells = 2.0 * sum(len(x) for x in qs) * sum(len(x) for x in sseqs)
    H = harness()
    # the driver's unit of parallelism is the batch: -frag is chosen so that every host core has one (the cell count
    # does not depend on it)
    frag = max(5000, -(-sub_bp // cores) + 2000)
    line = {"metric": METRIC, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int16 (AVX2 prefilter) / int32, f64 adjustment",
            "data": "synthetic", "impl": "reference", "gpu_launches": 0}
    with tempfile.TemporaryDirectory() as d:
        if H:
            tot = 0.0
            for it in range(min(args.warmup, 1) + args.steps):
                driver_inputs(W, d, snames, sseqs, genome)
                dt = driver_run("crossmatch", d, cores, frag=frag)
                if it >= min(args.warmup, 1):
                    tot += dt
            gcups = cells * args.steps / tot / 1e9
            sample = (f"the unmodified reference driver, RepeatMasker -engine crossmatch -pa {cores} -lib (500 families) on the first "
                      f"{sub_bp} bp of the workload's genome (-frag {frag}: one batch per core), restated crossmatch of this repository as "
                      f"cross_match; cells of the search stage / wall time of the whole run")
            kind = "port"
            ms = 1e3 * tot / args.steps
        else:
            tot_c = tot_s = 0.0
            for it in range(min(args.warmup, 1) + args.steps):
                c, dt, sample = oracle_sample(W, d, snames, sseqs, qn, qs, cores, 4)
                if it >= min(args.warmup, 1):
                    tot_c += c
                    tot_s += dt
            gcups, kind, ms = tot_c / tot_s / 1e9, "port", 1e3 * tot_s / args.steps
        c1, dt1, desc1 = oracle_sample(W, d, snames, sseqs, qn, qs, 1, 16)
    line.update({"value": gcups, "ms_per_step": ms,
                 "config": config_dict(args, W, snames, sseqs, len(qs), sum(len(x) for x in qs), 1, sub_bp),
                 "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample,
                                  "raw_oracle_one_core_gcups": c1 / dt1 / 1e9, "raw_oracle_sample": desc1,
                                  "note": "no timing against Phil Green's seeded, banded cross_match exists: it is not in the "
                                          "reference tree; this arm evaluates every cell like the GPU engine does"},
                 "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def traffic_from_profile(cells_step):
    """dram__bytes_read.sum + dram__bytes_write.sum of K1 from the committed `ncu --set full` capture
    (profiles/k1_traffic.json, taken on this workload), per launch; None when the capture is for another size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        if abs(t["cells_per_launch"] - cells_step) / cells_step < 0.02:
            return t["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def config_dict(args, W, snames, sseqs, n_frag, Q, world, genome_bp):
    L = sum(len(x) for x in sseqs)
    return {"workload": W["text"], "genome_bp": genome_bp, "families": len(sseqs), "library_bp": L, "fragments": n_frag,
            "query_bp_searched": Q, "cells_per_step": 2.0 * Q * L, "matrix": W["matrix"][:-7],
            "gap": [W["gap"]["gap_init"], W["gap"]["ins_gap_ext"], W["gap"]["del_gap_ext"]], "minscore": W["gap"]["minscore"],
            "seed": W["seed"],
            "parallelism": f"one genome, fragment b -> rank b mod {world}, library replicated, host-side gather of the hit records on "
                           f"rank 0 inside the timed region, no data-path collective",
            "l2": "L2 flushed between timed steps (256 MiB write)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--genome-bp", type=int, default=0, help="per-GPU (weak) or total (strong) genome size; 0 = the workload's")
    ap.add_argument("--families", type=int, default=0)
    ap.add_argument("--cpu-genome-bp", type=int, default=120_000, help="sub-genome of the --impl reference driver run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-driver", action="store_true", help="skip the RepeatMasker -engine gpu driver run of the bench line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from repeatmasker_b200 import build as rm_build
    from repeatmasker_b200 import engine, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")  # the host-side gather of the records
    if rank == 0:
        rm_build.build()
    if world > 1:
        dist.barrier()

    W, snames, sseqs, qn, qs, n_frag, genome_bp = workload(args, rank, world)
    tmp = tempfile.TemporaryDirectory()
    m = matrix_file(tmp.name, W["matrix"])
    params = engine.make_params(device=local, **W["gap"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    mine = sharding.my_fragments(n_frag, rank, world)

    sess = engine.Session(local)
    sess.set_matrix(m)
    sess.set_library(sseqs, snames)
    sess.set_queries(qs, qn)
    n_hits = [0]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def finish(rec):
        """End of a step: all records on rank 0, in fragment order (RepeatMasker:994-1017 concatenates .cat files so).
        On one GPU they already are where the caller wants them: C records in host memory, as the XS glue reads them."""
        st = rec.stats
        if world > 1:
            hits = rec.to_dicts()
            per_frag = sharding.split_hits_by_fragment(hits, qn)
            gathered = [None] * world if rank == 0 else None
            dist.gather_object((mine, per_frag), gathered, dst=0, group=host_group)
            if rank == 0:
                out = [None] * n_frag
                for ids, hs in gathered:
                    for f, h in zip(ids, hs):
                        out[f] = h
                assert all(h is not None for h in out)
                n_hits[0] = sum(len(h) for h in out)
        else:
            n_hits[0] = len(rec)
        return st

    def timed(fn, steps):
        tot, stats = 0.0, []
        for _ in range(steps):
            flush.fill_(1)
            sync()
            t0 = time.perf_counter()
            st = fn()
            sync()
            tot += time.perf_counter() - t0
            stats.append(st)
        return tot, stats

    def step_resident():
        return finish(sess.search_records(params))

    def step_e2e():  # host buffers in, host records out: library + batch upload, search, records (and the gather)
        sess.set_library(sseqs, snames)
        sess.set_queries(qs, qn)
        return finish(sess.search_records(params))

    timed(step_resident, args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    t_res, st_res = timed(step_resident, args.steps)
    timed(step_e2e, 1)
    t_e2e, st_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    cells_step = st_res[0]["cells"]
    tt = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
    cc = torch.tensor([cells_step, float(sum(len(x) for x in qs))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
    t_res_max, t_e2e_max = tt.tolist()
    cells_all, Q_all = cc.tolist()
    value = cells_all * args.steps / t_res_max / 1e9
    e2e = cells_all * args.steps / t_e2e_max / 1e9

    if rank == 0:
        k1_ms = sum(s["ms_score"] for s in st_res) / len(st_res)
        peak_cups, dpx_ips = engine.measure_dpx_peak(local)
        achieved = cells_step / (k1_ms * 1e-3) / 1e9
        issue_bound = dpx_ips * 32 * 2 / 5 / 1e9  # measured VIADDMNMX.S16x2 rate x 32 lanes x 2 cells / 5 instructions
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        Q = sum(len(x) for x in qs)
        s0 = st_res[0]
        line = {
            "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_res_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "int16x2 (DPX) score pass, int32 origin/traceback, f64 complexity adjustment", "data": "synthetic",
            "config": config_dict(args, W, snames, sseqs, n_frag, int(Q_all), world, genome_bp),
            "mbp_per_s": Q_all * args.steps / t_res_max / 1e6,
            "e2e": {"value": e2e, "unit": "GCUPS",
                    # counted by the library: packed 4-bit batch + offsets; library row codes, 1 B/base codes, group tables
                    "h2d_bytes_per_step": int(st_e2e[0]["h2d_bytes"]) + int(sess.library_h2d_bytes()),
                    "d2h_bytes_per_step": int(st_e2e[0]["d2h_bytes"]), "ms_per_step": 1e3 * t_e2e_max / args.steps,
                    "mbp_per_s": Q_all * args.steps / t_e2e_max / 1e6,
                    "what": "Session.set_library + set_queries (host buffers -> HBM) + search (records back on the host"
                            + (", gathered on rank 0" if world > 1 else "") + ") per step"},
            "gpu_launches": int(sum(s["launches"] for s in st_res)),
            "roofline": {"bound": "int-simd (DPX VIADDMNMX/VIMNMX3 16x2); not hbm, not tensor", "kernel": "swg_score_kernel<16,false>",
                         "achieved": achieved, "peak": issue_bound, "unit": "GCUPS", "frac": achieved / issue_bound,
                         "peak_source": "SURVEY 8(d): VIADDMNMX.S16x2 issue rate measured live on this GPU (rmgpu_measure_dpx_peak) "
                                        "x 32 lanes x 2 cells / 5 instructions per packed cell",
                         "traffic": traffic_from_profile(cells_step), "k1_ms_per_launch": k1_ms, "dpx_warp_instr_per_s": dpx_ips,
                         "microbenchmark_gcups": peak_cups / 1e9, "frac_of_microbenchmark": achieved / (peak_cups / 1e9),
                         "hbm_algorithmic_gbs": (Q / 2 + sum(len(x) for x in sseqs) * 8) / (k1_ms * 1e-3) / 1e9,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "stage_ms": {k: sum(s[k] for s in st_res) / len(st_res) for k in ("ms_score", "ms_origin", "ms_traceback", "ms_host_post", "ms_total")},
            "hits": n_hits[0], "pairs_passed": s0["n_pairs_passed"], "candidates": s0["n_candidates"], "origin_cells": s0["origin_cells"],
            "origin_redo": s0["n_origin_redo"], "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline and world == 1:
            with tempfile.TemporaryDirectory() as d:
                c, dt, desc = oracle_sample(W, d, snames, sseqs, qn, qs, 1, 4)
            line["cpu_baseline"] = {"value": c / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port", "sample": desc,
                                    "host_cores": os.cpu_count()}
        if not args.no_driver and world == 1 and harness() and args.workload == "config2":
            # the product path: the reference driver with -engine gpu through a persistent rmgpu_server (whole run: Perl
            # batching, one search call per batch from forked children, ProcessRepeats, masking)
            from repeatmasker_b200 import synth

            subprocess.run(["make", "-C", os.path.join(ROOT, "repeatmasker_b200", "perl")], check=False, capture_output=True)
            genome = make_genome(W, genome_bp, sseqs)
            with tempfile.TemporaryDirectory() as d:
                driver_inputs(W, d, snames, sseqs, genome)
                sess.close()
                with engine.Server(os.path.join(d, "rmgpu.sock"), gpus=[local]):
                    driver_run("gpu", d, 8)  # warm: CUDA context, library upload
                    driver_inputs(W, d, snames, sseqs, genome)
                    dt = driver_run("gpu", d, 8)
                line["driver"] = {"what": "RepeatMasker -engine gpu -pa 8 -lib, whole run through rmgpu_server (second of two runs)",
                                  "wall_s": dt, "gcups_whole_run": cells_step / dt / 1e9}
        print(json.dumps(line))
    sess.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
