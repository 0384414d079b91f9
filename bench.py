#!/usr/bin/env python3
"""bench.py -- the hot path on BASELINE.json's configuration 2:
   1 Mbp synthetic sequence with implanted 10-30 %-diverged copies of a 500-family consensus library, both strands.

A "step" = one pass of the search (K0 streams/profile, K1 score, K2 origin, K3 traceback, host post) over the whole
batch of SimpleBatcher fragments (60 kbp, 2 kbp overlap) against the whole library: Q*L*2 ~ 1e12 DP cells.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--genome-bp B] [--families F]

N > 1 is launched by torchrun (one rank per GPU, NCCL only for the barrier / max-reduce of the timing): every rank
searches its own 1 Mbp genome against the replicated library ("weak" scaling, no data-path collective).

--impl reference times the CPU arm: the restated crossmatch oracle (oracle/_build/cross_match; Phil Green's
cross_match is not part of the reference tree and cannot be built here), all host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "GCUPS (1e9 Smith-Waterman-Gotoh cell updates/s, Q*L*2 per stage) searched vs restated crossmatch on host cores"
GAP = dict(gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5, minscore=225, masklevel=101, minmatch=9, bandwidth=14)


def workload(args, rank, world=1):
    """Library (replicated) and this rank's fragments.  One seeded genome of world x genome_bp bases is cut into
    SimpleBatcher fragments and fragment b goes to rank b mod world (repeatmasker_b200.sharding): per-GPU work is
    fixed as N grows (weak scaling) and no rank ever needs another rank's data."""
    from repeatmasker_b200 import sharding, synth

    snames, sseqs = synth.make_library(1001, n_families=args.families)
    qn, qs = [], []
    for part in range(world):  # the genome is generated 1 Mbp at a time so that every N sees the same first Mbp
        genome, truth = synth.make_genome(1001 + 7919 * part, args.genome_bp, sseqs, coverage=0.45)
        n, q = synth.fragment(genome, f"synth{part}", frag=60000, overlap=2000)
        qn += n
        qs += q
    mine = sharding.my_fragments(len(qs), rank, world)
    return snames, sseqs, [qn[i] for i in mine], [qs[i] for i in mine]


def matrix_file(tmpdir):
    import util

    return util.write_matrix("20p41g.matrix", os.path.join(tmpdir, "20p41g.matrix"))


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


def cpu_sample(args, tmpdir, snames, sseqs, qn, qs, threads):
    """Bounded sample of the same workload for the CPU arm: the first fragment against `nfam` families taken at a
    constant stride through the library (family popularity falls along the library, so the first families alone
    would over-represent hit-dense pairs), dealt round-robin to `threads` oracle processes.
    Returns (cells, seconds, description)."""
    import util
    from repeatmasker_b200 import synth

    nfam = min(len(sseqs), args.cpu_families * threads)
    qpath = os.path.join(tmpdir, "cpu_q.fa")
    qseq = qs[0][: args.cpu_query_bp]
    synth.write_fasta(qpath, [qn[0]], [qseq])
    m = matrix_file(tmpdir)
    procs, cells = [], 0.0
    picked = [(i * len(sseqs)) // nfam for i in range(nfam)]
    for t in range(threads):
        sub = picked[t::threads]
        if not sub:
            continue
        sp = os.path.join(tmpdir, f"cpu_s{t}.fa")
        synth.write_fasta(sp, [snames[i] for i in sub], [sseqs[i] for i in sub])
        cells += 2.0 * len(qseq) * sum(len(sseqs[i]) for i in sub)
        procs.append(util.oracle_argv(qpath, sp, m, **GAP))
    t0 = time.perf_counter()
    running = [subprocess.Popen(a, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for a in procs]
    for p in running:
        p.wait()
    dt = time.perf_counter() - t0
    desc = (f"fragment 1 ({len(qseq)} bp) x {nfam} of {len(sseqs)} families (every {len(sseqs) // max(nfam, 1)}th), both strands, "
            f"{len(procs)} oracle processes (restated crossmatch, this repo)")
    return cells, dt, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "cross_match")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    threads = os.cpu_count() or 1
    snames, sseqs, qn, qs = workload(args, 0)
    with tempfile.TemporaryDirectory() as d:
        for _ in range(args.warmup if args.warmup < 2 else 1):
            cpu_sample(args, d, snames, sseqs, qn, qs, threads)
        tot_cells, tot_s = 0.0, 0.0
        for _ in range(args.steps):
            c, dt, desc = cpu_sample(args, d, snames, sseqs, qn, qs, threads)
            tot_cells += c
            tot_s += dt
    gcups = tot_cells / tot_s / 1e9
    line = {"metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "impl": "reference",
            "config": config_dict(args, snames, sseqs, qn, qs, 1),
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
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


def config_dict(args, snames, sseqs, qn, qs, world):
    Q, L = sum(len(x) for x in qs), sum(len(x) for x in sseqs)
    return {"workload": "BASELINE configs[1]: 1 Mbp synthetic genome (45% implanted 10-30%-diverged copies) x 500-family "
                        "synthetic consensus library, both strands, SimpleBatcher 60 kbp fragments / 2 kbp overlap",
            "genome_bp": args.genome_bp, "families": len(sseqs), "library_bp": L, "fragments": len(qs),
            "query_bp_searched": Q, "cells_per_step_per_gpu": 2.0 * Q * L, "matrix": "20p41g", "gap": [-30, -6, -5],
            "minscore": 225, "seed": 1001, "parallelism": f"fragment-sharded x{world}, library replicated, no collective",
            "l2": "L2 flushed between timed steps (256 MiB write)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-bp", type=int, default=1_000_000)
    ap.add_argument("--families", type=int, default=500)
    ap.add_argument("--cpu-families", type=int, default=8, help="families per oracle process in the --impl reference sample")
    ap.add_argument("--cpu-baseline-families", type=int, default=24, help="families in the single-thread cpu_baseline sample")
    ap.add_argument("--cpu-query-bp", type=int, default=60000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from repeatmasker_b200 import build as rm_build
    from repeatmasker_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        rm_build.build()
    if world > 1:
        dist.barrier()

    snames, sseqs, qn, qs = workload(args, rank, world)
    tmp = tempfile.TemporaryDirectory()
    m = matrix_file(tmp.name)
    params = engine.make_params(device=local, **GAP)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    sess = engine.Session(local)
    sess.set_matrix(m)
    sess.set_library(sseqs, snames)
    sess.set_queries(qs, qn)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

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
        return sess.search(params)[1]

    def step_e2e():  # host buffers in, host records out: library + batch upload, search, result download
        sess.set_library(sseqs, snames)
        sess.set_queries(qs, qn)
        return sess.search(params)[1]

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
    cc = torch.tensor([cells_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
    t_res_max, t_e2e_max = tt.tolist()
    cells_all = cc.item()
    value = cells_all * args.steps / t_res_max / 1e9
    e2e = cells_all * args.steps / t_e2e_max / 1e9

    if rank == 0:
        k1_ms = sum(s["ms_score"] for s in st_res) / len(st_res)
        peak_cups, dpx_ips = engine.measure_dpx_peak(local)
        achieved = cells_step / (k1_ms * 1e-3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        Q = sum(len(x) for x in qs)
        nS = st_res[0]
        line = {
            "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_res_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16x2 (DPX) score pass, int32 origin/traceback, f64 complexity adjustment", "data": "synthetic",
            "config": config_dict(args, snames, sseqs, qn, qs, world),
            "mbp_per_s": Q * world * args.steps / t_res_max / 1e6,
            "e2e": {"value": e2e, "unit": "GCUPS", # query records: 4-bit packed (counted by the library); library: 1 B/row codes in 16-row groups + 1 B/base codes
                    # for K2/K3 + 5 B of flags per group = 2.33 B/base (the 8 B/row profiles are built on the device)
                    "h2d_bytes_per_step": int(st_e2e[0]["h2d_bytes"]) + int(2.33 * sum(len(x) for x in sseqs)),
                    "d2h_bytes_per_step": int(st_e2e[0]["d2h_bytes"]), "ms_per_step": 1e3 * t_e2e_max / args.steps,
                    "mbp_per_s": Q * world * args.steps / t_e2e_max / 1e6,
                    "what": "Session.set_library + set_queries (host buffers -> HBM) + search (records back on the host) per step"},
            "gpu_launches": int(sum(s["launches"] for s in st_res)),
            "roofline": {"bound": "int-simd (DPX VIADDMNMX/VIMNMX3 16x2); not hbm, not tensor", "kernel": "swg_score_kernel<16,false>",
                         "achieved": achieved, "peak": peak_cups / 1e9, "unit": "GCUPS", "frac": achieved / (peak_cups / 1e9),
                         "peak_source": "rmgpu_measure_dpx_peak: register-resident loop of the same recurrence, measured live on this GPU",
                         "traffic": traffic_from_profile(cells_step), "k1_ms_per_launch": k1_ms, "dpx_warp_instr_per_s": dpx_ips,
                         # SURVEY 8(d)'s alternative denominator: raw DPX issue rate x 32 lanes x 2 cells / 5 instructions per cell pair
                         "frac_of_dpx_issue_bound": achieved * 1e9 / (dpx_ips * 32 * 2 / 5) if dpx_ips else None,
                         "hbm_algorithmic_gbs": (Q / 2 + sum(len(x) for x in sseqs) * 8) / (k1_ms * 1e-3) / 1e9,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "stage_ms": {k: sum(s[k] for s in st_res) / len(st_res) for k in ("ms_score", "ms_origin", "ms_traceback", "ms_host_post", "ms_total")},
            "hits": None, "pairs_passed": nS["n_pairs_passed"], "candidates": nS["n_candidates"], "origin_cells": nS["origin_cells"], "origin_redo": nS["n_origin_redo"],
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            with tempfile.TemporaryDirectory() as d:
                args.cpu_families = args.cpu_baseline_families
                c, dt, desc = cpu_sample(args, d, snames, sseqs, qn, qs, 1)
            line["cpu_baseline"] = {"value": c / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port", "sample": desc,
                                    "host_cores": os.cpu_count()}
        print(json.dumps(line))
    sess.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
