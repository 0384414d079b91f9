"""Run, on ONE GPU, the exact per-rank workloads that `bench.py --gpus N` gives ranks 0..N-1 (N = 4 and 8), one search
each: a dry run of the multi-GPU bench inputs (different genomes -> different hit patterns, table sizes, overflow
triples) without needing N GPUs.  usage: python tools/check_ranks.py [--worlds 4,8]"""
import argparse, os, sys, time, tempfile
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from repeatmasker_b200 import engine, sharding, synth

ap = argparse.ArgumentParser(); ap.add_argument("--worlds", default="4,8"); a = ap.parse_args()
args = argparse.Namespace(families=500, genome_bp=1_000_000)
tmp = tempfile.mkdtemp(); m = bench.matrix_file(tmp)
snames, sseqs = synth.make_library(1001, n_families=args.families)
s = engine.Session(); s.set_matrix(m); s.set_library(sseqs, snames)
p = engine.make_params(**bench.GAP)
worlds = [int(x) for x in a.worlds.split(",")]
parts = {}
for part in range(max(worlds)):
    g, _ = synth.make_genome(1001 + 7919 * part, args.genome_bp, sseqs, coverage=0.45)
    parts[part] = synth.fragment(g, f"synth{part}", frag=60000, overlap=2000)
for world in worlds:
    qn, qs = [], []
    for part in range(world):
        qn += parts[part][0]; qs += parts[part][1]
    for rank in range(world):
        mine = sharding.my_fragments(len(qs), rank, world)
        s.set_queries([qs[i] for i in mine], [qn[i] for i in mine])
        t0 = time.perf_counter(); hits, st = s.search(p); dt = time.perf_counter() - t0
        print(f"world {world} rank {rank}: {len(mine)} fragments, {len(hits)} hits, {st['cells'] / dt / 1e9:.0f} GCUPS, K1 {st['ms_score']:.1f} "
              f"K2 {st['ms_origin']:.1f} K3 {st['ms_traceback']:.1f} ms, passed {st['n_pairs_passed']}, redo {st['n_origin_redo']}", flush=True)
print("all ranks ok")
