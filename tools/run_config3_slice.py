"""BASELINE configuration 3 (46 Mbp genome x 1,500-family library, -s) on a slice: the full run is ~1.7e14 cells
(~50 s on one B200); this script searches `--mbp` of it to show the sustained rate with a 1.8 Mbp library and many
fragments per call (memory-bounded chunking of the batch).  usage: python tools/run_config3_slice.py [--mbp 4]
With --families 20000 --seed 1005 --mbp 0.24 the same script runs a slice of configuration 5 (24 Mbp Dfam-scale library)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import util
from repeatmasker_b200 import engine, synth
ap = argparse.ArgumentParser(); ap.add_argument("--mbp", type=float, default=4.0)
ap.add_argument("--families", type=int, default=1500); ap.add_argument("--seed", type=int, default=1002); a = ap.parse_args()
m = util.write_matrix("20p41g.matrix", "/tmp/20p41g.matrix")
snames, sseqs = synth.make_library(a.seed, n_families=a.families)
qn, qs = [], []
left, part = int(a.mbp * 1e6), 0
while left > 0:
    g, _ = synth.make_genome(a.seed + 31 * part, min(left, 1_000_000), sseqs, coverage=0.45); left -= 1_000_000; part += 1
    n, q = synth.fragment(g, f"chrpart{part}", 60000, 2000); qn += n; qs += q
s = engine.Session(); s.set_matrix(m); s.set_library(sseqs, snames); s.set_queries(qs, qn)
p = engine.make_params(minmatch=8)
s.search(p)
t0 = time.perf_counter(); hits, st = s.search(p); dt = time.perf_counter() - t0
print(f"library {sum(map(len, sseqs))} bp, {len(qs)} fragments, {sum(map(len, qs))} bp: {len(hits)} hits, {st['cells'] / dt / 1e9:.0f} GCUPS "
      f"({sum(map(len, qs)) / dt / 1e6:.2f} Mbp/s), K1 {st['ms_score']:.0f} ms K2 {st['ms_origin']:.0f} ms K3 {st['ms_traceback']:.0f} ms "
      f"post {st['ms_host_post']:.0f} ms, launches {st['launches']}, passed {st['n_pairs_passed']}, redo {st['n_origin_redo']}")
