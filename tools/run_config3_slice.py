"""BASELINE configuration 3 (46 Mbp genome x 1,500-family library, -s) on a slice: the full run is ~1.7e14 cells
(~30 s on one B200); this script searches `--mbp` of it to show the sustained rate with a 1.8 Mbp library and many
fragments per call (memory-bounded chunking of the batch), and with --oracle compares every fragment hit for hit
(lines, rows, '*', order) with the CPU oracle run on all host cores -- the CPU arm of this configuration.
usage: python tools/run_config3_slice.py [--mbp 4] [--oracle]
With --families 20000 --seed 1005 --mbp 0.24 the same script runs a slice of configuration 5 (24 Mbp Dfam-scale library)."""
import argparse, os, sys, tempfile, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import util
from repeatmasker_b200 import engine, sharding, synth
ap = argparse.ArgumentParser(); ap.add_argument("--mbp", type=float, default=4.0)
ap.add_argument("--families", type=int, default=1500); ap.add_argument("--seed", type=int, default=1002)
ap.add_argument("--oracle", action="store_true"); a = ap.parse_args()
m = util.write_matrix("20p41g.matrix", "/tmp/20p41g.matrix")
snames, sseqs = synth.make_library(a.seed, n_families=a.families)
parts, left, k = [], int(a.mbp * 1e6), 0
while left > 0:  # one sequence, generated 1 Mbp at a time
    g, _ = synth.make_genome(a.seed + 7919 * k, min(left, 1_000_000), sseqs, coverage=0.45); parts.append(g); left -= 1_000_000; k += 1
qn, qs = synth.fragment("".join(parts), "synth0", 60000, 2000)
s = engine.Session(); s.set_matrix(m); s.set_library(sseqs, snames); s.set_queries(qs, qn)
p = engine.make_params(minmatch=8)
s.search(p)
t0 = time.perf_counter(); hits, st = s.search(p); dt = time.perf_counter() - t0
print(f"library {sum(map(len, sseqs))} bp, {len(qs)} fragments, {sum(map(len, qs))} bp: {len(hits)} hits, {st['cells'] / dt / 1e9:.0f} GCUPS "
      f"({sum(map(len, qs)) / dt / 1e6:.2f} Mbp/s), K1 {st['ms_score']:.0f} ms K2 {st['ms_origin']:.0f} ms K3 {st['ms_traceback']:.0f} ms "
      f"post {st['ms_host_post']:.0f} ms, launches {st['launches']}, passed {st['n_pairs_passed']}, redo {st['n_origin_redo']}")
if a.oracle:
    per = sharding.split_hits_by_fragment(hits, qn)
    fields = ("score", "div", "dele", "ins", "qname", "qstart", "qend", "qleft", "strand", "sname", "sstart", "send", "sleft", "star", "qaln", "saln")
    t0 = time.perf_counter(); same = nh = 0
    with tempfile.TemporaryDirectory() as d:
        for f in range(len(qs)):
            q = os.path.join(d, "q.fa"); synth.write_fasta(q, [qn[f]], [qs[f]])
            want = util.run_oracle_parallel(q, snames, sseqs, m, d, minmatch=8)
            ok = len(want) == len(per[f]) and all(all(g[k] == w[k] for k in fields) for g, w in zip(per[f], want))
            same += ok; nh += len(want)
            if not ok: print(f"  fragment {f + 1}: DIFFERENT ({len(per[f])} vs {len(want)} hits)")
    dt2 = time.perf_counter() - t0
    print(f"oracle (restated crossmatch, {os.cpu_count()} host cores): {len(qs)} fragments in {dt2:.0f} s = {st['cells'] / dt2 / 1e9:.1f} GCUPS; "
          f"{same} of {len(qs)} fragments identical hit for hit ({nh} hits: every line, both rows, '*', order)")
    sys.exit(0 if same == len(qs) else 1)
