"""Randomised parity: librmgpu (through the C ABI, files in / records out) against the CPU oracle on random small
searches -- random libraries (short families, an occasional 6-9 kbp one for stripe chains and int16 overflow, simple
repeats), random genomes (coverage, divergence, IUPAC codes, N runs, several records of random size), random matrices
of the 46 the reference ships, gap recipes, minscore, masklevel, raw / adjusted.  Every hit must be identical
(score, percentages, coordinates, strand, '*', alignment rows).  usage: python tools/fuzz_parity.py [--cases 40] [--seed 1]"""
import argparse, os, sys, tempfile, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from repeatmasker_b200 import engine, synth
from test_gpu_parity import assert_same_hits

ap = argparse.ArgumentParser(); ap.add_argument("--cases", type=int, default=40); ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()
names = sorted(util.matrices())
recipes = [(-35, -7, -6), (-30, -6, -5), (-27, -6, -5), (-25, -5, -5), (-40, -8, -7), (-30, -5, -6)]
tot_cells, t0, fails = 0.0, time.time(), 0
for case in range(a.cases):
    seed = a.seed * 100003 + case
    rng = np.random.default_rng(seed)
    nfam = int(rng.integers(2, 30))
    snames, sseqs = synth.make_library(seed, n_families=nfam, lo=30, hi=int(rng.choice([400, 1500, 3000])))
    if rng.random() < 0.25:
        snames.append("long#LINE/L1"); sseqs.append(synth.random_dna(rng, int(rng.integers(6000, 9000))).tobytes().decode())
    if rng.random() < 0.3:
        snames.append("(CA)n#Simple_repeat"); sseqs.append("CA" * int(rng.integers(20, 90)))
    glen = int(rng.integers(1500, 16000))
    genome, _ = synth.make_genome(seed + 1, glen, sseqs, coverage=float(rng.uniform(0.2, 0.7)),
                                  div_lo=float(rng.uniform(0.0, 0.1)), div_hi=float(rng.uniform(0.12, 0.35)))
    g = list(genome)
    if rng.random() < 0.3:  # near-identical copy of the long subject -> scores past the int16 range
        src = sseqs[-1] if len(sseqs[-1]) > 5000 else max(sseqs, key=len)
        cp = synth.mutate(rng, src, 0.01, 0.001)
        at = int(rng.integers(0, max(1, len(g) - 10)))
        g[at:at] = list(cp if rng.random() < 0.5 else synth.revcomp(cp))
    if rng.random() < 0.4:
        for _ in range(int(rng.integers(1, 30))):
            g[int(rng.integers(0, len(g)))] = "NRYKMSWBDHVX"[int(rng.integers(12))]
    if rng.random() < 0.3:
        at = int(rng.integers(0, len(g) - 50)); g[at:at + int(rng.integers(5, 40))] = "N" * 10
    if rng.random() < 0.3:
        at = int(rng.integers(0, len(g) - 200)); g[at:at] = list("CA" * int(rng.integers(15, 80)))
    genome = "".join(g)
    frag = int(rng.choice([700, 3000, 60000]))
    qn, qs = synth.fragment(genome, f"c{case}", frag=frag, overlap=min(200, frag // 4))
    gi, ie, de = recipes[int(rng.integers(len(recipes)))]
    kw = dict(gap_init=gi, ins_gap_ext=ie, del_gap_ext=de, minscore=int(rng.choice([150, 180, 225, 300, 400])),
              masklevel=int(rng.choice([101, 101, 90, 80, 50, 0])), raw=bool(rng.random() < 0.3))
    mname = names[int(rng.integers(len(names)))]
    d = tempfile.mkdtemp()
    q, s, m = os.path.join(d, "q.fa"), os.path.join(d, "s.fa"), os.path.join(d, mname)
    synth.write_fasta(q, qn, qs); synth.write_fasta(s, snames, sseqs); util.write_matrix(mname, m)
    want, _ = util.run_oracle(q, s, m, **kw)
    got, st = engine.search_files(q, s, m, **kw)
    tot_cells += st["cells"]
    try:
        assert_same_hits(got, want)
        print(f"case {case} seed {seed}: ok  {len(want):4d} hits, {len(qs)} records x {len(sseqs)} subjects, {mname} {kw}", flush=True)
    except AssertionError as e:
        fails += 1
        print(f"case {case} seed {seed}: MISMATCH {str(e)[:600]}\n   {mname} {kw}", flush=True)
print(f"{a.cases - fails}/{a.cases} cases identical, {tot_cells:.3g} cells, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
