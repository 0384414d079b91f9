import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np, util
from repeatmasker_b200 import engine, synth
m = util.write_matrix("20p41g.matrix", "/tmp/m.matrix")
rng = np.random.default_rng(3)
genome = synth.random_dna(rng, 1_000_000).tobytes().decode()
qn, qs = synth.fragment(genome, "g", 60000, 2000)
def run(label, lens):
    seqs = [synth.random_dna(rng, int(n)).tobytes().decode() for n in lens]
    s = engine.Session(); s.set_matrix(m); s.set_library(seqs); s.set_queries(qs)
    p = engine.make_params()
    s.score_only(p)
    best, ms = s.score_only(p)
    cells = 2.0 * sum(len(x) for x in qs) * sum(lens)
    print(f"{label:40s} subjects {len(lens):6d} bases {sum(lens):8d}  K1 {ms:8.2f} ms  {cells/ms/1e6:8.1f} GCUPS", flush=True)
    s.close()
run("no chains: 2300 x 250", [250]*2300)
run("no chains: 2300 x 256", [256]*2300)
run("chains of 2: 1150 x 512", [512]*1150)
run("chains of 4: 575 x 1024", [1024]*575)
run("chains of 16: 144 x 4096", [4096]*144)
run("chains of 32: 72 x 8192", [8192]*72)
sn, ss = synth.make_library(1001, 500)
run("config-2 library (log-normal)", [len(x) for x in ss])
