import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np, util
from repeatmasker_b200 import engine, synth
m = util.write_matrix("20p41g.matrix", "/tmp/m.matrix")
rng = np.random.default_rng(3)
genome = synth.random_dna(rng, 1_000_000).tobytes().decode()
qn, qs = synth.fragment(genome, "g", 60000, 2000)
lens=[4096]*144
seqs = [synth.random_dna(rng, int(n)).tobytes().decode() for n in lens]
s = engine.Session(); s.set_matrix(m); s.set_library(seqs); s.set_queries(qs)
for i in range(3):
    try:
        best, ms = s.score_only(engine.make_params()); print("ok", ms, flush=True)
    except Exception as e:
        print("ERR", e, flush=True); break
