import ctypes as C, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np, util
from repeatmasker_b200 import engine, synth
m = util.write_matrix("20p43g.matrix", "/tmp/m.matrix"); om = util.parse_matrix(m)
op = util.OraParams(gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5); lib = util.oracle_lib()
def check(qseqs, sseqs, label):
    s = engine.Session(); s.set_matrix(m); s.set_library(sseqs); s.set_queries(qseqs)
    best, ms = s.score_only(engine.make_params()); bad = []
    for qi, q in enumerate(qseqs):
        for strand, ql in enumerate((q, synth.revcomp(q))):
            qc = util.encode(om, ql)
            for si, sq in enumerate(sseqs):
                want = lib.ora_sw_score(qc, len(qc), util.encode(om, sq), len(sq), C.byref(om), C.byref(op))
                got = best[(qi * len(sseqs) + si) * 2 + strand]
                if got != want: bad.append((qi, si, strand, len(q), len(sq), got, want))
    print(label, "pairs", len(qseqs)*len(sseqs)*2, "bad", len(bad), bad[:8]); s.close()
rng = np.random.default_rng(1)
r = lambda n: synth.random_dna(rng, n).tobytes().decode()
g = r(3000)
check([g], [r(100)], "1x100")
check([g], [g[500:700]], "copy200")
check([g], [r(256)], "1x256")
check([g], [r(257)], "1x257")
check([g], [g[100:900]], "copy800 (4 stripes)")
check([g], [r(100), r(50), r(7)], "3 small")
check([g, r(40), r(4)], [g[100:900], r(30)], "multi q")
check([r(31)], [r(100)], "q31"); check([r(32)], [r(100)], "q32"); check([r(33)], [r(100)], "q33")
snames, sseqs = synth.make_library(11, n_families=40, lo=20, hi=3000)
check([g], sseqs, "lib40")
for i in range(0, 40, 8): check([g], sseqs[i:i+8], f"lib[{i}:{i+8}] lens {[len(x) for x in sseqs[i:i+8]]}")
