import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import util
from repeatmasker_b200 import engine, synth
m = util.write_matrix("20p43g.matrix", "/tmp/m.matrix")
snames, sseqs = synth.make_library(1, n_families=24, lo=80, hi=1500)
genome, _ = synth.make_genome(2, 30000, sseqs, coverage=0.45)
synth.write_fasta("/tmp/q.fa", ["chrT"], [genome]); synth.write_fasta("/tmp/s.fa", snames, sseqs)
try:
    got, st = engine.search_files("/tmp/q.fa", "/tmp/s.fa", "/tmp/m.matrix")
    want, _ = util.run_oracle("/tmp/q.fa", "/tmp/s.fa", "/tmp/m.matrix")
    print(len(got), len(want), st["n_origin_redo"], got == want or [ (g,w) for g,w in zip(got,want) if any(g[k]!=w[k] for k in ("score","qstart","qend","sstart","send"))][:2])
except Exception as e:
    print("ERR", e)
