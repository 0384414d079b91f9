#!/usr/bin/env python3
"""The smallest search that runs every kernel of the library (K0 streams/profile, K1, K2b + resolve, the int32 score pass
K2w through an int16 overflow -- and, with RMGPU_NO_WIDE=1, the K2 sweep fallback instead --, K3 warp and block teams, K4,
K5 incl. masklevel, row gather, the mask kernel of the stage chain, both collection filters), for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
The result is compared with the oracle, so a sanitizer-clean run is also a correct one."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from repeatmasker_b200 import engine, synth  # noqa: E402

snames, sseqs = synth.make_library(11, n_families=10, lo=80, hi=900)
long_seq = synth.random_dna(__import__("numpy").random.default_rng(5), 4300).tobytes().decode()
snames.append("longfam#LINE/L1")
sseqs.append(long_seq)
genome, _ = synth.make_genome(12, 5000, sseqs[:10], coverage=0.5)
genome = genome[:2500] + long_seq[100:4200] + genome[2500:]  # an identical 4.1 kbp copy: raw score > 32767 (int16 overflow path)
with tempfile.TemporaryDirectory() as d:
    q, s, m = os.path.join(d, "q.fa"), os.path.join(d, "s.fa"), os.path.join(d, "m.matrix")
    synth.write_fasta(q, ["rec1", "rec2"], [genome, genome[1000:4000]])
    synth.write_fasta(s, snames, sseqs)
    util.write_matrix("20p43g.matrix", m)
    for ml in (101, 80):
        got, st = engine.search_files(q, s, m, device=0, masklevel=ml)
        want, _ = util.run_oracle(q, s, m, masklevel=ml)
        assert len(got) == len(want) > 0, (len(got), len(want))
        for g, w in zip(got, want):
            for k in ("score", "div", "dele", "ins", "qstart", "qend", "strand", "sname", "sstart", "send", "qaln", "saln", "star"):
                assert g[k] == w[k], (k, g, w)
        print(f"masklevel {ml}: {len(got)} hits identical to the oracle, launches {st['launches']}, origin redo {st['n_origin_redo']}, max raw {max(h['raw_score'] for h in got)}")
    # stage chain: mask the first record's hits on the device, search again, compare with the oracle on the masked file
    sess = engine.Session()
    sess.set_matrix(m)
    sess.set_library(sseqs, snames)
    sess.set_queries([genome, genome[1000:4000]], ["rec1", "rec2"])
    hits, _ = sess.search(engine.make_params())
    ranges = [(h["query_index"], h["qstart"], h["qend"]) for h in hits if h["query_index"] == 0 and h["score"] < 5000]
    sess.mask_queries(ranges)
    got, _ = sess.search(engine.make_params())
    b = bytearray(genome.encode())
    for _, lo, hi in ranges:
        b[lo - 1:hi] = b"X" * (hi - lo + 1)
    synth.write_fasta(q, ["rec1", "rec2"], [b.decode(), genome[1000:4000]])
    want, _ = util.run_oracle(q, s, m)
    assert [(g["score"], g["qstart"], g["qend"], g["qaln"]) for g in got] == [(w["score"], w["qstart"], w["qend"], w["qaln"]) for w in want]
    print(f"stage chain: {len(ranges)} ranges masked on the device, {len(got)} hits identical to the oracle on the masked batch")
    sess.close()
names = ["a"] * 6 + ["b"] * 4
starts = [1, 50, 60, 300, 310, 320, 1, 5, 400, 410]
ends = [200, 150, 400, 500, 480, 900, 100, 90, 800, 700]
scores = [900, 300, 500, 700, 250, 800, 400, 300, 600, 200]
print("masklevel filter keeps", engine.collection_filter(names, starts, ends, scores, 80))
print("contained filter keeps", engine.collection_filter(names, starts, ends, scores, 80, contained=True))
