"""BASELINE configuration 2 at FULL size (1 Mbp x 500 families, ~1.2e12 cells): the oracle would need hours, so the
result is checked through size-independent properties, with the oracle's scoring functions as the checker:
  * every reported alignment, re-scored column by column with the matrix and gap penalties, gives exactly the raw
    score the DP reported, and its complexity-adjusted score is the reported score;
  * coordinates, alignment rows and percentage strings are mutually consistent;
  * output order is the specified order; '*' flags are consistent; no two hits share (query, strand, subject, origin);
  * K1's best score per (query, subject, strand) bounds every hit of that triple and reaches it for the best one;
  * a second run returns the identical result (no dependence on the dynamic ticket schedule);
  * searching the fragments in two halves (two "ranks") returns the same hits as one batch.
Also a many-family library (BASELINE configuration 5 in miniature) against the oracle."""
import ctypes as C
import os

import pytest

import util
from repeatmasker_b200 import engine, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config2(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg2")
    snames, sseqs = synth.make_library(1001, n_families=500)
    genome, truth = synth.make_genome(1001, 1_000_000, sseqs, coverage=0.45)
    qn, qs = synth.fragment(genome, "synth0", frag=60000, overlap=2000)
    m = util.write_matrix("20p41g.matrix", str(d / "20p41g.matrix"))
    s = engine.Session()
    s.set_matrix(m)
    s.set_library(sseqs, snames)
    s.set_queries(qs, qn)
    p = engine.make_params()
    hits, st = s.search(p)
    return dict(s=s, p=p, hits=hits, st=st, snames=snames, sseqs=sseqs, qn=qn, qs=qs, m=m, truth=truth)


def test_full_size_alignments_rescore_exactly(config2):
    c = config2
    om = util.parse_matrix(c["m"])
    op = util.OraParams(gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5)
    lib = util.oracle_lib()
    assert len(c["hits"]) > 800
    assert c["st"]["cells"] == 2.0 * sum(map(len, c["qs"])) * sum(map(len, c["sseqs"]))
    qlen = dict(zip(c["qn"], map(len, c["qs"])))
    slen = dict(zip(c["snames"], map(len, c["sseqs"])))
    qseq = dict(zip(c["qn"], c["qs"]))
    sseq = dict(zip(c["snames"], c["sseqs"]))
    for h in c["hits"]:
        raw, adj = C.c_int(), C.c_int()
        lib.ora_rescore(h["qaln"].encode(), h["saln"].encode(), C.byref(om), C.byref(op), C.byref(raw), C.byref(adj))
        assert raw.value == h["raw_score"], h
        assert adj.value == h["score"] >= 225, h
        qa, sa = h["qaln"].replace("-", ""), h["saln"].replace("-", "")
        assert len(qa) == h["qend"] - h["qstart"] + 1 and len(sa) == h["send"] - h["sstart"] + 1
        assert h["qleft"] == qlen[h["qname"]] - h["qend"] and h["sleft"] == slen[h["sname"]] - h["send"]
        assert sa == sseq[h["sname"]][h["sstart"] - 1:h["send"]]
        frag = qseq[h["qname"]][h["qstart"] - 1:h["qend"]]
        assert qa == (synth.revcomp(frag) if h["strand"] == "C" else frag)
        assert h["qaln"][0] != "-" and h["saln"][0] != "-" and h["qaln"][-1] != "-" and h["saln"][-1] != "-"
        span = len(qa)
        nsub = sum(1 for a, b in zip(h["qaln"], h["saln"]) if a != b and a != "-" and b != "-")
        assert h["div"] == "%.2f" % (100.0 * nsub / span)
        assert h["dele"] == "%.2f" % (100.0 * h["qaln"].count("-") / span)
        assert h["ins"] == "%.2f" % (100.0 * h["saln"].count("-") / span)


def test_full_size_order_flags_and_uniqueness(config2):
    hits = config2["hits"]
    qi = {n: i for i, n in enumerate(config2["qn"])}
    si = {n: i for i, n in enumerate(config2["snames"])}
    key = lambda h: (qi[h["qname"]], h["qstart"], -si[h["sname"]], h["sstart"], h["qend"], h["strand"] == "C", h["send"])
    assert [key(h) for h in hits] == sorted(key(h) for h in hits)
    seen = set()
    for h in hits:
        k = (h["qname"], h["strand"], h["sname"], h["qstart"] if h["strand"] == "+" else h["qend"], h["sstart"])
        assert k not in seen, k  # one hit per alignment origin
        seen.add(k)
    byq = {}
    for h in hits:
        byq.setdefault(h["qname"], []).append(h)
    for hs in byq.values():
        for h in hs:
            star = any(o["score"] > h["score"] and o["qstart"] <= h["qend"] and o["qend"] >= h["qstart"] for o in hs)
            assert star == h["star"]


def test_full_size_score_pass_bounds_hits_and_is_deterministic(config2):
    c = config2
    best, ms = c["s"].score_only(c["p"])
    ns = len(c["sseqs"])
    qi = {n: i for i, n in enumerate(c["qn"])}
    si = {n: i for i, n in enumerate(c["snames"])}
    top = {}
    for h in c["hits"]:
        k = (qi[h["qname"]], si[h["sname"]], 1 if h["strand"] == "C" else 0)
        top[k] = max(top.get(k, 0), h["raw_score"])
    for (q, s, strand), raw in top.items():
        b = best[(q * ns + s) * 2 + strand]
        assert b == min(raw, 32767) or (b == 32767 and raw >= 32000), (q, s, strand, b, raw)
    again, _ = c["s"].search(c["p"])
    assert again == c["hits"]
    cells = 2.0 * sum(map(len, c["qs"])) * sum(map(len, c["sseqs"]))
    assert cells / (ms * 1e-3) / 1e9 > 1000, "K1 slower than 1,000 GCUPS on a B200: not running the DPX path?"


def test_full_size_sharded_equals_single_batch(config2):
    """Fragments are independent units: two 'ranks' each searching their round-robin share give the same hits."""
    from repeatmasker_b200 import sharding

    c = config2
    parts = []
    ids = []
    for rank in range(2):
        mine = sharding.my_fragments(len(c["qs"]), rank, 2)
        c["s"].set_queries([c["qs"][i] for i in mine], [c["qn"][i] for i in mine])
        hits, _ = c["s"].search(c["p"])
        parts.append(sharding.split_hits_by_fragment(hits, [c["qn"][i] for i in mine]))
        ids.append(mine)
    merged = [None] * len(c["qs"])
    for mine, part in zip(ids, parts):
        for f, h in zip(mine, part):
            merged[f] = h
    strip = lambda h: {k: v for k, v in h.items() if k != "query_index"}
    want = sharding.split_hits_by_fragment(c["hits"], c["qn"])
    assert [[strip(h) for h in f] for f in merged] == [[strip(h) for h in f] for f in want]
    c["s"].set_queries(c["qs"], c["qn"])


@pytest.mark.parametrize("frag", [0, 8, 17])
def test_full_size_fragment_equals_oracle_hit_for_hit(config2, tmp_path, frag):
    """Whole fragments of configuration 2 (57 kbp x all 500 families x both strands, 6.7e10 cells each) against the oracle
    run on every host core: every hit line, alignment row, '*' flag and the output order.  A hit the score pass or the
    hot-block mask dropped would show up here as a missing line."""
    from repeatmasker_b200 import sharding
    from test_gpu_parity import assert_same_hits

    c = config2
    q = str(tmp_path / "q.fa")
    synth.write_fasta(q, [c["qn"][frag]], [c["qs"][frag]])
    want = util.run_oracle_parallel(q, c["snames"], c["sseqs"], c["m"], str(tmp_path), gap_init=-30, ins_gap_ext=-6,
                                    del_gap_ext=-5, minscore=225)
    got = sharding.split_hits_by_fragment(c["hits"], c["qn"])[frag]
    assert len(want) > 20
    assert_same_hits(got, want)


def test_full_size_recall_of_implanted_copies(config2):
    """The generator's own record of what it planted (`truth`: position, length, family, strand, divergence) against the
    hits: every intact copy (not overwritten by a later one) of >= 200 bases and <= 25 % divergence must be found with
    its family and strand over at least half of its length; >= 90 % of all intact copies of >= 150 bases."""
    c = config2
    truth = c["truth"]
    offs = synth.fragment_offsets(1_000_000)
    assert len(offs) == len(c["qn"])
    off = dict(zip(c["qn"], offs))
    by = {}
    for h in c["hits"]:
        by.setdefault((h["sname"], h["strand"]), []).append((off[h["qname"]] + h["qstart"], off[h["qname"]] + h["qend"]))
    ends = sorted((p, p + l) for p, l, *_ in truth)
    strong = strong_found = weak = weak_found = 0
    for i, (pos, ln, fam, strand, div) in enumerate(truth):
        if ln < 150 or any(tp < pos + ln and tp + tl > pos for tp, tl, *_ in truth[i + 1:]):
            continue
        ivs = by.get((c["snames"][fam], "C" if strand else "+"), [])
        ok = any(min(b, pos + ln) - max(a, pos + 1) + 1 >= 0.5 * ln for a, b in ivs)
        weak += 1
        weak_found += ok
        if ln >= 200 and div <= 0.25:
            strong += 1
            strong_found += ok
    assert strong > 200 and strong_found == strong, (strong_found, strong)
    assert weak_found >= 0.9 * weak, (weak_found, weak)


def test_config3_slice_equals_oracle_hit_for_hit(tmp_path):
    """BASELINE configuration 3 (1,500-family, 1.76 Mbp library, -s parameters) on the first 0.5 Mbp of its genome: every one of
    the 9 fragments against the CPU arm (the oracle on all host cores), hit for hit.  tools/run_config3_slice.py --oracle does
    the same on 4 Mbp (profiles/r2_config3_slice.txt)."""
    from repeatmasker_b200 import sharding
    from test_gpu_parity import assert_same_hits

    snames, sseqs = synth.make_library(1002, n_families=1500)
    genome, _ = synth.make_genome(1002, 500_000, sseqs, coverage=0.45)
    qn, qs = synth.fragment(genome, "synth0", frag=60000, overlap=2000)
    m = util.write_matrix("20p41g.matrix", str(tmp_path / "20p41g.matrix"))
    s = engine.Session()
    s.set_matrix(m)
    s.set_library(sseqs, snames)
    s.set_queries(qs, qn)
    hits, st = s.search(engine.make_params(minmatch=8))
    per = sharding.split_hits_by_fragment(hits, qn)
    assert len(qs) == 9 and len(hits) > 200
    for f in range(len(qs)):
        q = str(tmp_path / "q.fa")
        synth.write_fasta(q, [qn[f]], [qs[f]])
        want = util.run_oracle_parallel(q, snames, sseqs, m, str(tmp_path), minmatch=8)
        assert_same_hits(per[f], want)


def test_many_family_library_against_oracle(tmp_path):
    """Configuration 5 in miniature: thousands of short families plus a few long ones, many small query records
    (several ticket-order groups, many fillers per stripe)."""
    snames, sseqs = synth.make_library(77, n_families=3000, mu=4.6, sigma=0.5, lo=40, hi=400)
    ln, ls = synth.make_library(78, n_families=6, mu=8.2, sigma=0.3, lo=2000, hi=6000)
    snames = [f"L{n}" for n in ln] + snames
    sseqs = ls + sseqs
    genome, _ = synth.make_genome(79, 24000, sseqs, coverage=0.5)
    qn, qs = synth.fragment(genome, "chrM", frag=700, overlap=60)
    assert len(qs) >= 30
    q, s, m = str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "m.matrix")
    synth.write_fasta(q, qn, qs)
    synth.write_fasta(s, snames, sseqs)
    util.write_matrix("20p43g.matrix", m)
    want, _ = util.run_oracle(q, s, m, minscore=200)
    got, st = engine.search_files(q, s, m, minscore=200)
    assert len(want) > 20
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for k in ("score", "div", "dele", "ins", "qname", "qstart", "qend", "strand", "sname", "sstart", "send", "star",
                  "qaln", "saln"):
            assert g[k] == w[k], (k, g, w)


def _gc_block_genome(seed, sseqs, block_bp):
    """BASELINE configuration 4's query: blocks at 37/41/45/51 % GC with implanted copies, 15 % of the length in planted
    (CA)n / (A)n / (GAA)n tracts and AT-rich low-complexity runs."""
    import numpy as np

    rng = np.random.default_rng(seed)
    blocks = []
    for k, gc in enumerate((0.37, 0.41, 0.45, 0.51)):
        g, _ = synth.make_genome(seed + 17 * k, block_bp, sseqs, coverage=0.35, gc=gc)
        g = list(g)
        planted = 0
        while planted < 0.15 * block_bp:
            n = int(rng.integers(40, 260))
            kind = int(rng.integers(4))
            if kind == 0:
                tract = ("CA" * n)[:n]
            elif kind == 1:
                tract = "A" * n
            elif kind == 2:
                tract = ("GAA" * n)[:n]
            else:
                tract = "".join("AT"[int(x)] if rng.random() < 0.9 else "ACGT"[int(rng.integers(4))] for x in rng.integers(0, 2, n))
            tract = "".join(c if rng.random() > 0.03 else "ACGT"[int(rng.integers(4))] for c in tract)  # a little decay
            at = int(rng.integers(0, block_bp - n))
            g[at:at + n] = tract
            planted += n
        blocks.append("".join(g))
    return "".join(blocks)


@pytest.mark.parametrize("matrix,gi,ie,de,raw,minscore", [("14p41g.matrix", -35, -7, -6, False, 225),
                                                          ("25p41g.matrix", -27, -6, -5, True, 300),
                                                          ("18p41g.matrix", -30, -6, -5, False, 180)])
def test_config4_scale_low_complexity_properties(tmp_path, matrix, gi, ie, de, raw, minscore):
    """BASELINE configuration 4 at a size the oracle cannot do in a test (240 kbp x 0.6 Mbp library + simple-repeat
    consensi, ~2.9e11 cells, tens of thousands of hits from the planted tracts): every reported alignment re-scores
    exactly (raw and complexity-adjusted, checked with the oracle's scoring function under the same matrix and gap
    penalties), output order holds, one hit per origin, and a second run is identical."""
    snames, sseqs = synth.make_library(1001, n_families=500)
    snames = snames + ["(CA)n#Simple_repeat", "(A)n#Simple_repeat", "(GAA)n#Simple_repeat", "AT_rich#Low_complexity"]
    sseqs = sseqs + ["CA" * 90, "A" * 150, "GAA" * 60, "AT" * 50 + "A" * 20 + "TA" * 40]
    genome = _gc_block_genome(1003, sseqs[:500], 60000)
    qn, qs = synth.fragment(genome, "gcblocks", frag=60000, overlap=2000)
    m = util.write_matrix(matrix, str(tmp_path / matrix))
    s = engine.Session()
    s.set_matrix(m)
    s.set_library(sseqs, snames)
    s.set_queries(qs, qn)
    p = engine.make_params(gap_init=gi, ins_gap_ext=ie, del_gap_ext=de, minscore=minscore, raw=raw)
    hits, st = s.search(p)
    assert len(hits) > 2000 and sum(h["sname"].endswith("Simple_repeat") for h in hits) > 1000
    om = util.parse_matrix(m)
    op = util.OraParams(gap_init=gi, ins_gap_ext=ie, del_gap_ext=de)
    lib = util.oracle_lib()
    for h in hits:
        r, adj = C.c_int(), C.c_int()
        lib.ora_rescore(h["qaln"].encode(), h["saln"].encode(), C.byref(om), C.byref(op), C.byref(r), C.byref(adj))
        assert r.value == h["raw_score"], h
        assert h["score"] == (r.value if raw else adj.value) >= minscore, h
    qi = {n: i for i, n in enumerate(qn)}
    si = {n: i for i, n in enumerate(snames)}
    key = lambda h: (qi[h["qname"]], h["qstart"], -si[h["sname"]], h["sstart"], h["qend"], h["strand"] == "C", h["send"])
    assert [key(h) for h in hits] == sorted(key(h) for h in hits)
    seen = set()
    for h in hits:
        k = (h["qname"], h["strand"], h["sname"], h["qstart"] if h["strand"] == "+" else h["qend"], h["sstart"])
        assert k not in seen, k
        seen.add(k)
    again, _ = s.search(p)
    assert again == hits
    s.close()


def test_config5_scale_library_properties(tmp_path):
    """BASELINE configuration 5's library scale (20,000 families, 23.9 Mbp, 46,700 stripes; 24 GB of column checkpoints
    per query record, so two records are searched in two chunks) against two fragments: alignments re-score exactly,
    order holds, and the chunked result equals the per-record results."""
    snames, sseqs = synth.make_library(1005, n_families=20000)
    genome, _ = synth.make_genome(1005, 100000, sseqs, coverage=0.45)
    qn, qs = synth.fragment(genome, "chr", frag=60000, overlap=2000)
    assert len(qs) == 2
    m = util.write_matrix("20p41g.matrix", str(tmp_path / "20p41g.matrix"))
    s = engine.Session()
    s.set_matrix(m)
    s.set_library(sseqs, snames)
    s.set_queries(qs, qn)
    p = engine.make_params()
    hits, st = s.search(p)
    assert len(hits) > 50 and st["cells"] == 2.0 * sum(map(len, qs)) * sum(map(len, sseqs))
    om = util.parse_matrix(m)
    op = util.OraParams(gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5)
    lib = util.oracle_lib()
    for h in hits:
        r, adj = C.c_int(), C.c_int()
        lib.ora_rescore(h["qaln"].encode(), h["saln"].encode(), C.byref(om), C.byref(op), C.byref(r), C.byref(adj))
        assert r.value == h["raw_score"] and adj.value == h["score"] >= 225, h
    qi = {n: i for i, n in enumerate(qn)}
    si = {n: i for i, n in enumerate(snames)}
    key = lambda h: (qi[h["qname"]], h["qstart"], -si[h["sname"]], h["sstart"], h["qend"], h["strand"] == "C", h["send"])
    assert [key(h) for h in hits] == sorted(key(h) for h in hits)
    single = []
    for k in range(2):
        s.set_queries([qs[k]], [qn[k]])
        part, _ = s.search(p)
        single += [{kk: v for kk, v in h.items() if kk != "query_index"} for h in part]
    assert single == [{kk: v for kk, v in h.items() if kk != "query_index"} for h in hits]
    s.close()


def _rescoring_holds(snames, sseqs, qn, qs):
    import tempfile

    m = util.write_matrix("20p41g.matrix", os.path.join(tempfile.mkdtemp(), "20p41g.matrix"))
    s = engine.Session()
    s.set_matrix(m)
    s.set_library(sseqs, snames)
    s.set_queries(qs, qn)
    p = engine.make_params()
    hits, st = s.search(p)
    om = util.parse_matrix(m)
    op = util.OraParams(gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5)
    lib = util.oracle_lib()
    for h in hits:
        r, adj = C.c_int(), C.c_int()
        lib.ora_rescore(h["qaln"].encode(), h["saln"].encode(), C.byref(om), C.byref(op), C.byref(r), C.byref(adj))
        assert r.value == h["raw_score"] and adj.value == h["score"] >= 225, h
    again, _ = s.search(p)
    assert again == hits
    s.close()
    return hits, st


def test_extreme_record_and_subject_lengths():
    """Far beyond a SimpleBatcher fragment: one 900 kbp query record (8,192-column hot blocks, an int16-overflowing triple
    swept in a window) and 30 / 50 kbp subjects with 20-30 kbp alignments (raw scores > 200,000, stripe chains of ~100
    stripes, band-doubling traceback): every alignment re-scores exactly and the result is deterministic."""
    import numpy as np

    snames, sseqs = synth.make_library(501, n_families=100)
    g, _ = synth.make_genome(502, 900000, sseqs, coverage=0.4)
    hits, st = _rescoring_holds(snames, sseqs, ["big"], [g])
    assert len(hits) > 300 and max(h["raw_score"] for h in hits) > 32767
    rng = np.random.default_rng(5)
    long1 = synth.random_dna(rng, 30000).tobytes().decode()
    long2 = synth.random_dna(rng, 50000).tobytes().decode()
    bg = synth.random_dna(rng, 20000).tobytes().decode()
    q = (bg[:3000] + synth.mutate(rng, long1[2000:22000], 0.05, 0.005) + bg[3000:6000] +
         synth.revcomp(synth.mutate(rng, long2[10000:40000], 0.1, 0.01)) + bg[6000:])
    hits, st = _rescoring_holds(["L30k#LINE/L1", "L50k#Satellite/x"] + snames[:20], [long1, long2] + sseqs[:20], ["q"], [q[:74000]])
    assert sorted(h["sname"] for h in hits)[:2] == ["L30k#LINE/L1", "L50k#Satellite/x"]
    assert max(h["raw_score"] for h in hits) > 200000
