"""GPU parity tests: librmgpu (through the C ABI) against the CPU oracle and the committed golden fixtures.
Bit-exact on every reported field.  Run on the B200 box: python -m pytest tests -m gpu"""
import ctypes as C
import json
import os
import random

import numpy as np
import pytest

import util
from repeatmasker_b200 import engine, synth

pytestmark = pytest.mark.gpu

FIELDS = ("score", "div", "dele", "ins", "qname", "qstart", "qend", "qleft", "strand", "sname", "sstart", "send",
          "sleft", "star")


def assert_same_hits(got, want, aln=True):
    assert len(got) == len(want), (len(got), len(want))
    for g, w in zip(got, want):
        for k in FIELDS:
            assert g[k] == w[k], (k, g, w)
        if aln:
            assert g["qaln"] == w["qaln"] and g["saln"] == w["saln"], (g, w)
            if "ti" in w:
                assert (g["ti"], g["tv"]) == (w["ti"], w["tv"])


def both(tmp_path, qnames, qseqs, snames, sseqs, matrix="20p43g.matrix", **kw):
    q, s, m = str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "m.matrix")
    synth.write_fasta(q, qnames, qseqs)
    synth.write_fasta(s, snames, sseqs)
    if os.path.exists(matrix):
        m = matrix
    else:
        util.write_matrix(matrix, m)
    want, _ = util.run_oracle(q, s, m, **kw)
    got, st = engine.search_files(q, s, m, **kw)
    return got, want, st


KAT = json.load(open(os.path.join(util.GOLD, "seq1_2.json")))


@pytest.mark.parametrize("ml", [101, 98, 97])
def test_kat_seq1_2_against_real_cross_match_fixture(tmp_path, ml):
    """t/seqs/masklevel/seq1-2-ml*.out: the hit lines real cross_match printed, through the GPU engine."""
    (tmp_path / "q.fa").write_text(KAT["query_fasta"])
    (tmp_path / "s.fa").write_text(KAT["subject_fasta"])
    (tmp_path / "m.matrix").write_text(KAT["matrix_text"])
    p = KAT["params"]
    kw = dict(gap_init=p["gap_init"], ins_gap_ext=p["ins_gap_ext"], del_gap_ext=p["del_gap_ext"], minscore=p["minscore"],
              masklevel=ml, minmatch=p["minmatch"], bandwidth=p["bandwidth"])
    got, st = engine.search_files(str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "m.matrix"), **kw)
    assert_same_hits(got, KAT[f"hits_ml{ml}"], aln=False)
    want, _ = util.run_oracle(str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "m.matrix"), **kw)
    assert_same_hits(got, want)
    assert st["launches"] >= 3 and st["cells"] == 2 * 2604 * 2797


def test_score_pass_equals_oracle_best(tmp_path):
    """K1 (int16x2 DPX) best score per (query, subject, strand) == the oracle's exhaustive best, ragged sizes."""
    rng = np.random.default_rng(7)
    m = util.write_matrix("20p43g.matrix", str(tmp_path / "m.matrix"))
    om = util.parse_matrix(m)
    snames, sseqs = synth.make_library(11, n_families=40, lo=20, hi=3000)
    sseqs[3] = sseqs[3][:7]  # shorter than one row group
    sseqs[5] = sseqs[5] * 3 if len(sseqs[5]) < 1500 else sseqs[5]
    genome, _ = synth.make_genome(12, 6000, sseqs, coverage=0.5)
    qseqs = [genome[:3111], genome[3111:3143], genome[3143:], "ACGT"]
    s = engine.Session()
    s.set_matrix(m)
    s.set_library(sseqs, snames)
    s.set_queries(qseqs)
    p = engine.make_params()
    best, ms = s.score_only(p)
    op = util.OraParams(gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5)
    lib = util.oracle_lib()
    for qi, q in enumerate(qseqs):
        for strand, ql in enumerate((q, synth.revcomp(q))):
            qc = util.encode(om, ql)
            for si, sq in enumerate(sseqs):
                want = lib.ora_sw_score(qc, len(qc), util.encode(om, sq), len(sq), C.byref(om), C.byref(op))
                assert best[(qi * len(sseqs) + si) * 2 + strand] == want, (qi, si, strand)
    s.close()


def test_synthetic_genome_full_parity(tmp_path):
    """Implanted diverged copies, both strands, several fragments: every hit field and alignment row identical."""
    snames, sseqs = synth.make_library(101, n_families=60, lo=60, hi=2500)
    genome, truth = synth.make_genome(102, 30000, sseqs, coverage=0.45)
    qn, qs = synth.fragment(genome, "chrS", frag=12000, overlap=500)
    got, want, st = both(tmp_path, qn, qs, snames, sseqs)
    assert len(want) > 10
    assert_same_hits(got, want)
    assert any(h["strand"] == "C" for h in got) and any(h["strand"] == "+" for h in got)


def test_long_subject_chains_and_int16_overflow(tmp_path):
    """A 9 kbp subject (35 stripes of 256 rows) and a near-identical 6 kbp copy: raw score > 32767 forces the
    int32 pass; stripe-to-stripe boundaries must be exact."""
    rng = np.random.default_rng(5)
    long_s = synth.random_dna(rng, 9000).tobytes().decode()
    other = synth.random_dna(rng, 700).tobytes().decode()
    copy = synth.mutate(rng, long_s[1500:7600], 0.02, 0.002)
    bg = synth.random_dna(rng, 9000).tobytes().decode()
    q = bg[:2000] + copy + bg[2000:4000] + synth.revcomp(synth.mutate(rng, other, 0.15, 0.02)) + bg[4000:]
    got, want, st = both(tmp_path, ["q1"], [q], ["longS#LINE/L1", "short#SINE/x"], [long_s, other])
    assert max(h["score"] for h in want) > 32767
    assert_same_hits(got, want)


def test_int16_overflow_windows_start_at_a_checkpoint(tmp_path):
    """Overflowing (> 32767) alignments in the middle of a 50 kbp query, one per strand: the int32 sweep of such a triple
    starts at the K1 checkpoint before its first hot column block and stops after the last one; no redo is needed."""
    rng = np.random.default_rng(77)
    long_s = synth.random_dna(rng, 9000).tobytes().decode()
    other = synth.random_dna(rng, 700).tobytes().decode()
    copy1 = synth.mutate(rng, long_s[1000:7400], 0.02, 0.002)
    copy2 = synth.revcomp(synth.mutate(rng, long_s[2000:8600], 0.03, 0.002))
    bg = synth.random_dna(rng, 38000).tobytes().decode()
    q = (bg[:12000] + copy1 + bg[12000:13000] + synth.mutate(rng, other, 0.12, 0.01) + bg[13000:22000] + copy2 +
         bg[22000:30000] + synth.mutate(rng, long_s[4000:4600], 0.2, 0.02) + bg[30000:])
    got, want, st = both(tmp_path, ["q1"], [q], ["longS#LINE/L1", "short#SINE/x"], [long_s, other])
    assert sum(h["score"] > 32767 for h in want) == 2
    assert_same_hits(got, want)
    assert st["n_origin_redo"] == 0


def test_int16_overflow_at_both_ends_of_the_query(tmp_path):
    """Two overflowing alignments of the same (query, subject, strand) triple, one starting at the first base of the
    query and one ending at its last base: the int32 window has no checkpoint to start from and reaches the end."""
    rng = np.random.default_rng(78)
    long_s = synth.random_dna(rng, 8200).tobytes().decode()
    copy1 = synth.mutate(rng, long_s[500:7000], 0.02, 0.002)
    copy2 = synth.mutate(rng, long_s[1200:8000], 0.03, 0.003)
    bg = synth.random_dna(rng, 12000).tobytes().decode()
    q = copy1 + bg + copy2
    got, want, st = both(tmp_path, ["q1"], [q], ["longS#LINE/L1"], [long_s])
    assert sum(h["score"] > 32767 for h in want) == 2
    assert_same_hits(got, want)


def test_int16_overflow_fan_leaves_the_window(tmp_path):
    """A 100,000-point alignment (12 kbp subject) followed by 40 kbp of background: the decaying fan to its right (-6 per
    column) is still above minscore where the stretched window of the int32 score pass (K2w) ends, so the pass raises
    its edge flag and the triple takes the one-sweep path (with that path's own full-width redo); a second, shorter
    wrapped copy on the other strand of a second subject stays on the K2w + block-parallel path.  Subject lengths are
    not multiples of the 64-row stripes."""
    rng = np.random.default_rng(79)
    long_s = synth.random_dna(rng, 12011).tobytes().decode()
    long_t = synth.random_dna(rng, 6521).tobytes().decode()
    copy1 = synth.mutate(rng, long_s[200:11800], 0.01, 0.001)
    copy2 = synth.revcomp(synth.mutate(rng, long_t[100:6400], 0.05, 0.004))
    bg = synth.random_dna(rng, 52000).tobytes().decode()
    q = bg[:1500] + copy1 + bg[1500:41500] + copy2 + bg[41500:44000]
    got, want, st = both(tmp_path, ["q1"], [q], ["longS#LINE/L1", "longT#LINE/L1"], [long_s, long_t])
    assert sum(h["score"] > 32767 for h in want) == 2 and max(h["score"] for h in want) > 95000
    assert_same_hits(got, want)
    assert st["n_origin_redo"] >= 1


def test_many_wrapped_triples_share_one_int32_pass(tmp_path):
    """Six query records, each with a > 32767-point copy of one of three long subjects at a different offset (first
    column block, middle, last block) and strand: all wrapped triples go through one K2w launch and one K2b launch."""
    rng = np.random.default_rng(80)
    subj = [synth.random_dna(rng, n).tobytes().decode() for n in (7000, 7777, 8191)]
    qn, qs = [], []
    for i in range(6):
        sq = subj[i % 3]
        cp = synth.mutate(rng, sq[50:len(sq) - 50], 0.02 + 0.01 * (i % 3), 0.002)
        if i & 1:
            cp = synth.revcomp(cp)
        bg = synth.random_dna(rng, 16000).tobytes().decode()
        off = (0, 7000, 16000)[i % 3] if i < 3 else (16000, 0, 7000)[i % 3]
        qn.append(f"q{i}")
        qs.append(bg[:off] + cp + bg[off:])
    got, want, st = both(tmp_path, qn, qs, [f"L{i}#LINE/L1" for i in range(3)], subj)
    assert sum(h["score"] > 32767 for h in want) == 6
    assert_same_hits(got, want)


def test_traceback_band_retries(tmp_path):
    """Alignments whose path leaves the first traceback band W = bandwidth + max(n, m) / 32 (an insertion and, further
    on, a deletion of the same size: the end points stay on the main diagonal, the path does not) are found with the
    doubled or quadrupled band; both teams of the traceback kernel (one warp, a block), with one, two and -- in the
    block team -- more cells per thread and wave, and the host-side reschedule after two failed widths."""
    rng = np.random.default_rng(4242)
    snames, sseqs, parts = [], [], []
    # (subject length, excursion): W = 14 + L/32; the excursion decides how many doublings are needed
    for i, (L, ex) in enumerate([(400, 40), (700, 60), (1100, 110), (3000, 150), (6000, 300), (8000, 400), (8000, 600), (5000, 700)]):
        sq = synth.random_dna(rng, L).tobytes().decode()
        snames.append(f"fam{i}#DNA/x")
        sseqs.append(sq)
        a, b = L // 3, 2 * L // 3
        ins = synth.random_dna(rng, ex).tobytes().decode()
        cp = synth.mutate(rng, sq[:a], 0.03, 0.0) + ins + synth.mutate(rng, sq[a:b], 0.03, 0.0) + synth.mutate(rng, sq[b + ex:], 0.03, 0.0)
        parts.append(synth.random_dna(rng, 300).tobytes().decode() + (cp if i % 2 == 0 else synth.revcomp(cp)))
    q = "".join(parts) + synth.random_dna(rng, 300).tobytes().decode()
    got, want, st = both(tmp_path, ["q1"], [q], snames, sseqs)
    # the flanks outscore the two gaps: each copy is one alignment through both
    assert sum(h["qend"] - h["qstart"] > 0.9 * len(sseqs[int(h["sname"][3])]) for h in want) >= 6
    assert_same_hits(got, want)


def test_library_packing_edge_lengths(tmp_path):
    """Subject lengths around the packing units (16-row groups, 512-row stripes, 4-stripe runs): subjects that end
    exactly on a group / stripe / run boundary, one-base and 15-base subjects, a subject longer than a run whose last
    stripe is topped up, and enough small ones to share stripes with the tails of longer ones."""
    rng = np.random.default_rng(123)
    lengths = [1, 15, 16, 17, 31, 32, 100, 511, 512, 513, 1024, 1040, 2047, 2048, 2049, 2560, 3000, 5000, 240, 250, 260, 270]
    seqs = [synth.random_dna(rng, n).tobytes().decode() for n in lengths]
    names = [f"s{i}_{n}#DNA/x" for i, n in enumerate(lengths)]
    bg = synth.random_dna(rng, 30000).tobytes().decode()
    parts, at = [], 0
    for i, sq in enumerate(seqs):
        if len(sq) < 60:
            continue
        copy = synth.mutate(rng, sq, 0.08 + 0.01 * (i % 5), 0.01)
        if i % 2:
            copy = synth.revcomp(copy)
        parts.append(bg[at:at + 600] + copy)
        at += 600
    q = "".join(parts) + bg[at:at + 3000]
    got, want, st = both(tmp_path, ["qpack"], [q], names, seqs)
    assert len(want) >= 15
    assert_same_hits(got, want)


def test_alphabet_edge_cases(tmp_path):
    """N runs, X (masked), lower-case x (excision sentinel), IUPAC codes incl. B/D/H/V, >7 residue classes."""
    rng = np.random.default_rng(9)
    cons = synth.random_dna(rng, 600).tobytes().decode()
    c2 = synth.random_dna(rng, 400).tobytes().decode()
    cons_amb = cons[:100] + "R" + cons[101:200] + "Y" + cons[201:300] + "N" + cons[301:]
    copy = list(synth.mutate(rng, cons, 0.1, 0.01))
    for pos, ch in ((50, "N"), (51, "N"), (120, "R"), (200, "K"), (260, "M"), (300, "S"), (310, "W"), (320, "B"),
                    (330, "D"), (340, "H"), (350, "V"), (400, "X")):
        copy[pos] = ch
    copy = "".join(copy)
    bg = synth.random_dna(rng, 3000).tobytes().decode()
    q = bg[:700] + copy + "NNNNNNNNNNNNNNNNNNNN" + bg[700:1500] + "x" + synth.mutate(rng, c2, 0.12, 0.01) + "XXXXXXXXXX" + bg[1500:]
    got, want, st = both(tmp_path, ["amb"], [q], ["cons#DNA/x", "c2#LTR/y"], [cons_amb, c2])
    assert len(want) >= 2
    assert_same_hits(got, want)


def test_letters_outside_the_iupac_alphabet(tmp_path):
    """Letters that are no nucleotide code (U, I, E ...) are scored as N, like the oracle does, but the rows show the original
    letter and a mismatch is a difference of LETTERS (U against N counts): the device spells such records with N and the
    host puts the letters back and recounts (rmgpu_session::respell_exotic)."""
    rng = np.random.default_rng(19)
    cons = synth.random_dna(rng, 500).tobytes().decode()
    copy = list(synth.mutate(rng, cons, 0.08, 0.01))
    for pos, ch in ((40, "U"), (41, "U"), (150, "I"), (300, "E"), (301, "N")):
        copy[pos] = ch
    cons_x = cons[:90] + "U" + cons[91:250] + "N" + cons[251:]
    bg = synth.random_dna(rng, 2000).tobytes().decode()
    q = bg[:500] + "".join(copy) + bg[500:1200] + synth.revcomp(synth.mutate(rng, cons, 0.1, 0.01)) + bg[1200:]
    got, want, st = both(tmp_path, ["exotic", "plain"], [q, bg[:300] + synth.mutate(rng, cons, 0.05, 0.0) + bg[300:600]],
                         ["consU#DNA/x"], [cons_x])
    assert len(want) >= 3
    assert any("U" in h["qaln"] for h in want) and any("U" in h["saln"] for h in want)
    assert_same_hits(got, want)


# the score pass has an instantiation per gap-penalty set of the crossmatch recipes (RepeatMasker:2127-2480), with the
# penalties as immediates, and a generic one (last two rows: penalties no recipe uses, del_gap_ext below gap_init)
@pytest.mark.parametrize("matrix,gi,ie,de", [("14p41g.matrix", -35, -7, -6), ("18p41g.matrix", -30, -6, -5),
                                            ("25p41g.matrix", -27, -6, -5), ("20p41g.matrix", -33, -5, -4),
                                            ("20p41g.matrix", -27, -5, -5), ("20p41g.matrix", -40, -15, -15),
                                            ("18p41g.matrix", -35, -10, -10),
                                            ("20p41g.matrix", -28, -4, -3), ("18p41g.matrix", -25, -5, -30)])
@pytest.mark.parametrize("raw", [False, True])
def test_matrix_sweep(tmp_path, matrix, gi, ie, de, raw):
    """BASELINE config 4 in miniature: GC-binned matrices x recipes (RepeatMasker:2146-2366), adjusted and -raw,
    with simple-repeat / low-complexity content."""
    rng = np.random.default_rng(31)
    snames, sseqs = synth.make_library(33, n_families=25, lo=80, hi=1200)
    sseqs.append("CA" * 90)
    snames.append("(CA)n#Simple_repeat")
    sseqs.append("A" * 150)
    snames.append("(A)n#Simple_repeat")
    genome, _ = synth.make_genome(34, 12000, sseqs[:25], coverage=0.4)
    g = genome[:5000] + "CA" * 40 + "CG" + "CA" * 30 + genome[5000:9000] + "A" * 70 + "G" + "A" * 40 + genome[9000:]
    got, want, st = both(tmp_path, ["gcblock"], [g], snames, sseqs, matrix=matrix, gap_init=gi, ins_gap_ext=ie,
                         del_gap_ext=de, minscore=180, raw=raw)
    assert len(want) > 5
    assert_same_hits(got, want)


def test_empty_and_tiny_inputs(tmp_path):
    got, want, _ = both(tmp_path, ["tiny"], ["ACGTACGTAC"], ["s"], ["ACGTTTTACG"])
    assert got == [] and want == []
    # a query record with no bases and a subject with no bases
    q, s, m = str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "m.matrix")
    open(q, "w").write(">empty\n>one\nACGTACGTAGCTAGCTAGCTAGCATCGATCGATCGATCGTAGCTAGCTAGCTAGCTAGTCGATCGATGCTAGCTAGCTAGC\n")
    open(s, "w").write(">nothing\n>same\nACGTACGTAGCTAGCTAGCTAGCATCGATCGATCGATCGTAGCTAGCTAGCTAGCTAGTCGATCGATGCTAGCTAGCTAGC\n")
    util.write_matrix("20p43g.matrix", m)
    want, _ = util.run_oracle(q, s, m, minscore=100)
    got, _ = engine.search_files(q, s, m, minscore=100)
    assert len(want) >= 1 and all(h['qname'] == 'one' and h['sname'] == 'same' for h in want)
    assert_same_hits(got, want)


def test_python_mirror_of_engine_interface(tmp_path):
    """The SearchEngineI-style call sequence RepeatMasker::search makes (RepeatMasker:2022-2066)."""
    snames, sseqs = synth.make_library(5, n_families=10, lo=100, hi=600)
    genome, _ = synth.make_genome(6, 5000, sseqs, coverage=0.5)
    q, s, m = str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "m.matrix")
    synth.write_fasta(q, ["g"], [genome])
    synth.write_fasta(s, snames, sseqs)
    util.write_matrix("20p43g.matrix", m)
    e = engine.GpuSearchEngine()
    e.setQuery(q); e.setSubject(s); e.setMatrix(m); e.setWordRaw(0)
    e.setScoreMode(engine.GpuSearchEngine.complexityAdjustedScoreMode); e.setGenerateAlignments(1)
    e.setGapInit(-30); e.setInsGapExt(-6); e.setDelGapExt(-5); e.setMaskLevel(101); e.setMinScore(225)
    status, hits = e.search(minMatch=9, bandwidth=14)
    want, _ = util.run_oracle(q, s, m)
    assert status == 0
    assert_same_hits(hits, want)
