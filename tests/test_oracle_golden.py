"""The oracle against every known answer the reference holds for this path.
CPU only; runs in seconds."""
import ctypes as C
import json
import os

import pytest

import util


def test_lambda_matches_matrix_pm(tmp_path):
    """Matrix.pm:397-464 run under perl (tests/golden/make_golden.py) vs the oracle, all 46 matrices, bit-exact."""
    for name, g in util.matrices().items():
        m = util.parse_matrix(util.write_matrix(name, str(tmp_path / name)))
        assert m.alpha.decode() == g["alphabet"], name
        assert [list(r)[: m.n] for r in list(m.score)[: m.n]] == g["rows"], name
        assert list(m.freq)[: m.n] == g["freqs"], name
        assert m.lambda_ == g["lambda_"], (name, m.lambda_, g["lambda_repr"])


CASES = json.load(open(os.path.join(util.GOLD, "searchresult_t.json")))


@pytest.mark.parametrize("idx", [0, 4, 5])
def test_rescore_known_alignments(tmp_path, idx):
    """t/SearchResult.t: raw 353 -> adjusted 253 (:88-95), adjusted 238 (:316), raw 2331 -> adjusted 1530 (:319,457)."""
    c = CASES[idx]
    e = c["expect"]
    m = util.parse_matrix(util.write_matrix(c["matrixName"], str(tmp_path / "m")))
    p = util.OraParams(gap_init=e["gap_init"], ins_gap_ext=e["ins_ext"], del_gap_ext=e["del_ext"])
    raw, adj = C.c_int(), C.c_int()
    util.oracle_lib().ora_rescore(c["queryString"].encode(), c["subjString"].encode(), C.byref(m), C.byref(p),
                                  C.byref(raw), C.byref(adj))
    if "raw" in e:
        assert raw.value == e["raw"]
    if "adjusted" in e:
        assert adj.value == e["adjusted"]


@pytest.mark.parametrize("idx", [0, 5])
def test_traceback_reproduces_cross_match_alignments(tmp_path, idx):
    """The traceback tie-break (E over F over diagonal) must rebuild, column for column, the alignments that real
    cross_match 1.080812 printed (t/SearchResult.t:21-60 strand C 101 columns; :319-529 1,590 columns, 78 gaps)."""
    c = CASES[idx]
    e = c["expect"]
    m = util.parse_matrix(util.write_matrix(c["matrixName"], str(tmp_path / "m")))
    p = util.OraParams(gap_init=e["gap_init"], ins_gap_ext=e["ins_ext"], del_gap_ext=-5, bandwidth=14)
    q = c["queryString"].replace("-", "")
    s = c["subjString"].replace("-", "")
    qa = C.create_string_buffer(len(q) + len(s) + 2)
    sa = C.create_string_buffer(len(q) + len(s) + 2)
    n = util.oracle_lib().ora_global_traceback(util.encode(m, q), len(q), util.encode(m, s), len(s), C.byref(m),
                                               C.byref(p), e["raw"], qa, sa, q.encode(), s.encode())
    assert n == len(c["queryString"])
    assert qa.value.decode() == c["queryString"]
    assert sa.value.decode() == c["subjString"]
    # and the exhaustive local score of the ungapped pair equals cross_match's raw score
    assert util.oracle_lib().ora_sw_score(util.encode(m, q), len(q), util.encode(m, s), len(s), C.byref(m),
                                          C.byref(p)) == e["raw"]


KAT = json.load(open(os.path.join(util.GOLD, "seq1_2.json")))


@pytest.mark.parametrize("ml", [101, 98, 97])
def test_search_kat_seq1_2(tmp_path, ml):
    """t/seqs/masklevel/seq1-2-ml{101,98,97}.out: every hit line of real cross_match 0.990329 (score, div/del/ins,
    coordinates, '*', order) is reproduced by the exhaustive oracle."""
    (tmp_path / "q.fa").write_text(KAT["query_fasta"])
    (tmp_path / "s.fa").write_text(KAT["subject_fasta"])
    (tmp_path / "m.matrix").write_text(KAT["matrix_text"])
    p = KAT["params"]
    hits, _ = util.run_oracle(str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "m.matrix"),
                              gap_init=p["gap_init"], ins_gap_ext=p["ins_gap_ext"], del_gap_ext=p["del_gap_ext"],
                              minscore=p["minscore"], masklevel=ml, minmatch=p["minmatch"], bandwidth=p["bandwidth"])
    want = KAT[f"hits_ml{ml}"]
    assert len(hits) == len(want)
    for h, w in zip(hits, want):
        for k in ("score", "div", "dele", "ins", "qname", "qstart", "qend", "qleft", "strand", "sname", "sstart", "send",
                  "sleft", "star"):
            assert h[k] == w[k], (k, h, w)
        # alignment rows must re-derive the coordinates
        assert len(h["qaln"]) == len(h["saln"]) > 0
        assert len(h["qaln"].replace("-", "")) == h["qend"] - h["qstart"] + 1
        assert len(h["saln"].replace("-", "")) == h["send"] - h["sstart"] + 1
