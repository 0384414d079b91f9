"""BASELINE configuration 1: the reference's own test inputs (copies under tests/golden/config1, made by
tests/golden/make_golden.py): t/seqs/general/{hum-1.fa,small-1.fa,NT_001493.seq} against t/libraries/{alu.fa,sinecutlib}.

  * the oracle's SIMD prefilter against its scalar pass (CPU);
  * concordance with REAL cross_match: the 3,882 hit lines of t/seqs/masklevel/hum-1-sine-ml101.out (cross_match 0.990329 on
    hum-1.fa; the subject library of that 2005 run is gone, today's sinecutlib / alu.fa hold 18 of its 22 families) --
    the measured floor of tools/concordance_hum1_sine.py, profiles/r2_concordance_hum1_sine.txt (CPU: the oracle on the 7
    families of sinecutlib; GPU: librmgpu on all 18);
  * librmgpu == oracle, hit for hit with alignments, on these inputs (GPU);
  * the reference driver with both engines on the invocations of t/RepeatMasker.t (-gc 43, -xsmall, -x, -qq) and without -gc:
    .out/.masked byte-identical (GPU)."""
import collections
import ctypes as C
import json
import os
import random
import subprocess

import pytest

import util

CFG1 = os.path.join(util.GOLD, "config1")


def library(tmp_path, name):
    """A library of the reference tree; "alu+ca" adds the (CA)n consensus that masks small-1.fa in t/RepeatMasker.t (there it
    comes from the species library, which the reference checkout does not hold)."""
    if name != "alu+ca":
        return os.path.join(CFG1, name)
    out = str(tmp_path / "alu_ca.fa")
    open(out, "w").write(open(os.path.join(CFG1, "alu.fa")).read() + ">(CA)n#Simple_repeat\n" + "CA" * 90 + "\n")
    return out
FX = json.load(open(os.path.join(util.GOLD, "hum1_sine.json")))
REF = [dict(zip(FX["keys"], h)) for h in FX["hits"]]
KW = dict(gap_init=-25, ins_gap_ext=-6, del_gap_ext=-6, minscore=250, masklevel=101, minmatch=10, bandwidth=14)


def fixture_matrix(tmp_path):
    m = str(tmp_path / "comparison.matrix")
    open(m, "w").write(FX["matrix_text"])
    return m


def concordance(got, families):
    """(identical lines, same coordinates, located at all, total) of the fixture's hits on `families`."""
    exact = set((g["qname"], g["qstart"], g["qend"], g["strand"], g["sname"], g["sstart"], g["send"], g["score"], g["div"],
                 g["dele"], g["ins"]) for g in got)
    coords = set((g["qname"], g["qstart"], g["qend"], g["strand"], g["sname"], g["sstart"], g["send"]) for g in got)
    by = collections.defaultdict(list)
    for g in got:
        by[(g["qname"], g["sname"], g["strand"])].append((g["qstart"], g["qend"]))
    ident = same = located = total = 0
    for r in REF:
        if r["sname"] not in families:
            continue
        total += 1
        k = (r["qname"], r["qstart"], r["qend"], r["strand"], r["sname"], r["sstart"], r["send"])
        if k + (r["score"], r["div"], r["dele"], r["ins"]) in exact:
            ident += 1
        if k in coords:
            same += 1
        if any(a <= r["qend"] and b >= r["qstart"] for a, b in by[(r["qname"], r["sname"], r["strand"])]):
            located += 1
    return ident, same, located, total


SINECUT_FAMILIES = {"AluJb#SINE/Alu", "AluSx#SINE/Alu", "AluY#SINE/Alu", "FLAM_C#SINE/Alu", "FRAM#SINE/Alu", "BC200#SINE/Alu",
                    "7SLRNA#SINE/Alu"}


# ------------------------------------------------------------------------------------------------------------ CPU
def test_oracle_batch_prefilter_equals_scalar_pass(tmp_path):
    lib = util.oracle_lib()
    if not hasattr(lib, "ora_sw_score_batch"):
        pytest.skip("oracle built without the AVX2 prefilter")
    m = util.parse_matrix(util.write_matrix("20p43g.matrix", str(tmp_path / "m")))
    rng = random.Random(5)
    for trial in range(6):
        p = util.OraParams(gap_init=-30 + trial, ins_gap_ext=-6, del_gap_ext=-5, minscore=225)
        Q = rng.randint(1, 900)
        q = bytes(rng.choice(b"ACGTNRYX") for _ in range(Q))
        subs = []
        for l in range(rng.randint(1, 16)):
            L = rng.randint(1, 700)
            s = bytearray(rng.choice(b"ACGT") for _ in range(L))
            if l % 3 == 0 and L > 50 and Q > 50:  # a planted copy so that some lanes score high
                n = min(L, Q) - 10
                s[5:5 + n] = q[3:3 + n]
            subs.append(bytes(s))
        qc = util.encode(m, q.decode())
        sc = [util.encode(m, s.decode()) for s in subs]
        arr = (C.c_char_p * len(sc))(*sc)
        lens = (C.c_int * len(sc))(*[len(s) for s in sc])
        out = (C.c_int * 16)()
        lib.ora_sw_score_batch(qc, Q, arr, lens, len(sc), C.byref(m), C.byref(p), out)
        for l, s in enumerate(sc):
            want = lib.ora_sw_score(qc, Q, s, len(s), C.byref(m), C.byref(p))
            assert out[l] == min(want, 32000), (trial, l, out[l], want)


def test_oracle_output_identical_with_and_without_simd_prefilter(tmp_path):
    m = util.write_matrix("20p43g.matrix", str(tmp_path / "m"))
    argv = util.oracle_argv(os.path.join(CFG1, "NT_001493.seq"), os.path.join(CFG1, "sinecutlib"), m, minscore=180)
    a = subprocess.run(argv, capture_output=True, text=True, check=True).stdout
    b = subprocess.run(argv, capture_output=True, text=True, check=True, env=dict(os.environ, RM_ORACLE_SCALAR="1")).stdout
    assert a == b and len(util.parse_crossmatch(a)) > 10


def test_oracle_concordance_with_real_cross_match(tmp_path):
    """The oracle against the hit lines of real cross_match on the 7 families today's sinecutlib still holds (1,162 of the
    3,882 fixture hits).  Measured (profiles/r2_concordance_hum1_sine.txt): 528 identical lines, 785 at identical
    coordinates (257 of them with a score exactly 9 higher: the first base of those consensi changed since 2005), 1,154
    located; 299 are parts of a longer alignment of ours (exhaustive DP merges what the 14-wide band splits)."""
    got, _ = util.run_oracle(os.path.join(CFG1, "hum-1.fa"), os.path.join(CFG1, "sinecutlib"), fixture_matrix(tmp_path), **KW)
    ident, same, located, total = concordance(got, SINECUT_FAMILIES)
    assert total == 1162
    assert ident >= 528 and same >= 785 and located >= 1154, (ident, same, located)


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("query,subject,matrix,kw", [
    ("hum-1.fa", "sinecutlib", None, KW),
    ("NT_001493.seq", "alu.fa", "20p43g.matrix", dict(minscore=225)),
    ("small-1.fa", "alu+ca", "18p43g.matrix", dict(minscore=180, gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5)),
    ("hum-1.fa", "alu.fa", "25p43g.matrix", dict(minscore=300, gap_init=-27, ins_gap_ext=-6, del_gap_ext=-5, raw=True)),
])
def test_gpu_equals_oracle_on_reference_inputs(tmp_path, query, subject, matrix, kw):
    from repeatmasker_b200 import engine
    from test_gpu_parity import assert_same_hits

    m = fixture_matrix(tmp_path) if matrix is None else util.write_matrix(matrix, str(tmp_path / matrix))
    q, s = os.path.join(CFG1, query), library(tmp_path, subject)
    want, _ = util.run_oracle(q, s, m, **kw)
    got, st = engine.search_files(q, s, m, **kw)
    assert len(want) > 0
    assert_same_hits(got, want)


@pytest.mark.gpu
def test_gpu_concordance_with_real_cross_match(tmp_path):
    """librmgpu against real cross_match on the 18 families for which the reference tree still holds a sequence (3,143 of the
    3,882 fixture hits; sinecutlib + the families only alu.fa has).  Measured floor: 1,621 identical lines (1,161 of 1,392 =
    83 % on the families whose consensus did not change: AluYa5/Ya8/Yb9/Yd8/Yg6/Yh9, FRAM, BC200, 7SLRNA; most of the rest
    are alignments the exhaustive pass extends through a gap that cross_match's 14-wide band splits), 2,203 at identical
    coordinates, 3,135 located."""
    from repeatmasker_b200 import engine

    lib = util.union_library(os.path.join(CFG1, "sinecutlib"), os.path.join(CFG1, "alu.fa"), str(tmp_path / "union.fa"))
    got, st = engine.search_files(os.path.join(CFG1, "hum-1.fa"), lib, fixture_matrix(tmp_path), **KW)
    for g in got:
        g["strand"] = "C" if g["strand"] in (1, "C") else "+"
    names = set(util.read_fasta(lib)[0])
    fams = set(r["sname"] for r in REF if r["sname"] in names)
    assert len(fams) == 18
    ident, same, located, total = concordance(got, fams)
    assert total == 3143
    assert ident >= 1621 and same >= 2203 and located >= 3135, (ident, same, located)
    unchanged = {f"{n}#SINE/Alu" for n in ("AluYa5", "AluYa8", "AluYb9", "AluYd8", "AluYg6", "AluYh9", "FRAM", "BC200", "7SLRNA")}
    i2, s2, l2, t2 = concordance(got, unchanged)
    assert t2 == 1392 and i2 >= 1161, (i2, t2)


@pytest.mark.gpu
@pytest.mark.skipif(not util.harness_ready(), reason="baseline/_ref harness not present (tools/make_rm_harness.sh)")
@pytest.mark.parametrize("query,lib,args", [
    ("hum-1.fa", "alu.fa", ["-gc", "43", "-qq"]),        # t/RepeatMasker.t test 5 (there with the human library)
    ("small-1.fa", "alu+ca", ["-gc", "43", "-xsmall"]),  # test 2
    ("NT_001493.seq", "sinecutlib", ["-gc", "43", "-x"]),  # test 3's flag on the third reference query
    ("hum-1.fa", "sinecutlib", []),                      # no -gc: chooseMatrices picks a GC bin per batch (RepeatMasker:4784-4821)
])
def test_reference_driver_on_reference_inputs(tmp_path, query, lib, args):
    outs = {}
    for eng in ("crossmatch", "gpu"):
        d = tmp_path / eng
        d.mkdir()
        (d / "genome.fa").write_text(open(os.path.join(CFG1, query)).read())
        (d / "lib.fa").write_text(open(library(tmp_path, lib)).read())
        log = util.run_repeatmasker(str(d), eng, args)
        if eng == "gpu":
            assert "Search Engine: rmgpu" in log
        outs[eng] = util.driver_outputs(str(d))
    a, b = outs["crossmatch"], outs["gpu"]
    assert a["out"] is not None and a["out"].count("\n") > 3
    assert a["out"] == b["out"]
    assert a["masked"] == b["masked"]
    assert util.strip_engine_trailer(a["cat"]) == util.strip_engine_trailer(b["cat"])
    assert util.strip_engine_trailer(a["tbl"]) == util.strip_engine_trailer(b["tbl"])
