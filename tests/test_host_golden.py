"""The double-precision host pieces of librmgpu, through its C ABI and without a GPU, against the same golden vectors
the oracle is pinned to: the matrix reader and lambda (Matrix.pm:308-364,397-464 run under perl, all 46 matrices,
bit-exact) and the complexity adjustment (t/SearchResult.t: raw 353 -> 253, 2331 -> 1530)."""
import json
import os

import pytest

import util
from repeatmasker_b200 import engine

CASES = json.load(open(os.path.join(util.GOLD, "searchresult_t.json")))


def test_matrix_reader_and_lambda_match_matrix_pm(tmp_path):
    mats = util.matrices()
    assert len(mats) == 46
    for name, g in mats.items():
        n, alphabet, lam, freqs = engine.matrix_info(util.write_matrix(name, str(tmp_path / name)))
        assert (n, alphabet) == (len(g["alphabet"]), g["alphabet"]), name
        assert lam == g["lambda_"], (name, lam, g["lambda_"])  # bit-exact: same doubling + bisection loop


@pytest.mark.parametrize("idx", [0, 5])
def test_complexity_adjustment_of_known_alignments(tmp_path, idx):
    c = CASES[idx]
    e = c["expect"]
    m = util.write_matrix(c["matrixName"], str(tmp_path / "m"))
    assert engine.adjusted_score(m, e["raw"], c["queryString"], c["subjString"]) == e["adjusted"]


def test_matrix_reader_fails_loudly(tmp_path):
    with pytest.raises(engine.RmgpuError, match="does not exist"):
        engine.matrix_info(str(tmp_path / "missing.matrix"))
    (tmp_path / "bad.matrix").write_text("this is not a matrix\n")
    with pytest.raises(engine.RmgpuError, match="could not parse"):
        engine.matrix_info(str(tmp_path / "bad.matrix"))


def test_fragment_arithmetic_equals_simplebatcher():
    """synth.fragment (what bench.py and the full-size tests cut genomes with) against the reference's own SimpleBatcher
    (tests/golden/simplebatcher.json: ids, lengths and content digests of its batches, made by running it under perl)."""
    import hashlib
    import json
    import os
    import random

    from repeatmasker_b200 import synth

    gold = json.load(open(os.path.join(util.GOLD, "simplebatcher.json")))
    for case in gold["cases"]:
        rng = random.Random(case["seed"])
        seq = "".join(rng.choice("ACGT") for _ in range(case["length"]))
        names, seqs = synth.fragment(seq, "s", frag=60000, overlap=2000)
        got = [[n, len(x), hashlib.md5(x.encode()).hexdigest()] for n, x in zip(names, seqs)]
        assert got == case["fragments"], case["length"]
        offs = synth.fragment_offsets(case["length"])
        assert all(seq[o:o + len(x)] == x for o, x in zip(offs, seqs))
