"""The device-side collection filters against the reference's own Perl on the reference's own fixtures
(tests/golden/masklevel.json: SearchResultCollection::maskLevelFilter / filterContainedResults, SearchResultCollection.pm:198-358
and :379-470, run under perl by tests/golden/make_golden.py on seq1-2-ml101.out and on the 3,882-hit hum-1-sine-ml101.out --
the two collections t/SearchResultCollection.t itself filters), and the engine's own masklevel (cross_match's flavour) against
the oracle on a dense collection."""
import json
import os

import pytest

import util
from repeatmasker_b200 import engine

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(util.GOLD, "masklevel.json")))["sets"]


@pytest.mark.parametrize("name", ["seq1_2", "hum1_sine"])
def test_masklevel_filter_equals_reference_perl(name):
    g = GOLD[name]
    names, starts, ends, scores = zip(*g["records"])
    for ml, want in g["masklevel"].items():
        got = engine.collection_filter(names, starts, ends, scores, int(ml))
        assert got == want, (name, ml, len(got), len(want))
    assert engine.collection_filter(names, starts, ends, scores, 101) == list(range(len(names)))


@pytest.mark.parametrize("name", ["seq1_2", "hum1_sine"])
def test_contained_filter_equals_reference_perl(name):
    g = GOLD[name]
    names, starts, ends, scores = zip(*g["records"])
    for ml, want in g["contained"].items():
        assert engine.collection_filter(names, starts, ends, scores, int(ml), contained=True) == want


def test_reference_test_expectations():
    """t/SearchResultCollection.t: 12 -> 2 at masklevel 30, -> 3 at 98; 3,882 -> 176 at 90."""
    a = GOLD["seq1_2"]
    b = GOLD["hum1_sine"]
    assert len(a["records"]) == 12 and len(a["masklevel"]["30"]) == 2 and len(a["masklevel"]["98"]) == 3
    assert len(b["records"]) == 3882 and len(b["masklevel"]["90"]) == 176


@pytest.mark.parametrize("ml", [0, 50, 80, 95, 99, 100])
def test_engine_masklevel_on_a_dense_collection(tmp_path, ml):
    """cross_match's own masklevel (engine flavour) on hundreds of overlapping Alu hits of one long query record."""
    from test_gpu_parity import assert_same_hits

    cfg = os.path.join(util.GOLD, "config1")
    m = util.write_matrix("20p43g.matrix", str(tmp_path / "m"))
    q, s = os.path.join(cfg, "NT_001493.seq"), os.path.join(cfg, "alu.fa")
    want, _ = util.run_oracle(q, s, m, masklevel=ml)
    got, st = engine.search_files(q, s, m, masklevel=ml)
    assert len(want) > 3
    assert_same_hits(got, want)
