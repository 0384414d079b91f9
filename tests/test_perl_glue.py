"""The Perl side of the boundary, checked on the CPU in the build container (needs perl and the reference's Perl
modules under /root/reference; skipped elsewhere): GpuSearchEngine::recordsToCollection must yield exactly what
the reference's CrossmatchSearchEngine->search() yields for the same search."""
import json
import os
import shutil
import subprocess

import pytest

import util
from repeatmasker_b200 import synth

REF = "/root/reference"
PERL_DIR = os.path.join(util.ROOT, "repeatmasker_b200", "perl")

pytestmark = pytest.mark.skipif(not (shutil.which("perl") and os.path.exists(os.path.join(REF, "CrossmatchSearchEngine.pm"))),
                                reason="needs perl and the reference's Perl modules")


def test_xs_glue_builds_and_loads():
    subprocess.run(["make", "-C", PERL_DIR], check=True, capture_output=True)
    out = subprocess.run(["perl", f"-I{PERL_DIR}", "-MRMGpu", "-e", "print RMGpu::version()"], capture_output=True, text=True,
                         check=True).stdout
    assert out.startswith("rmgpu version")


@pytest.mark.parametrize("library_kimura", [False, True])
def test_records_to_collection_equals_reference_parse(tmp_path, library_kimura):
    subprocess.run(["make", "-C", PERL_DIR], check=True, capture_output=True)
    snames, sseqs = synth.make_library(21, n_families=20, lo=80, hi=900)
    genome, _ = synth.make_genome(22, 9000, sseqs, coverage=0.5)
    q, s, m = str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "20p43g.matrix")
    synth.write_fasta(q, ["chrP"], [genome])
    synth.write_fasta(s, snames, sseqs)
    util.write_matrix("20p43g.matrix", m)
    hits, _ = util.run_oracle(q, s, m)
    assert any(h["strand"] == "C" for h in hits) and any(h["strand"] == "+" for h in hits)
    recs = [dict(score=h["score"], pctDiverge=h["div"], pctDelete=h["dele"], pctInsert=h["ins"], queryName=h["qname"],
                 queryStart=h["qstart"], queryEnd=h["qend"], queryRemaining=h["qleft"],
                 orientation="C" if h["strand"] == "C" else "", subjName=h["sname"], subjStart=h["sstart"],
                 subjEnd=h["send"], subjRemaining=h["sleft"], star=int(h["star"]), queryString=h["qaln"],
                 subjString=h["saln"]) for h in hits]
    if library_kimura:  # the fast path: the text computed by librmgpu (host code) instead of the Perl loop
        from repeatmasker_b200 import engine
        for r, h in zip(recs, hits):
            r["pctKimuraDiverge"] = engine.kimura_divergence(h["qaln"], h["saln"], h["strand"] == "C")[0]
    (tmp_path / "recs.json").write_text(json.dumps(recs))
    res = subprocess.run(["perl", f"-I{REF}", f"-I{PERL_DIR}", os.path.join(util.ROOT, "tests", "perl", "compare_collections.pl"),
                          util.ORACLE_BIN, m, q, s, str(tmp_path / "recs.json")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip() == f"OK {len(hits)}"
