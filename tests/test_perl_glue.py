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


def test_alignment_text_from_the_library_equals_reference_perl(tmp_path):
    """SURVEY 8f.1, the .cat blocks: GpuSearchResult::toStringFormatted (alignment block rendered by
    rmgpu_format_alignment) == SearchResult::toStringFormatted (SearchResult.pm:2486-2834) byte for byte, on the 244
    golden alignments, hand-made edge cases and 300 random rows; inherited formats, fallbacks, setters and clone()."""
    subprocess.run(["make", "-C", PERL_DIR], check=True, capture_output=True)
    cases = json.load(open(os.path.join(util.ROOT, "tests", "golden", "kimura.json")))["cases"]
    # the fixture holds rows as rmgpu_hit does (strand C: query reverse-complemented); the object holds the query forward
    comp = str.maketrans("ACGTYRMKHBVD", "TGCARYKMDVBH")
    rows = []
    for c in cases:
        q, s = c["qaln"], c["saln"]
        if c["strand"]:
            q, s = q[::-1].translate(comp), s[::-1].translate(comp)
        rows.append({"q": q, "s": s, "strand": c["strand"]})
    (tmp_path / "rows.json").write_text(json.dumps(rows))
    res = subprocess.run(["perl", f"-I{REF}", f"-I{PERL_DIR}", os.path.join(util.ROOT, "tests", "perl", "compare_alignment_text.pl"),
                          str(tmp_path / "rows.json")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.startswith(f"OK {2 * (len(rows) + 13 + 300)} texts"), res.stdout


def test_alignment_text_python_binding_matches_perl_binding(tmp_path):
    """The ctypes mirror calls the same C entry point as the XS glue."""
    from repeatmasker_b200 import engine
    subprocess.run(["make", "-C", PERL_DIR], check=True, capture_output=True)
    q, s = "ACGT" * 20 + "--AC" + "RYN" * 5, "ACTT" * 20 + "GGAC" + "RCN" * 5
    py = engine.format_alignment("chr1", "AluY#SINE/Alu", q, s, 101, 101 + 96, 5, 5 + 98, True, "25p41g.matrix", "3.21")
    pl = subprocess.run(["perl", f"-I{PERL_DIR}", "-MRMGpu", "-e",
                         'print RMGpu::format_alignment("chr1", "AluY#SINE/Alu", $ARGV[0], $ARGV[1], 101, 197, 5, 103, 1, '
                         '"25p41g.matrix", "3.21", undef, undef)', q, s], capture_output=True, text=True, check=True).stdout
    assert py == pl and py.startswith("\n  chr1 ") and "C AluY#SINE/Alu" in py and "Matrix = 25p41g.matrix\n" in py
    with pytest.raises(ValueError):
        engine.format_alignment("q", "s", "ACGT", "ACG", 1, 4, 1, 3)
