#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ from the reference tree.

Run in the BUILD container only (needs /root/reference, read-only, and perl):
    python tests/golden/make_golden.py

Nothing under tests/ or the product reads /root/reference at run time; the fixtures
written here are what travels to the GPU box.  Sources of each fixture (file:line in
/root/reference) are recorded inside the JSON.

Fixtures:
  searchresult_t.json   alignments + expected scores from t/SearchResult.t (cross_match
                        1.080812 output pasted into the reference's own unit test)
  seq1_2.json           the only self-contained search KAT: t/seqs/masklevel/seq-1.fa vs seq-2.fa
                        with its cross_match 0.990329 hit lists at masklevel 101/98/97
  (matrices.json also holds) Matrix.pm lambda for every Matrices/crossmatch/*.matrix, produced by
                        running the reference's own Matrix.pm under perl
  config1/              BASELINE configuration 1 inputs, copied verbatim: t/seqs/general/{hum-1.fa,small-1.fa,NT_001493.seq},
                        t/libraries/{alu.fa,sinecutlib} (data files of the reference's own tests, not source code)
  hum1_sine.json        the 3,882 hit lines of real cross_match 0.990329 in t/seqs/masklevel/hum-1-sine-ml101.out
                        (query hum-1.fa; subject library of 2005, absent -- t/libraries/sinecutlib is a newer file
                        with the same family names), its argv, and the matrix printed in its header
  masklevel.json        survivors of SearchResultCollection::maskLevelFilter / filterContainedResults
                        (SearchResultCollection.pm:198-358, :379-470) run under perl on the two fixtures above
  matrices.json         alphabet / FREQS / integer rows of every Matrices/crossmatch/*.matrix as parsed
                        by the reference's Matrix.pm (values only; tests re-materialise matrix files
                        from it with tests/util.py:write_matrix)
"""
import json, os, re, subprocess, sys, shutil

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def parse_searchresult_t():
    src = open(f"{REF}/t/SearchResult.t").read()
    cases = []
    # each 'new SearchResult( ... );' block
    for m in re.finditer(r"new SearchResult\(\s*(.*?)\);", src, re.S):
        body = m.group(1)
        line = src[: m.start()].count("\n") + 1

        def num(key):
            mm = re.search(key + r"\s*=>\s*([-\d\.]+)", body)
            return mm.group(1) if mm else None

        def strv(key):
            mm = re.search(key + r"\s*=>\s*((?:\"[^\"]*\"\s*\.?\s*)+)", body)
            if not mm:
                return None
            return "".join(re.findall(r"\"([^\"]*)\"", mm.group(1)))

        cases.append(
            dict(
                source=f"t/SearchResult.t:{line}",
                score=int(num("score")),
                pctDiverge=num("pctDiverge"), pctDelete=num("pctDelete"), pctInsert=num("pctInsert"),
                queryStart=int(num("queryStart")), queryEnd=int(num("queryEnd")),
                subjStart=int(num("subjStart")), subjEnd=int(num("subjEnd")),
                orientation=strv("orientation"),
                matrixName=strv("matrixName"),
                queryString=strv("queryString"), subjString=strv("subjString"),
            )
        )
    # expectations asserted by the reference test itself (t/SearchResult.t line numbers)
    expect = {
        0: dict(raw=353, adjusted=253, gap_init=-30, ins_ext=-6, del_ext=-6, transitions=18, transversions=11,
                note="t/SearchResult.t:88-101; header comment :21-22 gives cross_match -gap_init -30 -ins_gap_ext -6 -del_gap_ext -5"),
        1: dict(raw=10, gap_init=-30, ins_ext=-6, del_ext=-6, note="t/SearchResult.t:150 (scoreCpGMod=1 artificial example)"),
        4: dict(adjusted=238, gap_init=-27, ins_ext=-5, del_ext=-5, note="t/SearchResult.t:316"),
        5: dict(raw=2331, adjusted=1530, gap_init=-30, ins_ext=-6, del_ext=-5,
                note="hit line t/SearchResult.t:319 (raw 2331 from -raw run); adjusted 1530 is the 'score =>' of the object"),
    }
    for i, c in enumerate(cases):
        c["expect"] = expect.get(i, {})
    return cases


def parse_cm_hits(path):
    hits = []
    for ln in open(path):
        f = ln.split()
        if len(f) >= 12 and re.match(r"^\d+$", f[0]) and re.match(r"^\d+\.\d+$", f[1]):
            star = f[-1] == "*"
            if star:
                f = f[:-1]
            if f[8] == "C":
                hits.append(dict(score=int(f[0]), div=f[1], dele=f[2], ins=f[3], qname=f[4], qstart=int(f[5]),
                                 qend=int(f[6]), qleft=int(f[7].strip("()")), strand="C", sname=f[9],
                                 sleft=int(f[10].strip("()")), send=int(f[11]), sstart=int(f[12]), star=star))
            else:
                hits.append(dict(score=int(f[0]), div=f[1], dele=f[2], ins=f[3], qname=f[4], qstart=int(f[5]),
                                 qend=int(f[6]), qleft=int(f[7].strip("()")), strand="+", sname=f[8],
                                 sstart=int(f[9]), send=int(f[10]), sleft=int(f[11].strip("()")), star=star))
    return hits


def seq1_2():
    d = f"{REF}/t/seqs/masklevel"
    return dict(
        source="t/seqs/masklevel/{seq-1.fa,seq-2.fa,seq1-2-ml101.out,seq1-2-ml98.out,seq1-2-ml97.out}",
        argv=open(f"{d}/seq1-2-ml101.out").readline().strip(),
        query_fasta=open(f"{d}/seq-1.fa").read(),
        subject_fasta=open(f"{d}/seq-2.fa").read(),
        # the 11x11 matrix is printed in the fixture header (seq1-2-ml101.out:11-23); no FREQS => uniform
        matrix_text="".join(open(f"{d}/seq1-2-ml101.out").readlines()[11:23]),
        params=dict(minmatch=10, minscore=250, gap_init=-25, ins_gap_ext=-6, del_gap_ext=-6, bandwidth=14),
        hits_ml101=parse_cm_hits(f"{d}/seq1-2-ml101.out"),
        hits_ml98=parse_cm_hits(f"{d}/seq1-2-ml98.out"),
        hits_ml97=parse_cm_hits(f"{d}/seq1-2-ml97.out"),
    )


def lambdas():
    out = {}
    perl = r'''
use lib "%s"; use Matrix;
foreach my $f (@ARGV) { my $m = Matrix->new(fileName=>$f); my @p=split(/\//,$f);
  my @rows = map { join(",", @$_) } @{$m->getMatrixValuesRef()};
  printf "%%s %%.17g %%s %%s %%s\n", $p[-1], $m->getLambda(), join("",@{$m->getAlphabetRef()}), join(",",@{$m->getMatrixFreqsRef()}), join(";",@rows); }
''' % REF
    files = sorted(
        os.path.join(f"{REF}/Matrices/crossmatch", f) for f in os.listdir(f"{REF}/Matrices/crossmatch") if f.endswith(".matrix")
    )
    res = subprocess.run(["perl", "-e", perl] + files, capture_output=True, text=True, check=True)
    for ln in res.stdout.splitlines():
        name, lam, alpha, freqs, rows = ln.split()
        out[name] = dict(lambda_=float(lam), lambda_repr=lam, alphabet=alpha, freqs=[float(x) for x in freqs.split(",")],
                         has_freqs_line="FREQS" in open(f"{REF}/Matrices/crossmatch/{name}").read().upper(),
                         rows=[[int(v) for v in r.split(",")] for r in rows.split(";")])
    return dict(source="Matrix.pm:308-364,397-464 executed by perl on Matrices/crossmatch/*.matrix", matrices=out)


def config1():
    dst = f"{OUT}/config1"
    os.makedirs(dst, exist_ok=True)
    for rel in ("t/seqs/general/hum-1.fa", "t/seqs/general/small-1.fa", "t/seqs/general/NT_001493.seq",
                "t/libraries/alu.fa", "t/libraries/sinecutlib"):
        shutil.copyfile(f"{REF}/{rel}", f"{dst}/{os.path.basename(rel)}")
        os.chmod(f"{dst}/{os.path.basename(rel)}", 0o644)


HIT_KEYS = ["score", "div", "dele", "ins", "qname", "qstart", "qend", "qleft", "strand", "sname", "sstart", "send", "sleft", "star"]


def hum1_sine():
    f = f"{REF}/t/seqs/masklevel/hum-1-sine-ml101.out"
    lines = open(f).readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("Score matrix")) + 1
    hits = parse_cm_hits(f)
    return dict(
        source="t/seqs/masklevel/hum-1-sine-ml101.out (cross_match 0.990329, 2005); query t/seqs/general/hum-1.fa",
        argv=lines[0].strip(),
        matrix_text="".join(lines[start:start + 12]),
        params=dict(minmatch=10, minscore=250, gap_init=-25, ins_gap_ext=-6, del_gap_ext=-6, bandwidth=14, masklevel=101),
        keys=HIT_KEYS,
        hits=[[h[k] for k in HIT_KEYS] for h in hits],
    )


def masklevel():
    """The reference's own Perl on the reference's own fixtures: which records survive each filter."""
    perl = r'''
use lib "%s"; use CrossmatchSearchEngine; use SearchResultCollection; use SearchResult;
my ($file, $mode, $value) = @ARGV;
my $col = CrossmatchSearchEngine::parseOutput( searchOutput => $file );
for ( my $i = 0 ; $i < $col->size() ; $i++ ) { $col->get( $i )->setId( $i ); }
if ( $mode eq "masklevel" ) { $col->maskLevelFilter( value => $value ); }
else { $col->filterContainedResults( value => $value ); }
my @ids = map { $col->get( $_ )->getId() } ( 0 .. $col->size() - 1 );
print join( ",", @ids ) . "\n";
''' % REF
    out = {}
    for name, rel in (("seq1_2", "t/seqs/masklevel/seq1-2-ml101.out"), ("hum1_sine", "t/seqs/masklevel/hum-1-sine-ml101.out")):
        hits = parse_cm_hits(f"{REF}/{rel}")
        rec = dict(source=rel, records=[[h["qname"], h["qstart"], h["qend"], h["score"]] for h in hits], masklevel={}, contained={})
        for v in (0, 10, 30, 50, 80, 90, 95, 98, 100):
            r = subprocess.run(["perl", "-e", perl, f"{REF}/{rel}", "masklevel", str(v)], capture_output=True, text=True, check=True)
            rec["masklevel"][str(v)] = [int(x) for x in r.stdout.strip().split(",") if x != ""]
        r = subprocess.run(["perl", "-e", perl, f"{REF}/{rel}", "contained", "80"], capture_output=True, text=True, check=True)
        rec["contained"]["80"] = [int(x) for x in r.stdout.strip().split(",") if x != ""]
        out[name] = rec
    return dict(source="SearchResultCollection.pm:198-358 (maskLevelFilter), :379-470 (filterContainedResults) executed by perl", sets=out)


def simplebatcher():
    """The reference's own fragmenting (FastaDB + SimpleBatcher, SimpleBatcher.pm:592-800) of seeded random sequences:
    batch ids, lengths and content digests.  repeatmasker_b200/synth.py:fragment restates that arithmetic."""
    import hashlib, random, tempfile
    perl = r'''
use lib "%s"; use FastaDB; use SimpleBatcher;
my ($f, $tmp) = @ARGV;
my $db = FastaDB->new( fileName => $f, openMode => SeqDBI::ReadOnly, maxIDLength => 50 );
my $b = SimpleBatcher->new( $db, 60000, 2000 );
for my $i (1 .. $b->getBatchCount()) {
  $b->writeBatchFile($i, $tmp);
  open(my $fh, "<", $tmp); my ($id, $seq) = ("", "");
  while (<$fh>) { chomp; if (/^>(\S+)/) { $id = $1; } else { $seq .= $_; } }
  close($fh);
  print "$id $seq\n";
}
''' % REF
    out = []
    with tempfile.TemporaryDirectory() as d:
        for n in (59999, 60001, 123457, 151062, 1000000):
            rng = random.Random(n)
            seq = "".join(rng.choice("ACGT") for _ in range(n))
            open(f"{d}/s.fa", "w").write(">s\n" + seq + "\n")
            r = subprocess.run(["perl", "-e", perl, f"{d}/s.fa", f"{d}/batch.fa"], capture_output=True, text=True, check=True)
            frags = []
            for ln in r.stdout.splitlines():
                fid, fseq = ln.split()
                frags.append([fid, len(fseq), hashlib.md5(fseq.upper().encode()).hexdigest()])
            out.append(dict(length=n, seed=n, fragments=frags))
    return dict(source="SimpleBatcher.pm:592-800 (frag 60000, overlap 2000) executed by perl on random.Random(length) sequences of ACGT", cases=out)


def main():
    config1()
    json.dump(simplebatcher(), open(f"{OUT}/simplebatcher.json", "w"))
    json.dump(hum1_sine(), open(f"{OUT}/hum1_sine.json", "w"))
    json.dump(masklevel(), open(f"{OUT}/masklevel.json", "w"))
    json.dump(parse_searchresult_t(), open(f"{OUT}/searchresult_t.json", "w"), indent=1)
    json.dump(seq1_2(), open(f"{OUT}/seq1_2.json", "w"), indent=1)
    json.dump(lambdas(), open(f"{OUT}/matrices.json", "w"))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
