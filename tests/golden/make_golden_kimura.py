"""Golden vectors for the Kimura divergence (CpG-modified) that CrossmatchSearchEngine.pm:509-511 stores per hit.

Run in the build container (needs perl and /root/reference): random gapped alignment pairs in the orientation of
rmgpu_hit::qaln/saln (+ a strand flag) go through the reference's own Perl -- the strand-C flip of
CrossmatchSearchEngine.pm:824-837, then SearchResult::calcKimuraDivergence(divCpGMod => 1) (SearchResult.pm:1607-1683)
and sprintf("%4.2f") -- and the answers are committed as tests/golden/kimura.json.
usage: python tests/golden/make_golden_kimura.py
"""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

PERL = r'''
use strict; use JSON::PP; use SearchResult;
local $/; my $cases = decode_json(<STDIN>); my @out;
foreach my $c (@$cases) {
  my ($q, $s) = ($c->{qaln}, $c->{saln});
  if ($c->{strand}) { $q = reverse $q; $q =~ tr/ACGTYRMKHBVD/TGCARYKMDVBH/; $s = reverse $s; $s =~ tr/ACGTYRMKHBVD/TGCARYKMDVBH/; }
  my $r = SearchResult->new(queryName => "q", subjName => "s", queryStart => 1, queryEnd => 1, subjStart => 1, subjEnd => 1,
                            queryString => $q, subjString => $s);
  my ($div, $ti, $tv, $wcb, $cpg) = $r->calcKimuraDivergence(divCpGMod => 1);
  push @out, { text => sprintf("%4.2f", $div), transitions => $ti + 0, transversions => $tv + 0, wcb => $wcb + 0, cpg => $cpg + 0 };
}
print encode_json(\@out);
'''


def make_cases(n=240, seed=4711):
    rng = np.random.default_rng(seed)
    cases = []
    for k in range(n):
        length = int(rng.integers(1, 600))
        cpg_rich = k % 3 == 0
        s, q = [], []
        for _ in range(length):
            if cpg_rich and rng.random() < 0.25:
                pair = "CG"
            else:
                pair = "ACGT"[rng.integers(4)]
            for b in pair:
                r = rng.random()
                div = 0.02 + 0.3 * (k % 7) / 6
                if r < div * 0.6:        # transition
                    qb = {"A": "G", "G": "A", "C": "T", "T": "C"}[b]
                elif r < div:            # transversion
                    qb = {"A": "CT", "G": "CT", "C": "AG", "T": "AG"}[b][rng.integers(2)]
                else:
                    qb = b
                u = rng.random()
                if u < 0.02:
                    s.append(b); q.append("-")          # gap in the query row
                elif u < 0.04:
                    s.append("-"); q.append(qb)         # gap in the subject row
                elif u < 0.05:
                    s.append("NRYKMSWBDHVX"[rng.integers(12)]); q.append(qb)
                elif u < 0.06:
                    s.append(b); q.append("NRYKMSWBDHVX"[rng.integers(12)])
                else:
                    s.append(b); q.append(qb)
        cases.append({"qaln": "".join(q), "saln": "".join(s), "strand": int(k % 2)})
    cases.append({"qaln": "A", "saln": "-", "strand": 0})                 # no well-characterised pair: 100.00
    cases.append({"qaln": "ACGT" * 5, "saln": "CATG" * 5, "strand": 1})   # saturated: log operand <= 0
    cases.append({"qaln": "TGTGTG", "saln": "CGCGCG", "strand": 0})       # runs of CpG with single transitions
    cases.append({"qaln": "TATATA", "saln": "CGCGCG", "strand": 0})       # two transitions per CpG
    return cases


def main():
    cases = make_cases()
    with tempfile.NamedTemporaryFile("w", suffix=".pl", delete=False) as f:
        f.write(PERL)
    out = subprocess.run(["perl", f"-I{REF}", f.name], input=json.dumps(cases), capture_output=True, text=True, check=True).stdout
    os.unlink(f.name)
    exp = json.loads(out)
    for c, e in zip(cases, exp):
        c["expect"] = e
    with open(os.path.join(HERE, "kimura.json"), "w") as fh:
        json.dump({"source": "SearchResult::calcKimuraDivergence(divCpGMod => 1) of the reference, run by make_golden_kimura.py",
                   "cases": cases}, fh, separators=(",", ":"))
    print(len(cases), "cases ->", os.path.join(HERE, "kimura.json"))


if __name__ == "__main__":
    main()
