#!/usr/bin/env python3
"""Concordance of this repository's search definition with real cross_match on the largest
run the reference ships: t/seqs/masklevel/hum-1-sine-ml101.out (cross_match 0.990329, 3,882 hits; fixture copy
tests/golden/hum1_sine.json) against the oracle -- and, with --gpu, librmgpu -- on the same query
(tests/golden/config1/hum-1.fa), the argv of line 1 of the fixture and the matrix printed in its header.

The subject library of that 2005 run is not in the reference tree.  tests/golden/config1/sinecutlib is the newer
77-sequence file of the same name: it still holds 7 of the run's 22 families (AluJb, AluSx, AluY, FLAM_C, FRAM, BC200,
7SLRNA); 11 more Alu subfamilies are in t/libraries/alu.fa (--with-alu adds those whose name sinecutlib lacks); FAM,
FLAM_A, AluSg1 and SVA (today SVA_A..F) have no sequence of that name in the reference tree.  The comparison is
therefore by family NAME and query interval, and a consensus that changed since 2005 shifts scores.

  python tools/concordance_hum1_sine.py [--gpu] [--out profiles/r2_concordance_hum1_sine.txt]
"""
import argparse
import collections
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402


def key_exact(h):
    return (h["qname"], h["qstart"], h["qend"], h["strand"], h["sname"], h["sstart"], h["send"], h["score"], h["div"], h["dele"], h["ins"])


def overlap(a, b):
    return max(0, min(a["qend"], b["qend"]) - max(a["qstart"], b["qstart"]) + 1)


def classify(ref, got):
    """Every cross_match hit is put in exactly one class by its best partner among our hits of the same query, subject
    and strand."""
    by = collections.defaultdict(list)
    for g in got:
        by[(g["qname"], g["sname"], g["strand"])].append(g)
    exact = set(key_exact(g) for g in got)
    cls = collections.Counter()
    per_family = collections.defaultdict(collections.Counter)
    used = set()
    for r in ref:
        c = None
        if key_exact(r) in exact:
            c = "identical line (score, %, every coordinate)"
        else:
            best, bo = None, 0
            for g in by.get((r["qname"], r["sname"], r["strand"]), []):
                o = overlap(r, g)
                if o > bo:
                    best, bo = g, o
            if best is None:
                c = "no hit of ours overlaps it (same query, subject, strand)"
            else:
                used.add(id(best))
                rl, gl = r["qend"] - r["qstart"] + 1, best["qend"] - best["qstart"] + 1
                if (r["qstart"], r["qend"], r["sstart"], r["send"]) == (best["qstart"], best["qend"], best["sstart"], best["send"]):
                    c = "same coordinates, different score or percentages"
                elif bo >= 0.9 * rl and bo >= 0.9 * gl:
                    c = "same locus, extents within 10 %"
                elif bo >= 0.9 * rl:
                    c = "contained in a longer alignment of ours (rule 3 merges what cross_match splits)"
                elif bo >= 0.9 * gl:
                    c = "ours is a part of a longer cross_match alignment"
                else:
                    c = "partial overlap"
        cls[c] += 1
        per_family[r["sname"]][c] += 1
    return cls, per_family


def score_deltas(ref, got):
    pos = {(g["qname"], g["qstart"], g["qend"], g["strand"], g["sname"], g["sstart"], g["send"]): g["score"] for g in got}
    out = collections.Counter()
    for r in ref:
        k = (r["qname"], r["qstart"], r["qend"], r["strand"], r["sname"], r["sstart"], r["send"])
        if k in pos:
            out[pos[k] - r["score"]] += 1
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--with-alu", action="store_true", help="add the families of t/libraries/alu.fa that sinecutlib lacks")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    fx = json.load(open(os.path.join(util.GOLD, "hum1_sine.json")))
    ref = [dict(zip(fx["keys"], h)) for h in fx["hits"]]
    P = fx["params"]
    lines = []
    with tempfile.TemporaryDirectory() as d:
        m = os.path.join(d, "comparison.matrix")
        open(m, "w").write(fx["matrix_text"])
        q = os.path.join(util.GOLD, "config1", "hum-1.fa")
        s = os.path.join(util.GOLD, "config1", "sinecutlib")
        if args.with_alu:
            s = util.union_library(s, os.path.join(util.GOLD, "config1", "alu.fa"), os.path.join(d, "union.fa"))
        kw = dict(gap_init=P["gap_init"], ins_gap_ext=P["ins_gap_ext"], del_gap_ext=P["del_gap_ext"], minscore=P["minscore"],
                  masklevel=P["masklevel"], minmatch=P["minmatch"], bandwidth=P["bandwidth"])
        got, _ = util.run_oracle(q, s, m, **kw)
        arms = [("oracle (restated crossmatch, exhaustive Smith-Waterman-Gotoh)", got)]
        if args.gpu:
            from repeatmasker_b200 import engine
            ghits, st = engine.search_files(q, s, m, device=0, **kw)
            for g in ghits:
                g["strand"] = "C" if g["strand"] in (1, "C") else "+"
            arms.append(("librmgpu (B200)", ghits))
    lib_names = set(h["sname"] for h in arms[-1][1])
    ref_names = collections.Counter(h["sname"] for h in ref)
    lines.append(f"cross_match 0.990329 fixture: {len(ref)} hits over {len(ref_names)} subject names; argv: {fx['argv']}")
    for name, hits in arms:
        cls, fam = classify(ref, hits)
        lines.append("")
        lines.append(f"== {name}: {len(hits)} hits")
        for c, n in cls.most_common():
            lines.append(f"  {n:5d}  {100.0 * n / len(ref):5.1f} %  {c}")
        deltas = score_deltas(ref, hits)
        lines.append("  same coordinates, our score minus cross_match's: " + ", ".join(f"{k:+d}: {v}" for k, v in sorted(deltas.items(), key=lambda kv: -kv[1])[:8]))
        missing_names = sorted(n for n in ref_names if n not in lib_names)
        lines.append(f"  subject names of the 2005 run with no hit of ours at all: {len(missing_names)} "
                     f"({sum(ref_names[n] for n in missing_names)} fixture hits): {' '.join(missing_names[:12])}")
        lines.append("  per family (fixture hits: identical / same coordinates / same locus / merged / part / partial / none):")
        order = ["identical line (score, %, every coordinate)", "same coordinates, different score or percentages", "same locus, extents within 10 %",
                 "contained in a longer alignment of ours (rule 3 merges what cross_match splits)", "ours is a part of a longer cross_match alignment",
                 "partial overlap", "no hit of ours overlaps it (same query, subject, strand)"]
        for sname, n in ref_names.most_common():
            lines.append(f"    {sname:14s} {n:5d}: " + " / ".join(str(fam[sname][c]) for c in order))
    text = "\n".join(lines) + "\n"
    sys.stdout.write(text)
    if args.out:
        open(args.out, "w").write(text)


if __name__ == "__main__":
    main()
