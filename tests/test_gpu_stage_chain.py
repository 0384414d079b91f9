"""Stage chain on a RESIDENT batch: RepeatMasker's -species runs search one batch against a chain of
libraries and, between two stages, mask what the stage found (`$seqDB->setSubstr( $name, $begin - 1, $len, 'X' x $len )`,
RepeatMasker:6036-6090; the chain itself is RepeatMasker:4115-4654).  The engine's half of that: the batch stays in HBM,
rmgpu_session_mask_queries() masks the ranges on the device, the next stage's library is swapped in, and the next
search sees the masked batch.  Checked hit for hit (alignment rows included: they carry the X's) against the oracle run on
a batch file that was masked on the host the way the reference masks it."""
import os

import numpy as np
import pytest

import util
from repeatmasker_b200 import engine, synth
from test_gpu_parity import assert_same_hits

pytestmark = pytest.mark.gpu


def host_mask(seq, ranges):
    b = bytearray(seq.encode())
    for s, e in ranges:  # 1-based inclusive, as the reference's setSubstr( $name, $begin - 1, $len, ... )
        b[s - 1:e] = b"X" * (e - s + 1)
    return b.decode()


def test_two_stage_chain_with_device_masking(tmp_path):
    rng = np.random.default_rng(4115)
    # stage-1 library: short elements; stage-2 library: long elements, some of whose genomic copies carry a stage-1
    # element inserted in the middle (the case the masking exists for) or overlap one at an end
    n1, s1 = synth.make_library(31, n_families=12, lo=150, hi=400)
    n2, s2 = synth.make_library(32, n_families=10, lo=900, hi=2600)
    n1 = [f"st1_{x}" for x in n1]
    n2 = [f"st2_{x}" for x in n2]
    parts = []
    for i in range(14):
        parts.append(synth.random_dna(rng, int(rng.integers(300, 1500))).tobytes().decode())
        long_copy = synth.mutate(rng, s2[i % len(s2)], 0.08 + 0.01 * (i % 5), 0.01)
        short_copy = synth.mutate(rng, s1[i % len(s1)], 0.06, 0.005)
        if i % 3 == 0:  # nested: the short element sits inside the long one
            cut = len(long_copy) // 2
            long_copy = long_copy[:cut] + short_copy + long_copy[cut:]
        elif i % 3 == 1:  # abutting
            long_copy = long_copy + short_copy
        else:  # apart
            long_copy = long_copy + synth.random_dna(rng, 200).tobytes().decode() + short_copy
        parts.append(synth.revcomp(long_copy) if i & 1 else long_copy)
    genome = "".join(parts) + synth.random_dna(rng, 500).tobytes().decode()
    qn, qs = synth.fragment(genome, "chrC", frag=12000, overlap=500)
    assert len(qn) >= 3
    m = util.write_matrix("20p43g.matrix", str(tmp_path / "m.matrix"))
    q0, l1, l2 = str(tmp_path / "q0.fa"), str(tmp_path / "l1.fa"), str(tmp_path / "l2.fa")
    synth.write_fasta(q0, qn, qs)
    synth.write_fasta(l1, n1, s1)
    synth.write_fasta(l2, n2, s2)

    s = engine.Session()
    s.set_matrix(m)
    s.set_queries(qs, qn)
    p = engine.make_params()
    # ---- stage 1
    s.set_library(s1, n1)
    got1, _ = s.search(p)
    want1, _ = util.run_oracle(q0, l1, m)
    assert len(want1) >= 10
    assert_same_hits(got1, want1)
    # ---- mask on the device what stage 1 found (overlapping ranges included), and on the host for the oracle
    ranges = [(h["query_index"], h["qstart"], h["qend"]) for h in got1]
    s.mask_queries(ranges)
    per_record = {}
    for qi, b, e in ranges:
        per_record.setdefault(qi, []).append((b, e))
    masked = [host_mask(sq, per_record.get(i, [])) for i, sq in enumerate(qs)]
    assert sum(sq.count("X") for sq in masked) > 1500
    q1 = str(tmp_path / "q1.fa")
    synth.write_fasta(q1, qn, masked)
    # ---- stage 2 on the resident, masked batch
    s.set_library(s2, n2)
    got2, st2 = s.search(p)
    want2, _ = util.run_oracle(q1, l2, m)
    assert len(want2) >= 10
    assert_same_hits(got2, want2)
    assert any("X" in h["qaln"] for h in got2)  # alignments that run through a masked insertion
    # the unmasked batch gives different stage-2 hits: the masking is what made them equal
    unmasked2, _ = util.run_oracle(q0, l2, m)
    assert [(h["score"], h["qstart"], h["qend"]) for h in unmasked2] != [(h["score"], h["qstart"], h["qend"]) for h in want2]
    # ---- a third stage: mask again (ranges that overlap earlier ones), back to the first library: nothing of it is left
    s.mask_queries([(h["query_index"], h["qstart"], h["qend"]) for h in got2])
    s.set_library(s1, n1)
    got3, _ = s.search(p)
    for qi, b, e in [(h["query_index"], h["qstart"], h["qend"]) for h in got2]:
        masked[qi] = host_mask(masked[qi], [(b, e)])
    q2 = str(tmp_path / "q2.fa")
    synth.write_fasta(q2, qn, masked)
    want3, _ = util.run_oracle(q2, l1, m)
    assert_same_hits(got3, want3)
    s.close()


def test_mask_queries_argument_checks():
    s = engine.Session()
    with pytest.raises(engine.RmgpuError):
        s.mask_queries([(0, 1, 10)])  # no batch resident
    s.set_queries(["ACGTACGTACGT"], ["q"])
    with pytest.raises(engine.RmgpuError):
        s.mask_queries([(1, 1, 4)])  # record index out of range
    s.mask_queries([])  # nothing to do
    s.mask_queries([(0, -5, 3), (0, 11, 99)])  # clamped to the record
    s.close()
