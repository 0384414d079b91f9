"""N > 1 host logic on the CPU: two gloo ranks each search their round-robin share of the fragments (with the CPU
oracle standing in for the GPU engine: it is the checker), gather, and must reproduce the single-process result in
fragment order."""
import json
import os
import socket
import subprocess
import sys

import util

WORKER = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
import util
from repeatmasker_b200 import sharding, synth
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d = sys.argv[2]
snames, sseqs = synth.make_library(41, n_families=12, lo=80, hi=500)
genome, _ = synth.make_genome(42, 9000, sseqs, coverage=0.5)
names, frags = synth.fragment(genome, "chrG", frag=2500, overlap=200)
mine = sharding.my_fragments(len(frags), rank, world)
q = os.path.join(d, f"q{rank}.fa"); s = os.path.join(d, f"s{rank}.fa"); m = os.path.join(d, f"m{rank}.matrix")
synth.write_fasta(q, [names[i] for i in mine], [frags[i] for i in mine]); synth.write_fasta(s, snames, sseqs)
util.write_matrix("20p43g.matrix", m)
hits, _ = util.run_oracle(q, s, m)
local = sharding.split_hits_by_fragment(hits, [names[i] for i in mine])
allhits = sharding.gather_hits(local, mine, len(frags))
if rank == 0:
    json.dump(allhits, open(os.path.join(d, "gathered.json"), "w"))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_sharding_reproduces_single_process(tmp_path):
    from repeatmasker_b200 import sharding, synth

    (tmp_path / "worker.py").write_text(WORKER)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(tmp_path / "worker.py"),
                          util.ROOT, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    gathered = json.load(open(tmp_path / "gathered.json"))
    snames, sseqs = synth.make_library(41, n_families=12, lo=80, hi=500)
    genome, _ = synth.make_genome(42, 9000, sseqs, coverage=0.5)
    names, frags = synth.fragment(genome, "chrG", frag=2500, overlap=200)
    q, s, m = str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "m.matrix")
    synth.write_fasta(q, names, frags)
    synth.write_fasta(s, snames, sseqs)
    util.write_matrix("20p43g.matrix", m)
    want = sharding.split_hits_by_fragment(util.run_oracle(q, s, m)[0], names)
    assert len(gathered) == len(names) >= 4
    assert gathered == want
    assert sum(len(x) for x in want) > 5


def test_round_robin_assignment_is_a_partition():
    from repeatmasker_b200 import sharding

    for n in (0, 1, 7, 18, 53):
        for w in (1, 2, 4, 8):
            got = sorted(i for r in range(w) for i in sharding.my_fragments(n, r, w))
            assert got == list(range(n))
