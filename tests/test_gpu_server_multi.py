"""The persistent worker with one rmgpu_session per GPU (batches -> per-GPU queues, no collective): several
clients search concurrently through one rmgpu_server; every answer equals the direct (in-process) search of the same
files.  Runs with however many GPUs are visible (two or more exercise the per-GPU queues)."""
import os
import threading

import numpy as np
import pytest

import util
from repeatmasker_b200 import engine, synth

pytestmark = pytest.mark.gpu


def test_concurrent_clients_over_all_gpus(tmp_path):
    import torch

    ngpu = torch.cuda.device_count()
    snames, sseqs = synth.make_library(301, n_families=40, lo=80, hi=1500)
    m = util.write_matrix("20p43g.matrix", str(tmp_path / "20p43g.matrix"))
    s = str(tmp_path / "lib.fa")
    synth.write_fasta(s, snames, sseqs)
    queries = []
    for k in range(6):
        g, _ = synth.make_genome(310 + k, 20000 + 3000 * k, sseqs, coverage=0.4)
        q = str(tmp_path / f"q{k}.fa")
        synth.write_fasta(q, [f"frag{k}"], [g])
        queries.append(q)
    strip = lambda hits: [{k: v for k, v in h.items()} for h in hits]
    direct = [strip(engine.search_files(q, s, m)[0]) for q in queries]
    assert all(len(d) > 0 for d in direct)
    got = [None] * len(queries)
    errs = []
    with engine.Server(str(tmp_path / "rmgpu.sock"), gpus=list(range(ngpu))) as srv:
        def client(k):
            try:
                for _ in range(3):  # repeated calls: library stays resident on whichever GPU serves them
                    got[k] = strip(engine.search_files(queries[k], s, m)[0])
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=client, args=(k,)) for k in range(len(queries))]
        for t in th:
            t.start()
        for t in th:
            t.join()
    assert not errs, errs
    assert got == direct
