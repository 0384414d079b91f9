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


def test_pending_calls_are_coalesced_into_one_pass(tmp_path):
    """What RepeatMasker's forked batch children do (RepeatMasker:1040-1180): many calls at once with the same library, matrix
    and parameters.  The worker searches their batch files as one batch (one score pass per GPU) and every caller gets exactly
    the hits of its own records; a call with another matrix is not merged but still finds the library resident."""
    import re
    import subprocess

    snames, sseqs = synth.make_library(401, n_families=40, lo=80, hi=1500)
    m = util.write_matrix("20p43g.matrix", str(tmp_path / "20p43g.matrix"))
    m2 = util.write_matrix("18p41g.matrix", str(tmp_path / "18p41g.matrix"))
    s = str(tmp_path / "lib.fa")
    synth.write_fasta(s, snames, sseqs)
    queries = []
    for k in range(8):
        g, _ = synth.make_genome(410 + k, 9000 + 1500 * k, sseqs, coverage=0.4)
        q = str(tmp_path / f"q{k}.fa")
        synth.write_fasta(q, [f"batch{k}a", f"batch{k}b"], [g[: len(g) // 2], g[len(g) // 2:]])
        queries.append(q)
    direct = [engine.search_files(q, s, m)[0] for q in queries]
    direct2 = engine.search_files(queries[0], s, m2)[0]
    assert sum(len(d) for d in direct) > 20
    got = [None] * len(queries)
    got2 = [None]
    errs = []
    sock = str(tmp_path / "rmgpu.sock")
    proc = subprocess.Popen([engine.SERVER_BIN, "--socket", sock, "--gpus", "0", "--linger-ms", "300"], stderr=subprocess.PIPE, text=True)
    try:
        for _ in range(400):
            if os.path.exists(sock) and engine.lib().rmgpu_server_ping(sock.encode()) == 0:
                break
            import time
            time.sleep(0.025)
        os.environ["RMGPU_SERVER"] = sock

        def client(k):
            try:
                got[k] = engine.search_files(queries[k], s, m)[0]
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        def other():
            try:
                got2[0] = engine.search_files(queries[0], s, m2)[0]
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=client, args=(k,)) for k in range(len(queries))] + [threading.Thread(target=other)]
        for t in th:
            t.start()
        for t in th:
            t.join()
    finally:
        os.environ.pop("RMGPU_SERVER", None)
        engine.lib().rmgpu_server_quit(sock.encode())
        log = proc.communicate(timeout=30)[1]
    assert not errs, errs
    assert got == direct
    assert got2[0] == direct2
    mm = re.search(r"served (\d+) searches in (\d+) passes", log)
    assert mm and int(mm.group(1)) == 9 and int(mm.group(2)) <= 4, log
