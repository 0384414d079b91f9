"""CPU-side checks: librmgpu.so builds/loads, exports every symbol include/rmgpu.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls."""
import ctypes as C
import os
import re

import pytest

import util
from repeatmasker_b200 import build as rm_build
from repeatmasker_b200 import engine


def test_library_builds_and_exports_header_symbols():
    rm_build.build()
    L = engine.lib()
    hdr = open(os.path.join(util.ROOT, "include", "rmgpu.h")).read()
    declared = set(re.findall(r"\b(rmgpu_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/rmgpu.h but not exported"
    assert set(engine.EXPORTS) <= declared
    assert L.rmgpu_abi_version() == 4
    assert b"rmgpu version" in L.rmgpu_version()


def test_struct_layouts_match_header(tmp_path):
    """include/rmgpu.h must be plain C (gcc -std=c99) and the ctypes mirrors must have the same layout."""
    import subprocess

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rmgpu.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(rmgpu_params),sizeof(rmgpu_hit),sizeof(rmgpu_stats),sizeof(rmgpu_result),"
                   "offsetof(rmgpu_hit,qaln),offsetof(rmgpu_result,stats));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(util.ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(engine.Params), C.sizeof(engine.Hit), C.sizeof(engine.Stats),
                                     C.sizeof(engine.Result), engine.Hit.qaln.offset, engine.Result.stats.offset]


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.RmgpuError, match="no CUDA device"):
        engine.Session()


def test_python_mirror_croaks_like_reference(tmp_path):
    e = engine.GpuSearchEngine()
    e.setQuery(str(tmp_path / "missing.fa"))
    with pytest.raises(engine.RmgpuError, match="does not exist"):
        e.getParameters()
    assert e.setMinScore(225) is None and e.getMinScore() == 225


def test_persistent_worker_protocol_without_gpu(tmp_path):
    """rmgpu_server comes up, answers ping, relays a search request and its error text over the wire, and quits.
    (On the CPU box the search itself fails loudly -- no CUDA device -- which is exactly what must travel back.)"""
    import torch

    from repeatmasker_b200 import synth

    rm_build.build()
    q, s_, m = str(tmp_path / "q.fa"), str(tmp_path / "s.fa"), str(tmp_path / "m.matrix")
    synth.write_fasta(q, ["q"], ["ACGTACGTAGCTAGCTAGCTAGCATCGATCGATCG"])
    synth.write_fasta(s_, ["s"], ["ACGTACGTAGCTAGCTAGCTAGCATCGATCGATCG"])
    util.write_matrix("20p43g.matrix", m)
    sock = str(tmp_path / "rmgpu.sock")
    with engine.Server(sock) as srv:
        assert engine.lib().rmgpu_server_ping(sock.encode()) == 0
        assert os.environ["RMGPU_SERVER"] == sock
        if torch.cuda.is_available():
            hits, st = engine.search_files(q, s_, m, minscore=100)
            assert len(hits) >= 1 and max(h["qend"] - h["qstart"] for h in hits) == 34
            with pytest.raises(engine.RmgpuError, match="does not exist"):
                engine.search_files(q, str(tmp_path / "missing.fa"), m)
        else:
            with pytest.raises(engine.RmgpuError, match="no CUDA device"):
                engine.search_files(q, s_, m, minscore=100)
    assert "RMGPU_SERVER" not in os.environ
    assert srv.proc.returncode == 0
    with pytest.raises(engine.RmgpuError, match="cannot reach rmgpu_server"):
        os.environ["RMGPU_SERVER"] = sock
        try:
            engine.search_files(q, s_, m)
        finally:
            del os.environ["RMGPU_SERVER"]
