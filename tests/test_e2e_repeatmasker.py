"""North-star acceptance test: the reference driver itself, `RepeatMasker -engine gpu` (librmgpu through
GpuSearchEngine.pm / XS) against `RepeatMasker -engine crossmatch` (the CPU oracle installed as cross_match), same
inputs: .out and .masked byte-identical, .cat identical apart from the engine trailer line.

Needs baseline/_ref (built by tools/make_rm_harness.sh in the build container; git-ignored, travels with gpurun).
Nothing here reads /root/reference."""
import os
import shutil
import subprocess

import pytest

import util
from repeatmasker_b200 import synth

pytestmark = pytest.mark.gpu

HARNESS = os.path.join(util.ROOT, "baseline", "_ref")


def run_rm(workdir, engine, extra=()):
    cmd = (f"source {HARNESS}/env.sh && cd {workdir} && perl $RM_HOME/RepeatMasker -engine {engine} -lib lib.fa -gc 43 "
           f"-pa 2 {' '.join(extra)} genome.fa")
    res = subprocess.run(["bash", "-c", cmd], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


@pytest.mark.skipif(not (os.path.exists(os.path.join(HARNESS, "env.sh")) and shutil.which("perl")),
                    reason="baseline/_ref harness not present (tools/make_rm_harness.sh)")
@pytest.mark.parametrize("extra", [(), ("-s",)])
def test_repeatmasker_driver_outputs_identical(tmp_path, extra):
    subprocess.run(["make", "-C", os.path.join(util.ROOT, "repeatmasker_b200", "perl")], check=True, capture_output=True)
    snames, sseqs = synth.make_library(1, n_families=24, lo=80, hi=1500)
    genome, _ = synth.make_genome(2, 30000, sseqs, coverage=0.45)
    outs = {}
    for engine in ("crossmatch", "gpu"):
        d = tmp_path / engine
        d.mkdir()
        synth.write_fasta(str(d / "lib.fa"), snames, sseqs)
        synth.write_fasta(str(d / "genome.fa"), ["chrT"], [genome], width=60)
        log = run_rm(str(d), engine, extra)
        if engine == "gpu":
            assert "Search Engine: rmgpu" in log
        outs[engine] = {ext: open(d / f"genome.fa.{ext}").read() for ext in ("out", "masked", "cat", "tbl")}
    a, b = outs["crossmatch"], outs["gpu"]
    assert a["out"].count("\n") > 8
    assert a["out"] == b["out"]
    assert a["masked"] == b["masked"]
    strip = lambda t: "\n".join(l for l in t.splitlines() if not l.startswith("run with "))
    assert strip(a["cat"]) == strip(b["cat"])
    assert strip(a["tbl"]) == strip(b["tbl"])


@pytest.mark.skipif(not (os.path.exists(os.path.join(HARNESS, "env.sh")) and shutil.which("perl")),
                    reason="baseline/_ref harness not present (tools/make_rm_harness.sh)")
def test_repeatmasker_driver_through_persistent_worker(tmp_path):
    """The same driver run with RMGPU_SERVER set: the forked batch children are thin clients of one rmgpu_server
    (no CUDA context per child, library packed once); outputs must not change."""
    from repeatmasker_b200 import engine

    subprocess.run(["make", "-C", os.path.join(util.ROOT, "repeatmasker_b200", "perl")], check=True, capture_output=True)
    snames, sseqs = synth.make_library(1, n_families=24, lo=80, hi=1500)
    genome, _ = synth.make_genome(2, 30000, sseqs, coverage=0.45)
    outs = {}
    for mode in ("direct", "server"):
        d = tmp_path / mode
        d.mkdir()
        synth.write_fasta(str(d / "lib.fa"), snames, sseqs)
        synth.write_fasta(str(d / "genome.fa"), ["chrT"], [genome], width=60)
        if mode == "server":
            with engine.Server(str(tmp_path / "rmgpu.sock")) as srv:
                run_rm(str(d), "gpu")
            assert "served" in srv.log and "served 0 searches" not in srv.log
        else:
            run_rm(str(d), "gpu")
        outs[mode] = {ext: open(d / f"genome.fa.{ext}").read() for ext in ("out", "masked", "cat")}
    assert outs["direct"] == outs["server"]
    assert outs["direct"]["out"].count("\n") > 8
