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


def run_rm(workdir, engine, extra=(), env=None):
    cmd = (f"source {HARNESS}/env.sh && cd {workdir} && perl $RM_HOME/RepeatMasker -engine {engine} -lib lib.fa -gc 43 "
           f"-pa 4 {' '.join(extra)} genome.fa")
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run(["bash", "-c", cmd], capture_output=True, text=True, env=e)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


GENOME_BP = 140_000  # three SimpleBatcher fragments: frag-N ids, 2 kbp overlaps and the driver's re-basing flow


@pytest.mark.skipif(not (os.path.exists(os.path.join(HARNESS, "env.sh")) and shutil.which("perl")),
                    reason="baseline/_ref harness not present (tools/make_rm_harness.sh)")
@pytest.mark.parametrize("extra", [(), ("-s",)])
def test_repeatmasker_driver_outputs_identical(tmp_path, extra):
    subprocess.run(["make", "-C", os.path.join(util.ROOT, "repeatmasker_b200", "perl")], check=True, capture_output=True)
    snames, sseqs = synth.make_library(1, n_families=24, lo=80, hi=1500)
    genome, _ = synth.make_genome(2, GENOME_BP, sseqs, coverage=0.45)
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
    """The three ways the plug-in reaches the GPU give the same files: the worker the plug-in starts itself (default), an
    rmgpu_server named by RMGPU_SERVER, and RMGPU_NO_SERVER=1 (a CUDA context per search() call in the forked child)."""
    from repeatmasker_b200 import engine

    subprocess.run(["make", "-C", os.path.join(util.ROOT, "repeatmasker_b200", "perl")], check=True, capture_output=True)
    snames, sseqs = synth.make_library(1, n_families=24, lo=80, hi=1500)
    genome, _ = synth.make_genome(2, GENOME_BP, sseqs, coverage=0.45)
    outs = {}
    for mode in ("auto", "server", "inprocess"):
        d = tmp_path / mode
        d.mkdir()
        synth.write_fasta(str(d / "lib.fa"), snames, sseqs)
        synth.write_fasta(str(d / "genome.fa"), ["chrT"], [genome], width=60)
        if mode == "server":
            with engine.Server(str(tmp_path / "rmgpu.sock")) as srv:
                run_rm(str(d), "gpu")
            assert "served" in srv.log and "served 0 searches" not in srv.log
        elif mode == "inprocess":
            run_rm(str(d), "gpu", env={"RMGPU_NO_SERVER": "1"})
        else:
            run_rm(str(d), "gpu")
        outs[mode] = {ext: open(d / f"genome.fa.{ext}").read() for ext in ("out", "masked", "cat")}
    assert outs["auto"] == outs["server"] == outs["inprocess"]
    assert outs["auto"]["out"].count("\n") > 8


@pytest.mark.skipif(not (os.path.exists(os.path.join(HARNESS, "env.sh")) and shutil.which("perl")),
                    reason="baseline/_ref harness not present (tools/make_rm_harness.sh)")
def test_driver_gc_bins_and_low_complexity_without_gc_flag(tmp_path):
    """BASELINE configuration 4 through the driver, WITHOUT -gc: the genome is four blocks at 37/41/45/51 % GC with planted
    (CA)n / (A)n / (GAA)n tracts and AT-rich runs, so RepeatMasker::chooseMatrices (RepeatMasker:4784-4821) picks a different
    GC-bin matrix per batch (the plug-in's worker keeps the library resident across those matrix changes) and the simple-repeat
    consensi produce dense, overlapping, low-complexity hits whose scores hinge on the complexity adjustment.  Both engines
    must write the same files."""
    from test_gpu_fullsize import _gc_block_genome

    snames, sseqs = synth.make_library(1001, n_families=120)
    snames = snames + ["(CA)n#Simple_repeat", "(A)n#Simple_repeat", "(GAA)n#Simple_repeat", "AT_rich#Low_complexity"]
    sseqs = sseqs + ["CA" * 90, "A" * 150, "GAA" * 60, "AT" * 50 + "A" * 20 + "TA" * 40]
    genome = _gc_block_genome(1003, sseqs[:120], 50000)
    outs, logs = {}, {}
    for engine in ("crossmatch", "gpu"):
        d = tmp_path / engine
        d.mkdir()
        synth.write_fasta(str(d / "lib.fa"), snames, sseqs)
        synth.write_fasta(str(d / "genome.fa"), ["gcblocks"], [genome], width=60)
        logs[engine] = util.run_repeatmasker(str(d), engine, ["-frag", "50000"], pa=4)
        outs[engine] = util.driver_outputs(str(d))
    a, b = outs["crossmatch"], outs["gpu"]
    assert a["out"].count("\n") > 50
    assert a["out"] == b["out"]
    assert a["masked"] == b["masked"]
    assert util.strip_engine_trailer(a["cat"]) == util.strip_engine_trailer(b["cat"])
    # the batches really ran with different matrices (the .cat records name the matrix of every alignment)
    import re

    matrices = set(re.findall(r"Matrix = (\S+)", b["cat"]))
    assert len(matrices) >= 2, matrices


@pytest.mark.skipif(not (os.path.exists(os.path.join(HARNESS, "env.sh")) and shutil.which("perl")),
                    reason="baseline/_ref harness not present (tools/make_rm_harness.sh)")
def test_driver_configuration2_full_size_both_engines(tmp_path):
    """BASELINE configuration 2 at FULL size through the unmodified driver with both engines (1 Mbp x 500 families, 1.2e12
    cells: the CPU arm is the oracle on every host core, about half a minute on 16 cores): .out, .masked and .tbl
    byte-identical, .cat identical apart from the engine trailer -- the north star's acceptance criterion at a BASELINE size."""
    snames, sseqs = synth.make_library(1001, n_families=500)
    genome, _ = synth.make_genome(1001, 1_000_000, sseqs, coverage=0.45)
    outs = {}
    for engine in ("crossmatch", "gpu"):
        d = tmp_path / engine
        d.mkdir()
        synth.write_fasta(str(d / "lib.fa"), snames, sseqs)
        synth.write_fasta(str(d / "genome.fa"), ["synth0"], [genome], width=60)
        util.run_repeatmasker(str(d), engine, ["-gc", "41"], pa=os.cpu_count() or 4)
        outs[engine] = util.driver_outputs(str(d))
    a, b = outs["crossmatch"], outs["gpu"]
    assert a["out"].count("\n") > 800
    assert a["out"] == b["out"]
    assert a["masked"] == b["masked"]
    assert util.strip_engine_trailer(a["cat"]) == util.strip_engine_trailer(b["cat"])
    assert util.strip_engine_trailer(a["tbl"]) == util.strip_engine_trailer(b["tbl"])
