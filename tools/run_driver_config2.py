"""The real RepeatMasker driver on BASELINE configuration 2 (1 Mbp synthetic genome x 500-family library) with
`-engine gpu` through a persistent rmgpu_server (other sizes: --families 1500 --seed 1002 --mbp 10 --extra=-s is a
slice of configuration 3): wall time of the whole run (Perl driver, batching, the search calls of
every batch, ProcessRepeats) next to the time the GPU worker spent searching.  Needs baseline/_ref
(tools/make_rm_harness.sh).  usage: python tools/run_driver_config2.py [--mbp 1.0] [--pa 4]"""
import argparse, os, subprocess, sys, tempfile, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from repeatmasker_b200 import engine, synth

ap = argparse.ArgumentParser(); ap.add_argument("--mbp", type=float, default=1.0); ap.add_argument("--pa", type=int, default=4)
ap.add_argument("--families", type=int, default=500); ap.add_argument("--seed", type=int, default=1001)
ap.add_argument("--extra", default="", help="extra driver flags, e.g. -s")
ap.add_argument("--server-gpus", default="0", help="GPU ordinals of the rmgpu_server workers, e.g. 0,1 (0,0 = two sessions on one GPU)")
ap.add_argument("--also-crossmatch", type=int, default=0, metavar="PA",
                help="also run `-engine crossmatch -pa PA` (the CPU oracle as cross_match) on the same input and compare the outputs")
a = ap.parse_args()
H = os.path.join(ROOT, "baseline", "_ref")
assert os.path.exists(os.path.join(H, "env.sh")), "run tools/make_rm_harness.sh in the build container first"
subprocess.run(["make", "-C", os.path.join(ROOT, "repeatmasker_b200", "perl")], check=True, capture_output=True)
d = tempfile.mkdtemp()
snames, sseqs = synth.make_library(a.seed, n_families=a.families)
parts, left, k = [], int(a.mbp * 1e6), 0
while left > 0:  # generated 1 Mbp at a time, one sequence in the end
    g, _ = synth.make_genome(a.seed + 31 * k, min(left, 1_000_000), sseqs, coverage=0.45)
    parts.append(g); left -= 1_000_000; k += 1
genome = "".join(parts)
synth.write_fasta(os.path.join(d, "lib.fa"), snames, sseqs)
synth.write_fasta(os.path.join(d, "genome.fa"), ["synth0"], [genome], width=60)
trace = os.path.join(d, "passes.txt")
with engine.Server(os.path.join(d, "rmgpu.sock"), gpus=[int(x) for x in a.server_gpus.split(",")], trace=trace) as srv:
    t_unix0 = time.time()
    cmd = f"source {H}/env.sh && cd {d} && perl $RM_HOME/RepeatMasker -engine gpu -lib lib.fa -gc 41 -pa {a.pa} {a.extra} genome.fa"
    t0 = time.perf_counter()
    res = subprocess.run(["bash", "-c", cmd], capture_output=True, text=True)
    wall = time.perf_counter() - t0
print(res.stdout[-1500:])
if res.returncode:
    print(res.stderr[-3000:]); sys.exit(1)
out = open(os.path.join(d, "genome.fa.out")).read().splitlines()
masked = open(os.path.join(d, "genome.fa.masked")).read()
nN = sum(l.count("N") for l in masked.splitlines() if not l.startswith(">"))
print(f"RepeatMasker -engine gpu -pa {a.pa} {a.extra}: {len(genome)} bp x {len(sseqs)} families: wall {wall:.1f} s, {len(out) - 3} annotation lines, "
      f"{100.0 * nN / len(genome):.1f} % masked")
print("\n".join(l for l in (srv.log or "").splitlines() if "served" in l))
if os.path.exists(trace):
    # where the run's wall time goes: before the first search reaches a GPU (FastaDB index, batching), while searches are
    # in flight, after the last one (ProcessRepeats, masking); busy = per-GPU sum of the passes' durations
    P = [l.split() for l in open(trace)]
    if P:
        st, en = [float(x[0]) for x in P], [float(x[1]) for x in P]
        busy = {}
        for x in P:
            busy[x[3]] = busy.get(x[3], 0.0) + float(x[1]) - float(x[0])
        print(f"phases: {min(st) - t_unix0:.1f} s before the first search, {max(en) - min(st):.1f} s with searches in flight, "
              f"{t_unix0 + wall - max(en):.1f} s after the last one; {len(P)} passes; GPU busy " +
              ", ".join(f"gpu{g} {b:.1f} s" for g, b in sorted(busy.items())))
print(open(os.path.join(d, "genome.fa.tbl")).read()[:700])
if a.also_crossmatch:
    d2 = tempfile.mkdtemp()
    for f in ("lib.fa", "genome.fa"):
        os.symlink(os.path.join(d, f), os.path.join(d2, f))
    cmd = f"source {H}/env.sh && cd {d2} && perl $RM_HOME/RepeatMasker -engine crossmatch -lib lib.fa -gc 41 -pa {a.also_crossmatch} {a.extra} genome.fa"
    t0 = time.perf_counter()
    res = subprocess.run(["bash", "-c", cmd], capture_output=True, text=True)
    wall2 = time.perf_counter() - t0
    if res.returncode:
        print(res.stdout[-1500:], res.stderr[-3000:]); sys.exit(1)
    print(f"RepeatMasker -engine crossmatch (restated crossmatch of this repo) -pa {a.also_crossmatch} on {os.cpu_count()} host cores: wall {wall2:.1f} s "
          f"({wall2 / wall:.0f}x the -engine gpu run)")
    strip = lambda t: "\n".join(l for l in t.splitlines() if not l.startswith("run with "))
    for ext in ("out", "masked", "cat", "tbl"):
        x, y = open(os.path.join(d, "genome.fa." + ext)).read(), open(os.path.join(d2, "genome.fa." + ext)).read()
        same = (x == y) if ext in ("out", "masked") else (strip(x) == strip(y))
        print(f"  genome.fa.{ext}: {'identical' if same else 'DIFFERENT'} ({len(x)} bytes)")
