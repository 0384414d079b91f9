"""Test helpers: ctypes view of the CPU oracle, fixture materialisation, cross_match text parser.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/ (it is the checker).
"""
import ctypes as C
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
ORACLE_BUILD = os.path.join(ROOT, "oracle", "_build")
ORACLE_BIN = os.path.join(ORACLE_BUILD, "cross_match")
MAXA = 16


class OraMatrix(C.Structure):
    _fields_ = [("n", C.c_int), ("alpha", C.c_char * (MAXA + 1)), ("score", (C.c_int * MAXA) * MAXA),
                ("freq", C.c_double * MAXA), ("lambda_", C.c_double), ("code_of", C.c_int * 256)]


class OraParams(C.Structure):
    _fields_ = [("gap_init", C.c_int), ("ins_gap_ext", C.c_int), ("del_gap_ext", C.c_int), ("minscore", C.c_int),
                ("masklevel", C.c_int), ("minmatch", C.c_int), ("bandwidth", C.c_int), ("raw", C.c_int),
                ("alignments", C.c_int)]


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ORACLE_BUILD, "librm_oracle.so"))
        _lib.ora_matrix_parse.argtypes = [C.c_char_p, C.POINTER(OraMatrix), C.c_char_p, C.c_int]
        _lib.ora_rescore.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(OraMatrix), C.POINTER(OraParams),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.ora_global_traceback.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(OraMatrix),
                                              C.POINTER(OraParams), C.c_int, C.c_char_p, C.c_char_p, C.c_char_p,
                                              C.c_char_p]
        _lib.ora_sw_score.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(OraMatrix),
                                      C.POINTER(OraParams)]
        if hasattr(_lib, "ora_sw_score_batch"):
            _lib.ora_sw_score_batch.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int,
                                                C.POINTER(OraMatrix), C.POINTER(OraParams), C.POINTER(C.c_int)]
            _lib.ora_sw_score_batch.restype = None
    return _lib


def matrices():
    return json.load(open(os.path.join(GOLD, "matrices.json")))["matrices"]


def write_matrix(name, path):
    """Re-materialise Matrices/crossmatch/<name> from the golden values (same grammar: optional FREQS
    line, alphabet row, integer rows)."""
    m = matrices()[name]
    with open(path, "w") as f:
        if m["has_freqs_line"]:
            f.write("FREQS " + " ".join(f"{a} {fr:g}" for a, fr in zip(m["alphabet"], m["freqs"]) if fr > 0) + "\n")
        f.write(" " + " ".join(f"{a:>3}" for a in m["alphabet"]) + "\n")
        for row in m["rows"]:
            f.write(" ".join(f"{v:>3}" for v in row) + "\n")
    return path


def parse_matrix(path):
    m = OraMatrix()
    err = C.create_string_buffer(512)
    rc = oracle_lib().ora_matrix_parse(path.encode(), C.byref(m), err, 512)
    assert rc == 0, err.value
    return m


def encode(m, letters):
    return bytes(m.code_of[ord(c)] for c in letters)


_HIT = re.compile(r"^\s*\d+\s+\d+\.\d+\s+\d+\.\d+\s+\d+\.\d+\s")
_ALN = re.compile(r"^(C?)\s+(\S+)\s+(\d+)\s+(\S+)\s+(\d+)\s*$")


def parse_crossmatch(text):
    """Minimal reader of cross_match text with the grammar CrossmatchSearchEngine.pm:549-879 accepts.
    Alignment rows are kept in DISPLAY orientation."""
    hits = []
    cur = None
    pos = 0
    for ln in text.splitlines():
        if ln.startswith("Score histogram:"):
            break
        if _HIT.match(ln):
            f = ln.split()
            star = f[-1] == "*"
            if star:
                f = f[:-1]
            if f[8] == "C":
                cur = dict(score=int(f[0]), div=f[1], dele=f[2], ins=f[3], qname=f[4], qstart=int(f[5]), qend=int(f[6]),
                           qleft=int(f[7].strip("()")), strand="C", sname=f[9], sleft=int(f[10].strip("()")),
                           send=int(f[11]), sstart=int(f[12]), star=star, qaln="", saln="")
            else:
                cur = dict(score=int(f[0]), div=f[1], dele=f[2], ins=f[3], qname=f[4], qstart=int(f[5]), qend=int(f[6]),
                           qleft=int(f[7].strip("()")), strand="+", sname=f[8], sstart=int(f[9]), send=int(f[10]),
                           sleft=int(f[11].strip("()")), star=star, qaln="", saln="")
            hits.append(cur)
            pos = 0
            continue
        m = _ALN.match(ln)
        if m and cur is not None:
            if pos == 0:
                cur["qaln"] += m.group(4)
            else:
                cur["saln"] += m.group(4)
            pos ^= 1
            continue
        mm = re.search(r"Transitions / transversions = \S+ \((\d+) / (\d+)\)", ln)
        if mm and cur is not None:
            cur["ti"], cur["tv"] = int(mm.group(1)), int(mm.group(2))
    return hits


def oracle_argv(query, subject, matrix, gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5, minscore=225, masklevel=101,
                minmatch=9, bandwidth=14, raw=False, alignments=True):
    argv = [ORACLE_BIN]
    if raw:
        argv.append("-raw")
    if alignments:
        argv.append("-alignments")
    argv += ["-gap_init", str(gap_init), "-ins_gap_ext", str(ins_gap_ext), "-del_gap_ext", str(del_gap_ext), "-minmatch",
             str(minmatch), "-minscore", str(minscore), "-bandwidth", str(bandwidth), "-masklevel", str(masklevel),
             "-matrix", matrix, query, subject]
    return argv


def run_oracle(query, subject, matrix, **kw):
    res = subprocess.run(oracle_argv(query, subject, matrix, **kw), capture_output=True, text=True, check=True)
    return parse_crossmatch(res.stdout), res.stdout


def read_fasta(path):
    names, seqs = [], []
    for ln in open(path):
        if ln.startswith(">"):
            names.append(ln[1:].split()[0] if ln[1:].split() else "")
            seqs.append([])
        elif names:
            seqs[-1].append(ln.strip())
    return names, ["".join(x) for x in seqs]


def union_library(first, second, out):
    """Records of `first`, then the records of `second` whose name `first` does not hold."""
    n1, s1 = read_fasta(first)
    n2, s2 = read_fasta(second)
    have = set(n1)
    with open(out, "w") as f:
        for n, s in list(zip(n1, s1)) + [(n, s) for n, s in zip(n2, s2) if n not in have]:
            f.write(f">{n}\n")
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
    return out


HARNESS = os.path.join(ROOT, "baseline", "_ref")


def harness_ready():
    import shutil

    return os.path.exists(os.path.join(HARNESS, "env.sh")) and shutil.which("perl") is not None


def run_repeatmasker(workdir, engine, args, query="genome.fa", lib="lib.fa", pa=4, env=None):
    """The reference driver itself (baseline/_ref, built by tools/make_rm_harness.sh): `-engine crossmatch` runs the
    oracle as cross_match, `-engine gpu` runs librmgpu through GpuSearchEngine.pm.  Returns its stdout."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "repeatmasker_b200", "perl")], check=True, capture_output=True)
    cmd = (f"source {HARNESS}/env.sh && cd {workdir} && perl $RM_HOME/RepeatMasker -engine {engine} -lib {lib} "
           f"-pa {pa} {' '.join(args)} {query}")
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run(["bash", "-c", cmd], capture_output=True, text=True, env=e)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


def driver_outputs(workdir, query="genome.fa", exts=("out", "masked", "cat", "tbl")):
    out = {}
    for ext in exts:
        p = os.path.join(workdir, f"{query}.{ext}")
        out[ext] = open(p).read() if os.path.exists(p) else None
    return out


def strip_engine_trailer(text):
    """.cat / .tbl name the engine and its version ("run with ..."); everything else must be identical."""
    return None if text is None else "\n".join(l for l in text.splitlines() if not l.startswith("run with "))


def run_oracle_parallel(query, snames, sseqs, matrix, workdir, procs=None, **kw):
    """The oracle over a whole library on several processes: the subjects are dealt round-robin to `procs` oracle
    processes (each sees the one query file), the hit lists are merged into the specified output order and the '*' flags
    (which depend on every hit of the query) are recomputed.  masklevel 101 only; one query record."""
    from repeatmasker_b200 import synth

    assert kw.get("masklevel", 101) >= 101
    procs = procs or os.cpu_count() or 1
    procs = max(1, min(procs, len(snames)))
    # longest subjects first, dealt round-robin: even work per process
    order = sorted(range(len(snames)), key=lambda i: -len(sseqs[i]))
    running = []
    for t in range(procs):
        sub = sorted(order[t::procs])
        sp = os.path.join(workdir, f"ora_part{t}.fa")
        synth.write_fasta(sp, [snames[i] for i in sub], [sseqs[i] for i in sub])
        running.append(subprocess.Popen(oracle_argv(query, sp, matrix, **kw), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
    hits = []
    for p in running:
        out, _ = p.communicate()
        assert p.returncode == 0
        hits += parse_crossmatch(out)
    sidx = {n: i for i, n in enumerate(snames)}
    hits.sort(key=lambda h: (h["qstart"], -sidx[h["sname"]], h["sstart"], h["qend"], h["strand"] == "C", h["send"]))
    for h in hits:
        h["star"] = any(o["score"] > h["score"] and o["qstart"] <= h["qend"] and o["qend"] >= h["qstart"] for o in hits)
    return hits
