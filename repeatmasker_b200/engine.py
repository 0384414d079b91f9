"""ctypes binding of librmgpu.so and a Python mirror of the reference's engine interface.

`GpuSearchEngine` follows SearchEngineI.pm:132-681 (the same setters) and CrossmatchSearchEngine.pm:381-524
(`search()` returning `(status, results)`), so parity tests read like the reference's own engine tests.  The Perl
plug-in that RepeatMasker itself loads is repeatmasker_b200/perl/GpuSearchEngine.pm; both sit on the same C ABI
(include/rmgpu.h).  There is no CPU fallback: without the CUDA library or a device every call raises.
"""
import ctypes as C
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RMGPU_LIB") or os.path.join(HERE, "librmgpu.so")


class RmgpuError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("gap_init", C.c_int32), ("ins_gap_ext", C.c_int32), ("del_gap_ext", C.c_int32),
                ("minscore", C.c_int32), ("masklevel", C.c_int32), ("minmatch", C.c_int32), ("bandwidth", C.c_int32),
                ("raw", C.c_int32), ("word_raw", C.c_int32), ("want_alignments", C.c_int32), ("device", C.c_int32),
                ("reserved", C.c_int32 * 5)]


class Hit(C.Structure):
    _fields_ = [("score", C.c_int32), ("raw_score", C.c_int32), ("query_index", C.c_int32),
                ("subject_index", C.c_int32), ("query_name", C.c_char_p), ("subject_name", C.c_char_p),
                ("qstart", C.c_int32), ("qend", C.c_int32), ("qleft", C.c_int32), ("sstart", C.c_int32),
                ("send", C.c_int32), ("sleft", C.c_int32), ("strand", C.c_uint8), ("overlap_star", C.c_uint8),
                ("pad_", C.c_uint8 * 2), ("pct_div", C.c_char * 8), ("pct_del", C.c_char * 8),
                ("pct_ins", C.c_char * 8), ("transitions", C.c_int32), ("transversions", C.c_int32),
                ("ambiguous", C.c_int32), ("gap_inits", C.c_int32), ("gap_bases", C.c_int32), ("qaln", C.c_char_p),
                ("saln", C.c_char_p), ("pct_kimura", C.c_char * 8), ("cpg_sites", C.c_int32),
                ("well_characterized", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("cells", C.c_double), ("ms_total", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double),
                ("ms_score", C.c_double), ("ms_origin", C.c_double), ("ms_traceback", C.c_double),
                ("ms_host_post", C.c_double), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("launches", C.c_int32), ("score_launches", C.c_int32), ("n_pairs_passed", C.c_int32),
                ("n_candidates", C.c_int32), ("origin_cells", C.c_double), ("n_origin_redo", C.c_int32),
                ("pad_", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("n_hits", C.c_int32), ("hits", C.POINTER(Hit)), ("stats", Stats), ("opaque", C.c_void_p)]


_lib = None

EXPORTS = ["rmgpu_version", "rmgpu_abi_version", "rmgpu_last_error", "rmgpu_search", "rmgpu_free",
           "rmgpu_session_open", "rmgpu_session_close", "rmgpu_session_set_matrix_file", "rmgpu_session_set_library",
           "rmgpu_session_set_library_fasta", "rmgpu_session_set_queries", "rmgpu_session_set_queries_fasta",
           "rmgpu_session_search", "rmgpu_session_score_only", "rmgpu_session_counts", "rmgpu_measure_dpx_peak",
           "rmgpu_server_ping", "rmgpu_server_quit", "rmgpu_kimura_divergence", "rmgpu_matrix_info",
           "rmgpu_adjusted_score", "rmgpu_masklevel_filter", "rmgpu_contained_filter",
           "rmgpu_session_library_h2d_bytes", "rmgpu_session_set_queries_fasta_multi", "rmgpu_device_count",
           "rmgpu_session_mask_queries", "rmgpu_format_alignment", "rmgpu_debug_k1_schedule"]


def lib():
    """Load librmgpu.so (built in-tree by repeatmasker_b200/build.py).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RmgpuError(f"{LIB_PATH} is missing: run `python -m repeatmasker_b200.build` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    L.rmgpu_version.restype = C.c_char_p
    L.rmgpu_last_error.restype = C.c_char_p
    L.rmgpu_search.argtypes = [C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.POINTER(Result))]
    L.rmgpu_free.argtypes = [C.POINTER(Result)]
    L.rmgpu_session_open.restype = C.c_void_p
    L.rmgpu_session_open.argtypes = [C.c_int]
    L.rmgpu_session_close.argtypes = [C.c_void_p]
    L.rmgpu_session_set_matrix_file.argtypes = [C.c_void_p, C.c_char_p]
    seqargs = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]
    L.rmgpu_session_set_library.argtypes = seqargs
    L.rmgpu_session_set_queries.argtypes = seqargs
    L.rmgpu_session_set_library_fasta.argtypes = [C.c_void_p, C.c_char_p]
    L.rmgpu_session_set_queries_fasta.argtypes = [C.c_void_p, C.c_char_p]
    L.rmgpu_session_search.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(C.POINTER(Result))]
    L.rmgpu_session_score_only.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.rmgpu_session_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64)]
    L.rmgpu_session_library_h2d_bytes.argtypes = [C.c_void_p]
    L.rmgpu_session_mask_queries.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.rmgpu_session_library_h2d_bytes.restype = C.c_int64
    L.rmgpu_measure_dpx_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    fargs = [C.c_int, C.c_int32] + [C.POINTER(C.c_int32)] * 4 + [C.c_int32, C.POINTER(C.c_uint8)]
    L.rmgpu_masklevel_filter.argtypes = fargs
    L.rmgpu_contained_filter.argtypes = fargs
    L.rmgpu_server_ping.argtypes = [C.c_char_p]
    L.rmgpu_server_quit.argtypes = [C.c_char_p]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise RmgpuError(lib().rmgpu_last_error().decode())


def _hits_to_dicts(res):
    out = []
    r = res.contents
    for i in range(r.n_hits):
        h = r.hits[i]
        out.append(dict(score=h.score, raw_score=h.raw_score, div=h.pct_div.decode(), dele=h.pct_del.decode(),
                        ins=h.pct_ins.decode(), qname=h.query_name.decode(), qstart=h.qstart, qend=h.qend,
                        qleft=h.qleft, strand="C" if h.strand else "+", sname=h.subject_name.decode(),
                        sstart=h.sstart, send=h.send, sleft=h.sleft, star=bool(h.overlap_star),
                        qaln=(h.qaln or b"").decode(), saln=(h.saln or b"").decode(), ti=h.transitions,
                        tv=h.transversions, query_index=h.query_index, subject_index=h.subject_index,
                        kimura=h.pct_kimura.decode(), cpg=h.cpg_sites))
    st = {k: getattr(r.stats, k) for k, _ in Stats._fields_}
    return out, st


def collection_filter(names, starts, ends, scores, masklevel, contained=False, device=-1):
    """rmgpu_masklevel_filter / rmgpu_contained_filter: the survivors' indices, in collection order."""
    n = len(names)
    ids = {x: k for k, x in enumerate(sorted(set(names)))}  # ranks in string order, like perl's cmp
    nid = (C.c_int32 * n)(*[ids[x] for x in names])
    arr = lambda v: (C.c_int32 * n)(*v)
    keep = (C.c_uint8 * n)()
    fn = lib().rmgpu_contained_filter if contained else lib().rmgpu_masklevel_filter
    _check(fn(device, n, nid, arr(starts), arr(ends), arr(scores), masklevel, keep))
    return [i for i in range(n) if keep[i]]


def kimura_divergence(qaln, saln, strand_c=False):
    """rmgpu_kimura_divergence(): (text "%4.2f", transitions, transversions, well-characterised pairs, CpG sites) of one
    alignment in the orientation of rmgpu_hit::qaln/saln (SearchResult.pm:1607-1683).  Host-only, needs no GPU."""
    text = C.create_string_buffer(8)
    ti, tv, wcb, cpg = C.c_double(), C.c_int32(), C.c_int32(), C.c_int32()
    f = lib().rmgpu_kimura_divergence
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int32),
                  C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    f.restype = None
    f(qaln.encode(), saln.encode(), 1 if strand_c else 0, text, C.byref(ti), C.byref(tv), C.byref(wcb), C.byref(cpg))
    return text.value.decode(), ti.value, tv.value, wcb.value, cpg.value


def format_alignment(query_name, subject_name, qaln, saln, qstart, qend, sstart, send, strand_c=False, matrix_name=None,
                     pct_kimura=None, cpg_sites=None, pct_raw_kimura=None):
    """rmgpu_format_alignment(): the alignment block of a .cat entry (SearchResult.pm:2486-2834, mode AlignWithQuerySeq)
    from the fields a SearchResult holds (rows with the query forward).  Host-only, needs no GPU."""
    f = lib().rmgpu_format_alignment
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                  C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64]
    f.restype = C.c_int64
    opt = [None if v is None else str(v).encode() for v in (matrix_name, pct_kimura, cpg_sites, pct_raw_kimura)]
    args = [query_name.encode(), subject_name.encode(), qaln.encode(), saln.encode(), qstart, qend, sstart, send,
            1 if strand_c else 0] + opt
    need = f(*args, None, 0)
    if need < 0:
        raise ValueError("alignment rows differ in length")
    buf = C.create_string_buffer(need + 1)
    f(*args, buf, need + 1)
    return buf.value.decode()


def k1_schedule(n_tickets, n_seg, seg_window, claim, n_blocks, seg_min_blocks, level):
    """rmgpu_debug_k1_schedule(): (has_work, ticket, segment, first block, end block) of one claim of the score pass's
    column-segment schedule.  Host-only test hook."""
    f = lib().rmgpu_debug_k1_schedule
    f.argtypes = [C.c_int32] * 7 + [C.POINTER(C.c_int32)] * 4
    f.restype = C.c_int
    t, s, b0, b1 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    rc = f(n_tickets, n_seg, seg_window, claim, n_blocks, seg_min_blocks, level, C.byref(t), C.byref(s), C.byref(b0), C.byref(b1))
    if rc < 0:
        raise ValueError("bad schedule arguments")
    return rc, t.value, s.value, b0.value, b1.value


def matrix_info(path):
    """rmgpu_matrix_info(): (n, alphabet, lambda, freqs) as the library's matrix reader sees the file.  Host-only."""
    n, lam = C.c_int32(), C.c_double()
    alpha = C.create_string_buffer(17)
    fr = (C.c_double * 16)()
    f = lib().rmgpu_matrix_info
    f.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _check(f(path.encode(), C.byref(n), alpha, C.byref(lam), fr))
    return n.value, alpha.value.decode(), lam.value, list(fr)[: n.value]


def adjusted_score(matrix_path, raw, qaln, saln):
    """rmgpu_adjusted_score(): complexity-adjusted score of one alignment (SearchResult.pm:2111-2150).  Host-only."""
    adj = C.c_int32()
    f = lib().rmgpu_adjusted_score
    f.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32)]
    _check(f(matrix_path.encode(), raw, qaln.encode(), saln.encode(), C.byref(adj)))
    return adj.value


def make_params(gap_init=-30, ins_gap_ext=-6, del_gap_ext=-5, minscore=225, masklevel=101, minmatch=9, bandwidth=14,
                raw=False, word_raw=False, alignments=True, device=-1):
    return Params(gap_init=gap_init, ins_gap_ext=ins_gap_ext, del_gap_ext=del_gap_ext, minscore=minscore,
                  masklevel=masklevel, minmatch=minmatch, bandwidth=bandwidth, raw=int(raw), word_raw=int(word_raw),
                  want_alignments=int(alignments), device=device)


def _seq_arrays(names, seqs):
    n = len(seqs)
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    bn = [(names[i] if names else f"seq{i}").encode() for i in range(n)]
    return (n, (C.c_char_p * n)(*bn), (C.c_char_p * n)(*bs), (C.c_int32 * n)(*[len(b) for b in bs]))


class Records:
    """Host-side result of one search as the library returns it (rmgpu_result: flat rmgpu_hit records + row buffer)."""

    def __init__(self, res):
        self._res = res
        self.stats = {k: getattr(res.contents.stats, k) for k, _ in Stats._fields_}

    def __len__(self):
        return self._res.contents.n_hits

    def to_dicts(self):
        return _hits_to_dicts(self._res)[0]

    def __del__(self):
        if self._res is not None:
            lib().rmgpu_free(self._res)
            self._res = None


class Session:
    """One GPU worker: matrix + library resident in HBM, query batches streamed through (rmgpu_session_*)."""

    def __init__(self, device=-1):
        self._h = lib().rmgpu_session_open(device)
        if not self._h:
            raise RmgpuError(lib().rmgpu_last_error().decode())

    def close(self):
        if self._h:
            lib().rmgpu_session_close(self._h)
            self._h = None

    __del__ = close

    def set_matrix(self, path):
        _check(lib().rmgpu_session_set_matrix_file(self._h, path.encode()))

    def set_library(self, seqs, names=None):
        _check(lib().rmgpu_session_set_library(self._h, *_seq_arrays(names, seqs)))

    def set_library_fasta(self, path):
        _check(lib().rmgpu_session_set_library_fasta(self._h, path.encode()))

    def set_queries(self, seqs, names=None):
        _check(lib().rmgpu_session_set_queries(self._h, *_seq_arrays(names, seqs)))

    def set_queries_fasta(self, path):
        _check(lib().rmgpu_session_set_queries_fasta(self._h, path.encode()))

    def mask_queries(self, ranges):
        """Mask (query index, qstart, qend) ranges -- 1-based inclusive, as in the hit records -- of the RESIDENT batch
        with X on the device: what RepeatMasker does to a batch between two stages (RepeatMasker:6036-6090)."""
        n = len(ranges)
        arr = lambda k: (C.c_int32 * max(n, 1))(*[int(r[k]) for r in ranges])
        _check(lib().rmgpu_session_mask_queries(self._h, n, arr(0), arr(1), arr(2)))

    def counts(self):
        nq, ns, qb, sb = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
        lib().rmgpu_session_counts(self._h, C.byref(nq), C.byref(ns), C.byref(qb), C.byref(sb))
        return nq.value, ns.value, qb.value, sb.value

    def library_h2d_bytes(self):
        return lib().rmgpu_session_library_h2d_bytes(self._h)

    def search_records(self, params):
        """The search with the result left as the C records rmgpu_session_search() returns (what the XS glue walks):
        a `Records` object -- len(), .stats, .to_dicts() -- that frees the result when it is dropped."""
        res = C.POINTER(Result)()
        _check(lib().rmgpu_session_search(self._h, C.byref(params), C.byref(res)))
        return Records(res)

    def search(self, params):
        res = C.POINTER(Result)()
        dbg = "RMGPU_DEBUG_HOST" in os.environ
        t0 = time.perf_counter()
        _check(lib().rmgpu_session_search(self._h, C.byref(params), C.byref(res)))
        t1 = time.perf_counter()
        try:
            return _hits_to_dicts(res)
        finally:
            lib().rmgpu_free(res)
            if dbg:
                print(f"[host dbg] Session.search: C call {1e3 * (t1 - t0):.2f} ms, records -> dicts + free {1e3 * (time.perf_counter() - t1):.2f} ms",
                      file=sys.stderr)

    def score_only(self, params):
        """K1 only: (best[nq][ns][2] as a flat list, kernel ms)."""
        nq, ns, _, _ = self.counts()
        buf = (C.c_int32 * max(nq * ns * 2, 1))()
        ms = C.c_double()
        _check(lib().rmgpu_session_score_only(self._h, C.byref(params), buf, C.byref(ms)))
        return list(buf)[: nq * ns * 2], ms.value


def search_files(query, subject, matrix, **kw):
    """rmgpu_search(): the call that stands where CrossmatchSearchEngine.pm:444 popen()s cross_match."""
    p = make_params(**kw)
    res = C.POINTER(Result)()
    _check(lib().rmgpu_search(C.byref(p), query.encode(), subject.encode(), matrix.encode(), C.byref(res)))
    try:
        return _hits_to_dicts(res)
    finally:
        lib().rmgpu_free(res)


SERVER_BIN = os.path.join(HERE, "rmgpu_server")


class Server:
    """Context manager around the persistent worker process (rmgpu_server): `with Server(sock, gpus=[0]) as s: ...`
    While it is open, RMGPU_SERVER points at it, so rmgpu_search() / search_files() / GpuSearchEngine.search() in this
    process and in its children run as thin clients."""

    def __init__(self, socket_path, gpus=(0,), trace=None):
        self.socket_path, self.gpus, self.proc, self._old, self.trace = socket_path, list(gpus), None, None, trace

    def __enter__(self):
        import subprocess
        import time

        argv = [SERVER_BIN, "--socket", self.socket_path, "--gpus", ",".join(map(str, self.gpus))]
        if self.trace:
            argv += ["--trace", self.trace]
        self.proc = subprocess.Popen(argv, stderr=subprocess.PIPE, text=True)
        for _ in range(200):
            if os.path.exists(self.socket_path) and lib().rmgpu_server_ping(self.socket_path.encode()) == 0:
                break
            if self.proc.poll() is not None:
                raise RmgpuError("rmgpu_server exited: " + self.proc.stderr.read())
            time.sleep(0.05)
        else:
            raise RmgpuError("rmgpu_server did not come up")
        self._old = os.environ.get("RMGPU_SERVER")
        os.environ["RMGPU_SERVER"] = self.socket_path
        return self

    def __exit__(self, *exc):
        if self._old is None:
            os.environ.pop("RMGPU_SERVER", None)
        else:
            os.environ["RMGPU_SERVER"] = self._old
        try:
            lib().rmgpu_server_quit(self.socket_path.encode())
            self.proc.wait(timeout=20)
        except Exception:
            self.proc.kill()
        self.log = self.proc.stderr.read()
        return False


def measure_dpx_peak(device=-1):
    cups, ips = C.c_double(), C.c_double()
    _check(lib().rmgpu_measure_dpx_peak(device, C.byref(cups), C.byref(ips)))
    return cups.value, ips.value


class GpuSearchEngine:
    """Python mirror of the engine plug-in contract (SearchEngineI.pm), GPU sibling of CrossmatchSearchEngine.pm."""

    basicScoreMode = 1  # SearchEngineI.pm:77
    complexityAdjustedScoreMode = 2  # SearchEngineI.pm:78

    def __init__(self, device=-1):
        self._v = dict(query=None, subject=None, matrix=None, gapInit=None, insGapExt=None, delGapExt=None,
                       minMatch=None, minScore=None, bandwidth=None, maskLevel=None, scoreMode=2,
                       generateAlignments=0, wordRaw=0, device=device)
        self.last_stats = None

    def __getattr__(self, name):
        if name.startswith("set") or name.startswith("get"):
            key = name[3].lower() + name[4:]
            if key in self.__dict__.get("_v", {}):
                if name.startswith("set"):
                    def setter(value):
                        old = self._v[key]
                        self._v[key] = value
                        return old
                    return setter
                return lambda: self._v[key]
        raise AttributeError(name)

    def getVersion(self):
        return lib().rmgpu_version().decode()

    def getParameters(self):
        v = self._v
        for k in ("query", "subject", "matrix"):
            if not v[k] or not os.path.isfile(v[k]):
                raise RmgpuError(f"GpuSearchEngine::getParameters: Error...{k} ({v[k]}) does not exist!")
        s = "rmgpu"
        if v["scoreMode"] == self.basicScoreMode:
            s += " -raw"
        if v["wordRaw"]:
            s += " -word_raw"
        if v["generateAlignments"]:
            s += " -alignments"
        for opt, key in (("gap_init", "gapInit"), ("ins_gap_ext", "insGapExt"), ("del_gap_ext", "delGapExt"),
                         ("minmatch", "minMatch"), ("minscore", "minScore"), ("bandwidth", "bandwidth"),
                         ("masklevel", "maskLevel")):
            if v[key] is not None:
                s += f" -{opt} {v[key]}"
        return s + f" -matrix {v['matrix']} {v['query']} {v['subject']}"

    def search(self, **kw):
        """Returns (status, hits) like CrossmatchSearchEngine.pm:381-524; status 0 = success."""
        for k, val in kw.items():
            getattr(self, "set" + k[0].upper() + k[1:])(val)
        self.getParameters()  # croaks on missing files like the reference
        v = self._v
        d = lambda x, dflt: dflt if x is None else x
        hits, st = search_files(v["query"], v["subject"], v["matrix"], gap_init=d(v["gapInit"], -4),
                                ins_gap_ext=d(v["insGapExt"], -3), del_gap_ext=d(v["delGapExt"], -3),
                                minscore=d(v["minScore"], 30), masklevel=d(v["maskLevel"], 80),
                                minmatch=d(v["minMatch"], 14), bandwidth=d(v["bandwidth"], 14),
                                raw=v["scoreMode"] == self.basicScoreMode, word_raw=bool(v["wordRaw"]),
                                alignments=bool(v["generateAlignments"]), device=v["device"])
        self.last_stats = st
        return 0, hits
