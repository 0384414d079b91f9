/*
 * rmgpu.h -- C ABI of librmgpu.so, the B200 (sm_100a) repeat-search engine that stands where
 * RepeatMasker's CrossmatchSearchEngine.pm popen()s the external `cross_match` binary.
 *
 * Plain C types only (no torch / CUDA types): this is what an FFI/XS stub binds.  The Perl-side
 * binding is repeatmasker_b200/perl/{GpuSearchEngine.pm,RMGpu.xs}; INTEGRATION.md shows the two
 * driver edits.  Each entry point cites the reference interface it replaces (paths are inside the
 * RepeatMasker tree).
 *
 * Threading / fork: no CUDA call is made before the first search in a given PID
 * (RepeatMasker forks one child per batch, RepeatMasker:1043; a CUDA context cannot cross fork).
 * A session is bound to the thread's current device and must be used from one thread at a time.
 */
#ifndef RMGPU_H
#define RMGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMGPU_ABI_VERSION 4

/* Search parameters.  Replaces the argv that CrossmatchSearchEngine.pm:277-363 (getParameters) builds from the
 * SearchEngineI.pm:132-681 setters:
 *   -raw (scoreMode==basicScoreMode) -word_raw -alignments -gap_init -ins_gap_ext -del_gap_ext
 *   -minmatch -minscore -bandwidth -masklevel -matrix <file> <query> <subject>                                  */
typedef struct rmgpu_params {
  int32_t gap_init;        /* setGapInit,   e.g. -30: cost of the first base of a gap                  */
  int32_t ins_gap_ext;     /* setInsGapExt, e.g. -6 : further bases of a gap in the SUBJECT row          */
  int32_t del_gap_ext;     /* setDelGapExt, e.g. -5 : further bases of a gap in the QUERY row            */
  int32_t minscore;        /* setMinScore                                                               */
  int32_t masklevel;       /* setMaskLevel 0..101 (RepeatMasker:2920-2926 always passes 101)            */
  int32_t minmatch;        /* setMinMatch: accepted for interface parity; discovery is exhaustive       */
  int32_t bandwidth;       /* setBandwidth: initial half-width of the traceback band                    */
  int32_t raw;             /* 1 = basicScoreMode (-raw), 0 = complexityAdjustedScoreMode                */
  int32_t word_raw;        /* setWordRaw: accepted, unused                                              */
  int32_t want_alignments; /* setGenerateAlignments                                                     */
  int32_t device;          /* CUDA device ordinal, -1 = current                                         */
  int32_t reserved[5];
} rmgpu_params;

/* One reported match.  Replaces one hit line + alignment block of cross_match text as parsed by
 * CrossmatchSearchEngine.pm:549-879 into SearchResult->new(...) (field list SearchResult.pm:235-262).      */
typedef struct rmgpu_hit {
  int32_t score;     /* reported score (complexity-adjusted unless params.raw)                         */
  int32_t raw_score; /* Smith-Waterman-Gotoh score                                                      */
  int32_t query_index, subject_index; /* 0-based record numbers in the query / subject input          */
  const char *query_name, *subject_name; /* owned by the result                                        */
  int32_t qstart, qend, qleft; /* 1-based inclusive, forward-strand query coordinates                  */
  int32_t sstart, send, sleft;
  uint8_t strand;       /* 0 forward, 1 = "C" (reverse-complemented query vs forward subject)          */
  uint8_t overlap_star; /* the trailing '*' of the hit line                                             */
  uint8_t pad_[2];
  char pct_div[8], pct_del[8], pct_ins[8]; /* the "%.2f" text cross_match prints (query-span denominators) */
  int32_t transitions, transversions, ambiguous, gap_inits, gap_bases;
  const char *qaln, *saln; /* NUL-terminated gapped rows in cross_match DISPLAY orientation: for strand C the
                              query row is reverse-complemented (descending coordinates), subject row forward.
                              NULL when want_alignments == 0.                                           */
  char pct_kimura[8];      /* sprintf("%4.2f") of SearchResult::calcKimuraDivergence(divCpGMod => 1)
                              (SearchResult.pm:1607-1683), as CrossmatchSearchEngine.pm:509-511 stores it; computed
                              on the rows as SearchResult holds them (strand C: both reversed and complemented).
                              "" when want_alignments == 0.                                             */
  int32_t cpg_sites, well_characterized; /* by-products of the same pass                                */
} rmgpu_hit;

/* Device-side accounting of one search (all times from CUDA events on the engine's stream). */
typedef struct rmgpu_stats {
  double cells;          /* DP cells of the exhaustive pass = sum Q*L*2                                 */
  double ms_total;       /* whole call, host clock                                                      */
  double ms_h2d, ms_d2h; /* copies                                                                      */
  double ms_score;       /* K1: strand-paired int16x2 DPX score pass (the dominant kernel)              */
  double ms_origin;      /* K2: int32 origin-tracking pass over (query,strand,subject) triples that passed K1 */
  double ms_traceback;   /* K3: banded traceback                                                        */
  double ms_host_post;   /* adjusted score, percentages, ordering, masklevel, strings                   */
  int64_t h2d_bytes, d2h_bytes;
  int32_t launches;      /* kernels launched                                                            */
  int32_t score_launches;
  int32_t n_pairs_passed; /* (query,strand,subject) triples with K1 best >= minscore                    */
  int32_t n_candidates;   /* origins with raw >= minscore                                               */
  double origin_cells;    /* cells K2 recomputed                                                        */
  int32_t n_origin_redo;  /* K2 tasks redone with origin propagation over the whole fragment            */
  int32_t pad_;
} rmgpu_stats;

typedef struct rmgpu_result {
  int32_t n_hits;
  rmgpu_hit *hits;
  rmgpu_stats stats;
  void *opaque; /* string arena */
} rmgpu_result;

/* "cross_match version ..." banner replacement (CrossmatchSearchEngine.pm:244-260 scrapes a version that the driver
 * prints in the .cat/.tbl trailer, RepeatMasker:1267-1270). */
const char *rmgpu_version(void);
int rmgpu_abi_version(void);

/* Last error text of the calling thread ("" if none).  Errors are also returned as non-zero status, which the
 * driver's retry ladder interprets (RepeatMasker:2068-2107). */
const char *rmgpu_last_error(void);

/* The one call that replaces `open("$cmdLine 2>/dev/null |")` + parseOutput (CrossmatchSearchEngine.pm:444-480):
 * query FASTA, subject FASTA and matrix file paths as set by setQuery/setSubject/setMatrix.
 * Returns 0 on success; *out must be released with rmgpu_free. Host buffers in, host buffers out.             */
int rmgpu_search(const rmgpu_params *p, const char *query_fasta, const char *subject_fasta, const char *matrix_file,
                 rmgpu_result **out);
void rmgpu_free(rmgpu_result *r);

/* ---- Session API: library resident in HBM across batches (one session per GPU worker). ---------- */
typedef struct rmgpu_session rmgpu_session;

rmgpu_session *rmgpu_session_open(int device); /* NULL on failure (no CUDA device => loud error, no fallback) */
void rmgpu_session_close(rmgpu_session *s);
/* Matrix file in the Matrix.pm:308-364 grammar (FREQS line, alphabet row, integer rows). */
int rmgpu_session_set_matrix_file(rmgpu_session *s, const char *matrix_file);
/* Subject library: n records, upper/lower-case IUPAC letters; packed to the device layout and kept resident. */
int rmgpu_session_set_library(rmgpu_session *s, int32_t n, const char *const *names, const char *const *seqs,
                              const int32_t *lens);
int rmgpu_session_set_library_fasta(rmgpu_session *s, const char *subject_fasta);
/* Query batch (SimpleBatcher fragments, SimpleBatcher.pm:592-800): encoded and copied to the device. */
int rmgpu_session_set_queries(rmgpu_session *s, int32_t n, const char *const *names, const char *const *seqs,
                              const int32_t *lens);
int rmgpu_session_set_queries_fasta(rmgpu_session *s, const char *query_fasta);
/* The batch files of several callers as ONE batch (what rmgpu_server does with the pending calls of RepeatMasker's forked
 * children, RepeatMasker:1040-1180): records of file f are first_record[f] .. first_record[f+1]-1 of the batch
 * (first_record holds n_files + 1 entries); rmgpu_hit::query_index tells the results apart. */
int rmgpu_session_set_queries_fasta_multi(rmgpu_session *s, int32_t n_files, const char *const *files, int32_t *first_record);
/* Stage chain on a resident batch: mask what a stage found -- n ranges, record index as in
 * rmgpu_hit::query_index, 1-based inclusive coordinates -- with X, on the device, as RepeatMasker does between two
 * stages of a -species run (`$seqDB->setSubstr( $name, $begin - 1, $len, 'X' x $len )`, RepeatMasker:6036-6090; stage
 * chain :4115-4654).  The next rmgpu_session_search (usually after rmgpu_session_set_library* with the next stage's
 * library, which may already be resident) sees the masked batch; nothing is uploaded again.  (Added in ABI 3.) */
int rmgpu_session_mask_queries(rmgpu_session *s, int32_t n, const int32_t *query_index, const int32_t *qstart,
                               const int32_t *qend);
/* Run the search on the resident batch x resident library. */
int rmgpu_session_search(rmgpu_session *s, const rmgpu_params *p, rmgpu_result **out);
/* K1 only, on resident data: best raw Smith-Waterman-Gotoh score per (query, subject, strand).
 * best must hold n_queries*n_subjects*2 int32 (strand fastest).  Used by tests and by bench.py's kernel-only leg;
 * returns the kernel time in *ms (CUDA events). */
int rmgpu_session_score_only(rmgpu_session *s, const rmgpu_params *p, int32_t *best, double *ms);
/* Bytes set_library copied to the device (row codes in 16-row groups, 1 B/base codes for K2-K4, group / stripe tables); the
 * 8 B/row score profiles are built on the device. */
int64_t rmgpu_session_library_h2d_bytes(const rmgpu_session *s);
int rmgpu_session_counts(const rmgpu_session *s, int32_t *n_queries, int32_t *n_subjects, int64_t *query_bases,
                         int64_t *subject_bases);

/* ---- Persistent worker (repeatmasker_b200/rmgpu_server).  With RMGPU_SERVER=<unix socket path> in the environment,
 * rmgpu_search() makes no CUDA call in the calling process: it sends the request to the server, which keeps one
 * session (matrix + packed library resident in HBM) per GPU.  This is what a forked, short-lived RepeatMasker batch child
 * (RepeatMasker:1043-1180) should use.  The two calls below are management helpers: 0 = server answered.            */
int rmgpu_device_count(void); /* visible CUDA devices (0 and an error text when there is none); `rmgpu_server --gpus all` */
int rmgpu_server_ping(const char *socket_path);
int rmgpu_server_quit(const char *socket_path);

/* ---- The two collection filters RepeatMasker runs on an engine's results (RepeatMasker:2838-2987), on the device, with
 * the reference's own arithmetic and tie order.  n records in collection order: an integer per distinct query name that
 * orders like the names do under perl's `cmp` (the contained filter's scan crosses name boundaries, as the reference's
 * does), 1-based inclusive query coordinates, score.  keep[i] = 1 for the survivors; the order of the collection is not changed.
 *   rmgpu_masklevel_filter   SearchResultCollection::maskLevelFilter( value => masklevel )   SearchResultCollection.pm:198-358
 *                            (sort by name, score descending, span ascending; pairwise sprintf("%0.0f") test against every
 *                            kept domain; masklevel >= 101 keeps everything)
 *   rmgpu_contained_filter   SearchResultCollection::filterContainedResults( value => masklevel )     :379-470
 * device -1 = current device.  Both return 0 on success. */
int rmgpu_masklevel_filter(int device, int32_t n, const int32_t *name_id, const int32_t *qstart, const int32_t *qend,
                           const int32_t *score, int32_t masklevel, uint8_t *keep);
int rmgpu_contained_filter(int device, int32_t n, const int32_t *name_id, const int32_t *qstart, const int32_t *qend,
                           const int32_t *score, int32_t masklevel, uint8_t *keep);

/* Measured integer-SIMD ("DPX") peak of this GPU: a register-resident loop of the bare recurrence.
 * Returns 16-bit-lane cell updates per second (GCUPS*1e9) in *cups and raw VIADDMNMX.S16x2 warp-instructions/s. */
int rmgpu_measure_dpx_peak(int device, double *cups, double *dpx_instr_per_s);

/* Host-only test hook for the score pass's column segments: claim number `claim` of a launch with n_tickets
 * tickets cut into n_seg segments and claimed in windows of seg_window tickets -> the ticket and segment it stands for and
 * the 32-column blocks [*bs, *be) of that segment for a record of n_blocks blocks whose stripe lies at depth `level` of its
 * chain (records with fewer than seg_min_blocks blocks are not cut).  Computed by the functions the kernel itself calls.
 * Returns 1 (the sub-ticket has work), 0 (empty: a short record's segment > 0) or -1 (bad arguments).
 * tests/test_k1_schedule.py checks on the CPU that every sub-ticket only depends on sub-tickets claimed before it. */
int rmgpu_debug_k1_schedule(int32_t n_tickets, int32_t n_seg, int32_t seg_window, int32_t claim, int32_t n_blocks,
                            int32_t seg_min_blocks, int32_t level, int32_t *ticket, int32_t *seg, int32_t *bs, int32_t *be);

/* Host-only helper behind rmgpu_hit::pct_kimura, exported so that it can be checked against the reference's Perl
 * (SearchResult.pm:1607-1683) without a GPU: rows in the orientation of rmgpu_hit::qaln/saln, strand as in rmgpu_hit.
 * Writes the "%4.2f" text; the optional outputs are transitions (fractional with the CpG rule), transversions,
 * well-characterised pairs and CpG sites. */
void rmgpu_kimura_divergence(const char *qaln, const char *saln, int strand, char text[8], double *transitions,
                             int32_t *transversions, int32_t *well_characterized, int32_t *cpg_sites);

/* Host-only views of the two other double-precision pieces of the path, exported for the same reason:
 * the matrix reader (grammar of Matrix.pm:308-364, lambda of Matrix.pm:397-464) and the complexity adjustment of a
 * raw score over the aligned query letters of one alignment (SearchResult.pm:2111-2150).  Both return 0 on success. */
int rmgpu_matrix_info(const char *matrix_file, int32_t *n, char alphabet[17], double *lambda, double freqs[16]);
int rmgpu_adjusted_score(const char *matrix_file, int32_t raw, const char *qaln, const char *saln, int32_t *adjusted);

/* Host-only: the alignment block of one .cat entry, i.e. everything SearchResult::_toCrossMatchFormat
 * (SearchResult.pm:2486-2834) appends to the annotation line in mode AlignWithQuerySeq -- the 50-column chunks with
 * coordinates, the line of difference marks (i v - ?) between the rows, and the Matrix / Kimura / transitions /
 * gap statistics trailer -- from the fields a SearchResult object holds: names, rows with the query forward
 * (SearchResult::getQueryString / getSubjString), 1-based coordinates, strand 1 = orientation "C" (subject coordinates
 * descend).  matrix_name, pct_kimura, cpg_sites, pct_raw_kimura may be NULL (undef in the object; the last two are printed
 * only together).  Writes at most cap-1 bytes + NUL into out (out may be NULL) and returns the full length, or -1 when
 * the rows differ in length.  GpuSearchEngine.pm's GpuSearchResult::_toCrossMatchFormat calls it, so the driver's own
 * `$collection->write( $catfile, SearchResult::AlignWithQuerySeq )` (RepeatMasker:1489) no longer walks every column in
 * Perl.  (Added in ABI 4.) */
int64_t rmgpu_format_alignment(const char *query_name, const char *subject_name, const char *qaln, const char *saln,
                               int64_t qstart, int64_t qend, int64_t sstart, int64_t send, int strand,
                               const char *matrix_name, const char *pct_kimura, const char *cpg_sites,
                               const char *pct_raw_kimura, char *out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
