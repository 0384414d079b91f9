/*
 * rm_oracle.h -- CPU restatement ("oracle") of the crossmatch search path of RepeatMasker.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (repeatmasker_b200/, include/) includes,
 * links or executes anything under oracle/.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker / CPU arm.
 *
 * What it restates.  RepeatMasker's `-engine crossmatch` path is
 *   CrossmatchSearchEngine.pm:381-524 (search) -> popen of the external `cross_match` binary
 *   (CrossmatchSearchEngine.pm:444) -> CrossmatchSearchEngine.pm:549-879 (parseOutput).
 * The arithmetic lives in Phil Green's cross_match (phrap/swat distribution; no pinned version, the
 * reference accepts any `cross_match version [\.\d]+`, CrossmatchSearchEngine.pm:251-254); it is
 * NOT part of /root/reference and is not installed, so it cannot be compiled as oracle/_ref.
 * This file restates the published algorithm at the reference's own call sites:
 *   - scoring model: matrix[subject][query] (SearchResult.pm:1894-1917), first gap base costs
 *     gap_init, later ones ins_gap_ext (gap in subject row) / del_gap_ext (gap in query row)
 *     (SearchResult.pm:1829-1852);
 *   - complexity-adjusted score (SearchResult.pm:2111-2150) with lambda from Matrix.pm:397-464;
 *   - matrix file grammar Matrix.pm:308-364;
 *   - hit line / alignment block grammar CrossmatchSearchEngine.pm:58-139 and :549-879;
 *   - masklevel semantics repeatmasker.help:714-728 + fixtures t/seqs/masklevel/seq1-2-ml{97,98,101}.out.
 *
 * PARITY PINNING (see DESIGN.md "Oracle"): pinned against the reference's own known answers for
 *   raw score / adjusted score / alignment strings (t/SearchResult.t: 353->253, 2331->1530, both
 *   alignments reproduced column for column by the traceback tie-break used here), lambda for all
 *   46 matrices (Matrix.pm run under perl), the 12-hit search KAT seq1-2-ml101.out and its
 *   masklevel 98/97 variants.  Hit DISCOVERY is exhaustive Smith-Waterman-Gotoh, not cross_match's
 *   word-seeded bands: which sub-optimal alignments are enumerated is defined here (one hit per
 *   alignment origin cell), and is "parity unpinned" beyond that KAT.
 */
#ifndef RM_ORACLE_H
#define RM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAXA 16 /* matrix alphabets of the crossmatch path have 11-12 symbols */

typedef struct {
  int n;                    /* alphabet size */
  char alpha[ORA_MAXA + 1]; /* column/row letters, upper case */
  int score[ORA_MAXA][ORA_MAXA]; /* [subject][query] */
  double freq[ORA_MAXA];
  double lambda;
  int code_of[256]; /* letter -> column (after the fallback rule, see ora_matrix_finish) */
} ora_matrix;

typedef struct {
  int n;
  char **name; /* first token of the header */
  char **seq;  /* upper-cased letters */
  int *len;
} ora_seqset;

typedef struct {
  int gap_init, ins_gap_ext, del_gap_ext; /* negative */
  int minscore;
  int masklevel; /* 0..101 */
  int minmatch;  /* accepted, unused: discovery is exhaustive */
  int bandwidth; /* seeds the traceback band schedule */
  int raw;       /* 1: report raw scores (cross_match -raw) */
  int alignments;
} ora_params;

typedef struct {
  int score, raw_score;
  int qidx, sidx, strand; /* strand 1 = "C" */
  int qstart, qend, qleft; /* 1-based, forward-strand query coordinates */
  int sstart, send, sleft;
  char div[16], del[16], ins[16]; /* the 2-decimal strings cross_match prints */
  int transitions, transversions, ambiguous, gap_inits, gap_bases;
  int star;
  char *qaln, *saln; /* as DISPLAYED by cross_match: for strand C the query row is the reverse
                        complement (descending coordinates), the subject row is forward */
} ora_hit;

typedef struct {
  int n, cap;
  ora_hit *hit;
  double cells; /* DP cells of the exhaustive pass: sum Q*L*2 */
} ora_result;

int ora_matrix_parse(const char *path, ora_matrix *m, char *err, int errlen);
void ora_matrix_finish(ora_matrix *m); /* freqs default, lambda, letter map */
int ora_fasta_read(const char *path, ora_seqset *s, char *err, int errlen);
void ora_seqset_free(ora_seqset *s);
void ora_revcomp(const char *in, int n, char *out);

int ora_search(const ora_seqset *q, const ora_seqset *s, const ora_matrix *m, const ora_params *p, ora_result *out);
void ora_result_free(ora_result *r);

/* pieces exposed for unit tests */
int ora_sw_score(const unsigned char *q, int Q, const unsigned char *s, int L, const ora_matrix *m,
                 const ora_params *p); /* best local raw score */
/* the same for up to 16 subjects at once (AVX2, x86-64 only; saturates at 32000): the prefilter ora_search() uses */
void ora_sw_score_batch(const unsigned char *q, int Q, const unsigned char *const *s, const int *L, int nsub,
                        const ora_matrix *m, const ora_params *p, int *best_out);
int ora_complexity_adjust(int raw, const long counts[ORA_MAXA], const ora_matrix *m);
int ora_global_traceback(const unsigned char *q, int n, const unsigned char *s, int mlen, const ora_matrix *m,
                         const ora_params *p, int want_score, char *qaln, char *saln,
                         const char *qlet, const char *slet); /* returns alignment length or -1 */
int ora_rescore(const char *qaln, const char *saln, const ora_matrix *m, const ora_params *p, int *raw_out,
                int *adj_out);
void ora_print_hit(const ora_hit *h, const ora_seqset *q, const ora_seqset *s, int alignments, void *FILE_out);

#ifdef __cplusplus
}
#endif
#endif
