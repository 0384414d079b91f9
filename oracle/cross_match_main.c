/*
 * cross_match_main.c -- the oracle packaged as a `cross_match`-compatible executable.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY ("restated crossmatch (this repo)"; never Phil Green's
 * cross_match).  It accepts the argv that CrossmatchSearchEngine.pm:277-363 builds and prints the
 * grammar CrossmatchSearchEngine.pm:549-879 parses, so the UNMODIFIED reference driver can run it:
 *   RepeatMasker -engine crossmatch  with CROSSMATCH_DIR pointing at oracle/_build/.
 * Run with no arguments it prints a banner holding "cross_match version <v>"
 * (CrossmatchSearchEngine.pm:251-254 scrapes it).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#include "rm_oracle.h"

#define ORACLE_VERSION "0.1.0"

static int looks_numeric(const char *s) {
  if (*s == '-' || *s == '+') s++;
  if (!*s) return 0;
  for (; *s; s++)
    if (!isdigit((unsigned char)*s) && *s != '.') return 0;
  return 1;
}

int main(int argc, char **argv) {
  ora_params p;
  ora_matrix m;
  ora_seqset q, s;
  ora_result r;
  const char *files[2] = {NULL, NULL}, *matrix = NULL;
  char err[512];
  int nf = 0, a, have_ins = 0, have_del = 0, gap_ext = -3;
  memset(&p, 0, sizeof(p));
  p.gap_init = -4; /* cross_match defaults: penalty -2 => gap_init penalty-2, gap_ext penalty-1 */
  p.minscore = 30;
  p.masklevel = 80;
  p.minmatch = 14;
  p.bandwidth = 14;
  if (argc < 2) {
    printf("cross_match version %s (rmgpu oracle: restated crossmatch, exhaustive Smith-Waterman-Gotoh)\n",
           ORACLE_VERSION);
    printf("usage: cross_match query.fa subject.fa [-matrix M] [-gap_init n] [-ins_gap_ext n] [-del_gap_ext n]\n"
           "       [-minscore n] [-masklevel n] [-minmatch n] [-bandwidth n] [-raw] [-word_raw] [-alignments]\n");
    return 0;
  }
  for (a = 1; a < argc; a++) {
    const char *o = argv[a];
    if (o[0] == '-' && !looks_numeric(o)) {
      int has_val = a + 1 < argc && looks_numeric(argv[a + 1]);
      if (!strcmp(o, "-raw")) p.raw = 1;
      else if (!strcmp(o, "-word_raw")) ;
      else if (!strcmp(o, "-alignments")) p.alignments = 1;
      else if (!strcmp(o, "-matrix") && a + 1 < argc) matrix = argv[++a];
      else if (!strcmp(o, "-gap_init") && has_val) p.gap_init = atoi(argv[++a]);
      else if (!strcmp(o, "-gap_ext") && has_val) gap_ext = atoi(argv[++a]);
      else if (!strcmp(o, "-ins_gap_ext") && has_val) { p.ins_gap_ext = atoi(argv[++a]); have_ins = 1; }
      else if (!strcmp(o, "-del_gap_ext") && has_val) { p.del_gap_ext = atoi(argv[++a]); have_del = 1; }
      else if (!strcmp(o, "-minscore") && has_val) p.minscore = atoi(argv[++a]);
      else if (!strcmp(o, "-masklevel") && has_val) p.masklevel = atoi(argv[++a]);
      else if (!strcmp(o, "-minmatch") && has_val) p.minmatch = atoi(argv[++a]);
      else if (!strcmp(o, "-bandwidth") && has_val) p.bandwidth = atoi(argv[++a]);
      else if (has_val) a++; /* unknown option with a value: accepted and ignored */
    } else if (nf < 2)
      files[nf++] = o;
  }
  if (!have_ins) p.ins_gap_ext = gap_ext;
  if (!have_del) p.del_gap_ext = gap_ext;
  if (nf < 2) {
    fprintf(stderr, "cross_match (rmgpu oracle): need a query and a subject file\n");
    return 1;
  }
  if (matrix) {
    if (ora_matrix_parse(matrix, &m, err, sizeof(err))) {
      fprintf(stderr, "%s\n", err);
      return 1;
    }
  } else { /* cross_match default: +1 match, -2 mismatch over ACGTN */
    int i, j;
    memset(&m, 0, sizeof(m));
    m.n = 5;
    strcpy(m.alpha, "ACGTN");
    for (i = 0; i < 5; i++)
      for (j = 0; j < 5; j++) m.score[i][j] = (i == 4 || j == 4) ? 0 : (i == j ? 1 : -2);
    for (i = 0; i < 4; i++) m.freq[i] = 0.25;
    ora_matrix_finish(&m);
  }
  if (ora_fasta_read(files[0], &q, err, sizeof(err)) || ora_fasta_read(files[1], &s, err, sizeof(err))) {
    fprintf(stderr, "%s\n", err);
    return 1;
  }
  for (a = 0; a < argc; a++) printf("%s ", a ? argv[a] : "cross_match");
  printf("\ncross_match version %s (rmgpu oracle: restated crossmatch, exhaustive Smith-Waterman-Gotoh)\n\n", ORACLE_VERSION);
  printf("Query file(s):  %s\nSubject file(s):   %s\n", files[0], files[1]);
  printf("Pairwise comparison algorithm: exhaustive Smith-Waterman-Gotoh, one match per alignment origin\n\n");
  printf("Gap penalties: gap_init: %d, ins_gap_ext: %d, del_gap_ext: %d, \n", p.gap_init, p.ins_gap_ext, p.del_gap_ext);
  printf("%s lambda: %.15g\n", p.raw ? "Using raw scores." : "Using complexity-adjusted scores.", m.lambda);
  printf("minmatch: %d (unused), minscore: %d, bandwidth: %d, masklevel: %d\n\n", p.minmatch, p.minscore, p.bandwidth,
         p.masklevel);
  printf("Maximal single base matches (low complexity regions):\n\n");
  ora_search(&q, &s, &m, &p, &r);
  for (a = 0; a < r.n; a++) ora_print_hit(&r.hit[a], &q, &s, p.alignments, stdout);
  {
    int nq = 0, last = -1;
    for (a = 0; a < r.n; a++)
      if (r.hit[a].qidx != last) {
        nq++;
        last = r.hit[a].qidx;
      }
    printf("\n%d matching entries (first file).\n", nq);
  }
  fprintf(stderr, "cells %.0f\n", r.cells);
  ora_result_free(&r);
  ora_seqset_free(&q);
  ora_seqset_free(&s);
  return 0;
}
