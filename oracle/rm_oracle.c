/*
 * rm_oracle.c -- plain-C CPU restatement of the crossmatch search path (see rm_oracle.h).
 * TEST INFRASTRUCTURE ONLY: the product never links or calls this file.
 *
 * The search "specification" restated here (DESIGN.md section 3 is the prose version):
 *
 *  for every query sequence, strand in {forward, reverse-complemented query}, subject sequence:
 *    exhaustive Smith-Waterman-Gotoh over the full Q x L matrix
 *        E[i][j] = max(E[i-1][j] + ins_gap_ext, H[i-1][j] + gap_init)     gap in the subject row
 *        F[i][j] = max(F[i][j-1] + del_gap_ext, H[i][j-1] + gap_init)     gap in the query row
 *        H[i][j] = max(0, E[i][j], F[i][j], H[i-1][j-1] + M[s_j][q_i])
 *    (scoring convention SearchResult.pm:1829-1852, :1894-1917).  Every cell with H>0 carries the
 *    ORIGIN (i0,j0) of its alignment: the cell where its path left a zero.  Ties are broken
 *    E over F over diagonal, and extension over opening inside E and F -- the rule that
 *    reproduces, column for column, both real cross_match alignments held by t/SearchResult.t.
 *    A hit = for each origin, the cell of maximal H (ties: smaller i, then smaller j), kept when
 *    that raw score >= minscore.
 *    The alignment is recovered by a banded global DP between origin and end cell (band schedule
 *    below), then div/del/ins percentages over the query span, the complexity-adjusted score
 *    (SearchResult.pm:2111-2150), minscore on the reported score, output order, masklevel, '*'.
 */
#include "rm_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NEG (-(1 << 29))

/* ------------------------------------------------------------------------------------------ */
/* Matrix file: grammar of Matrix.pm:308-364                                                   */
/* ------------------------------------------------------------------------------------------ */

static int is_int_token(const char *t) {
  if (!*t) return 0;
  for (; *t; t++)
    if (!(isdigit((unsigned char)*t) || *t == '-')) return 0;
  return 1;
}

int ora_matrix_parse(const char *path, ora_matrix *m, char *err, int errlen) {
  FILE *f = fopen(path, "r");
  char line[4096];
  int row = 0, have_freq = 0;
  double fh[256];
  int i;
  memset(m, 0, sizeof(*m));
  for (i = 0; i < 256; i++) fh[i] = -1.0;
  if (!f) {
    snprintf(err, errlen, "could not open matrix %s", path);
    return -1;
  }
  while (fgets(line, sizeof(line), f)) {
    char *tok[256];
    int nt = 0;
    char *p;
    for (p = line; *p; p++) *p = (char)toupper((unsigned char)*p); /* Matrix.pm:326 */
    /* FREQS A 0.325 C 0.175 ...  (Matrix.pm:330-332) */
    if ((p = strstr(line, "FREQS")) != NULL && isspace((unsigned char)p[5])) {
      char *q = p + 5;
      char *a, *b;
      have_freq = 1;
      while ((a = strtok(q, " \t\r\n")) != NULL) {
        q = NULL;
        b = strtok(NULL, " \t\r\n");
        if (!b) break;
        fh[(unsigned char)a[0]] = atof(b);
      }
      continue;
    }
    for (p = strtok(line, " \t\r\n"); p && nt < 256; p = strtok(NULL, " \t\r\n")) tok[nt++] = p;
    if (nt == 0) continue;
    /* alphabet row: >=4 single letters (Matrix.pm:337-340) */
    {
      int all_letters = nt >= 4;
      for (i = 0; i < nt && all_letters; i++)
        if (!(strlen(tok[i]) == 1 && isupper((unsigned char)tok[i][0]))) all_letters = 0;
      if (all_letters) {
        if (nt > ORA_MAXA) {
          snprintf(err, errlen, "matrix alphabet of %d symbols > %d", nt, ORA_MAXA);
          fclose(f);
          return -1;
        }
        m->n = nt;
        for (i = 0; i < nt; i++) m->alpha[i] = tok[i][0];
        m->alpha[nt] = 0;
        continue;
      }
    }
    /* value row, optional leading row letter (Matrix.pm:341-349) */
    if (m->n > 0) {
      int off = 0;
      if (strlen(tok[0]) == 1 && !isdigit((unsigned char)tok[0][0])) off = 1;
      if (nt - off >= 4 && is_int_token(tok[off]) && is_int_token(tok[off + 1]) && is_int_token(tok[off + 2]) &&
          is_int_token(tok[off + 3])) {
        if (row < ORA_MAXA)
          for (i = 0; i + off < nt && i < ORA_MAXA; i++) m->score[row][i] = atoi(tok[i + off]);
        row++;
      }
    }
  }
  fclose(f);
  if (m->n == 0 || row < m->n) {
    snprintf(err, errlen, "matrix %s: alphabet %d rows %d", path, m->n, row);
    return -1;
  }
  /* Matrix.pm:353-363: default uniform ACGT, else the FREQS hash; absent letters get 0 */
  for (i = 0; i < m->n; i++) {
    unsigned char c = (unsigned char)m->alpha[i];
    if (have_freq)
      m->freq[i] = fh[c] > 0 ? fh[c] : 0.0;
    else
      m->freq[i] = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? 0.25 : 0.0;
  }
  ora_matrix_finish(m);
  return 0;
}

/* Matrix.pm:397-464, reproduced loop for loop (doubling, then bisection to 1e-5). */
static double lambda_sum(const ora_matrix *m, double lambda) {
  double sum = 0;
  int i, j;
  for (i = 0; i < m->n; i++)
    for (j = 0; j < m->n; j++)
      if (m->freq[i] != 0 && m->freq[j] != 0) sum += m->freq[i] * m->freq[j] * exp(lambda * m->score[i][j]);
  return sum;
}

void ora_matrix_finish(ora_matrix *m) {
  double lo = 0, hi, lambda = 0.5, sum;
  int c, i;
  do {
    sum = lambda_sum(m, lambda);
    if (sum < 1.0) {
      lo = lambda;
      lambda *= 2.0;
    }
  } while (sum < 1.0);
  hi = lambda;
  while (hi - lo > .00001) {
    lambda = (lo + hi) / 2.0;
    sum = lambda_sum(m, lambda);
    if (sum >= 1.0)
      hi = lambda;
    else
      lo = lambda;
  }
  m->lambda = lambda;
  /* letter -> column.  Letters absent from the matrix (B D H V, and anything else) score as N when
     the matrix has N, else as its last column; the excision sentinel 'x' (RepeatMasker:6059) is X.
     cross_match's own treatment of these is unpinned. */
  {
    const char *pn = strchr(m->alpha, 'N');
    int fallback = pn ? (int)(pn - m->alpha) : m->n - 1;
    for (c = 0; c < 256; c++) m->code_of[c] = fallback;
    for (i = 0; i < m->n; i++) {
      m->code_of[(unsigned char)m->alpha[i]] = i;
      m->code_of[(unsigned char)tolower((unsigned char)m->alpha[i])] = i;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* FASTA                                                                                       */
/* ------------------------------------------------------------------------------------------ */

int ora_fasta_read(const char *path, ora_seqset *s, char *err, int errlen) {
  FILE *f = fopen(path, "r");
  char *line = NULL;
  size_t cap = 0;
  ssize_t n;
  int cur = -1, scap = 0;
  size_t *caps = NULL;
  memset(s, 0, sizeof(*s));
  if (!f) {
    snprintf(err, errlen, "could not open sequence file %s", path);
    return -1;
  }
  while ((n = getline(&line, &cap, f)) >= 0) {
    if (line[0] == '>') {
      char *p = line + 1, *e;
      if (s->n == scap) {
        scap = scap ? scap * 2 : 16;
        s->name = (char **)realloc(s->name, scap * sizeof(char *));
        s->seq = (char **)realloc(s->seq, scap * sizeof(char *));
        s->len = (int *)realloc(s->len, scap * sizeof(int));
        caps = (size_t *)realloc(caps, scap * sizeof(size_t));
      }
      while (*p == ' ' || *p == '\t') p++;
      e = p;
      while (*e && !isspace((unsigned char)*e)) e++;
      cur = s->n++;
      s->name[cur] = strndup(p, (size_t)(e - p));
      caps[cur] = 1024;
      s->seq[cur] = (char *)malloc(caps[cur]);
      s->len[cur] = 0;
    } else if (cur >= 0) {
      ssize_t k;
      for (k = 0; k < n; k++) {
        unsigned char c = (unsigned char)line[k];
        if (!isalpha(c)) continue;
        if ((size_t)s->len[cur] + 2 > caps[cur]) {
          caps[cur] *= 2;
          s->seq[cur] = (char *)realloc(s->seq[cur], caps[cur]);
        }
        s->seq[cur][s->len[cur]++] = (char)toupper(c); /* FastaDB.pm:1110-1118 upper-cases on load */
      }
    }
  }
  for (cur = 0; cur < s->n; cur++) s->seq[cur][s->len[cur]] = 0;
  free(line);
  free(caps);
  fclose(f);
  return 0;
}

void ora_seqset_free(ora_seqset *s) {
  int i;
  for (i = 0; i < s->n; i++) {
    free(s->name[i]);
    free(s->seq[i]);
  }
  free(s->name);
  free(s->seq);
  free(s->len);
  memset(s, 0, sizeof(*s));
}

static char comp_letter(char c) { /* same table as CrossmatchSearchEngine.pm:833 */
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    case 'Y': return 'R';
    case 'R': return 'Y';
    case 'M': return 'K';
    case 'K': return 'M';
    case 'H': return 'D';
    case 'B': return 'V';
    case 'V': return 'B';
    case 'D': return 'H';
    default: return c;
  }
}

void ora_revcomp(const char *in, int n, char *out) {
  int i;
  for (i = 0; i < n; i++) out[i] = comp_letter(in[n - 1 - i]);
  out[n] = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Score-only Smith-Waterman-Gotoh (exact prefilter: adjusted <= raw, so raw max < minscore     */
/* means no hit).                                                                              */
/* ------------------------------------------------------------------------------------------ */

int ora_sw_score(const unsigned char *q, int Q, const unsigned char *s, int L, const ora_matrix *m,
                 const ora_params *p) {
  int *Hc = (int *)malloc((size_t)(L + 1) * sizeof(int));
  int *Ec = (int *)malloc((size_t)(L + 1) * sizeof(int));
  int i, j, best = 0;
  const int gi = p->gap_init, ie = p->ins_gap_ext, de = p->del_gap_ext;
  for (j = 0; j <= L; j++) {
    Hc[j] = 0;
    Ec[j] = NEG;
  }
  for (i = 1; i <= Q; i++) {
    int qi = q[i - 1];
    int hdiag = 0, hup = 0, F = NEG;
    for (j = 1; j <= L; j++) {
      int e1 = Ec[j] + ie, e2 = Hc[j] + gi;
      int E = e1 > e2 ? e1 : e2;
      int f1 = F + de, f2 = hup + gi;
      int d = hdiag + m->score[s[j - 1]][qi];
      int h;
      F = f1 > f2 ? f1 : f2;
      h = E > F ? E : F;
      if (d > h) h = d;
      if (h < 0) h = 0;
      hdiag = Hc[j];
      Hc[j] = h;
      Ec[j] = E;
      hup = h;
      if (h > best) best = h;
    }
  }
  free(Hc);
  free(Ec);
  return best;
}

/* The same score-only pass over up to 16 subjects at once, one int16 lane per subject (inter-sequence SIMD, AVX2).
   Saturating arithmetic: a lane whose best reaches `cap` is reported as cap ("at least cap"), so the caller may only
   use the result as the boolean best >= minscore with minscore <= cap.  Lanes shorter than the longest subject are
   padded with rows that score PADSC against everything, which cannot raise a lane's best (every value in a pad row
   descends from a real cell through penalties).  ora_search() uses it when the CPU has AVX2 and RM_ORACLE_SCALAR is
   not set; tests/test_oracle_golden.py checks it against ora_sw_score(). */
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define ORA_HAVE_AVX2 1
#define ORA_LANES 16
#define ORA_CAP 32000
__attribute__((target("avx2"))) void ora_sw_score_batch(const unsigned char *q, int Q, const unsigned char *const *s,
                                                        const int *L, int nsub, const ora_matrix *m,
                                                        const ora_params *p, int *best_out) {
  const short PADSC = -100;
  int Lmax = 0, l, j, c, i;
  __m256i *prof, *Hc, *Ec, best;
  const __m256i vgi = _mm256_set1_epi16((short)p->gap_init), vie = _mm256_set1_epi16((short)p->ins_gap_ext),
                vde = _mm256_set1_epi16((short)p->del_gap_ext), vzero = _mm256_setzero_si256(),
                vcap = _mm256_set1_epi16(ORA_CAP), vneg = _mm256_set1_epi16(-32768);
  short tmp[ORA_LANES];
  for (l = 0; l < nsub; l++)
    if (L[l] > Lmax) Lmax = L[l];
  if (Lmax == 0 || Q == 0) {
    for (l = 0; l < nsub; l++) best_out[l] = 0;
    return;
  }
  prof = (__m256i *)aligned_alloc(32, (size_t)m->n * (size_t)Lmax * sizeof(__m256i));
  Hc = (__m256i *)aligned_alloc(32, (size_t)(Lmax + 1) * sizeof(__m256i));
  Ec = (__m256i *)aligned_alloc(32, (size_t)(Lmax + 1) * sizeof(__m256i));
  for (c = 0; c < m->n; c++)
    for (j = 0; j < Lmax; j++) {
      for (l = 0; l < ORA_LANES; l++) tmp[l] = (l < nsub && j < L[l]) ? (short)m->score[s[l][j]][c] : PADSC;
      prof[(size_t)c * Lmax + j] = _mm256_loadu_si256((const __m256i *)tmp);
    }
  for (j = 0; j <= Lmax; j++) {
    Hc[j] = vzero;
    Ec[j] = vneg;
  }
  best = vzero;
  for (i = 1; i <= Q; i++) {
    const __m256i *prow = prof + (size_t)q[i - 1] * Lmax;
    __m256i hdiag = vzero, hup = vzero, F = vneg;
    for (j = 1; j <= Lmax; j++) {
      const __m256i hl = Hc[j];
      const __m256i E = _mm256_max_epi16(_mm256_adds_epi16(Ec[j], vie), _mm256_adds_epi16(hl, vgi));
      const __m256i d = _mm256_adds_epi16(hdiag, prow[j - 1]);
      __m256i h;
      F = _mm256_max_epi16(_mm256_adds_epi16(F, vde), _mm256_adds_epi16(hup, vgi));
      h = _mm256_max_epi16(_mm256_max_epi16(E, F), _mm256_max_epi16(d, vzero));
      h = _mm256_min_epi16(h, vcap);
      hdiag = hl;
      Hc[j] = h;
      Ec[j] = E;
      hup = h;
      best = _mm256_max_epi16(best, h);
    }
  }
  _mm256_storeu_si256((__m256i *)tmp, best);
  for (l = 0; l < nsub; l++) best_out[l] = tmp[l];
  free(prof);
  free(Hc);
  free(Ec);
}
#endif

/* ------------------------------------------------------------------------------------------ */
/* Origin-tracking pass                                                                        */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  uint64_t origin; /* (i0 << 32) | j0, 0 = empty */
  int score, i1, j1;
} cand_t;

typedef struct {
  cand_t *tab;
  size_t cap, n;
} candmap;

static void candmap_init(candmap *c) {
  c->cap = 1024;
  c->n = 0;
  c->tab = (cand_t *)calloc(c->cap, sizeof(cand_t));
}

static cand_t *candmap_slot(candmap *c, uint64_t origin) {
  size_t h = (size_t)((origin * 0x9E3779B97F4A7C15ull) >> 20) & (c->cap - 1);
  while (c->tab[h].origin && c->tab[h].origin != origin) h = (h + 1) & (c->cap - 1);
  return &c->tab[h];
}

static void candmap_update(candmap *c, uint64_t origin, int score, int i, int j) {
  cand_t *s = candmap_slot(c, origin);
  if (!s->origin) {
    if ((c->n + 1) * 2 > c->cap) { /* grow */
      cand_t *old = c->tab;
      size_t ocap = c->cap, k;
      c->cap *= 2;
      c->tab = (cand_t *)calloc(c->cap, sizeof(cand_t));
      for (k = 0; k < ocap; k++)
        if (old[k].origin) *candmap_slot(c, old[k].origin) = old[k];
      free(old);
      s = candmap_slot(c, origin);
    }
    s->origin = origin;
    s->score = score;
    s->i1 = i;
    s->j1 = j;
    c->n++;
    return;
  }
  /* cells are visited in increasing (i, j): strict > keeps the smallest (i, j) among equals */
  if (score > s->score) {
    s->score = score;
    s->i1 = i;
    s->j1 = j;
  }
}

#define SELI(c, a, b) ((b) ^ (((a) ^ (b)) & -(int)(c)))
#define SELO(c, a, b) ((b) ^ (((a) ^ (b)) & (uint64_t) - (int64_t)(c)))

static void origin_pass(const unsigned char *q, int Q, const unsigned char *s, int L, const ora_matrix *m,
                        const ora_params *p, candmap *cm) {
  int *Hc = (int *)malloc((size_t)(L + 1) * sizeof(int));
  int *Ec = (int *)malloc((size_t)(L + 1) * sizeof(int));
  uint64_t *oH = (uint64_t *)calloc((size_t)(L + 1), sizeof(uint64_t));
  uint64_t *oE = (uint64_t *)calloc((size_t)(L + 1), sizeof(uint64_t));
  const int gi = p->gap_init, ie = p->ins_gap_ext, de = p->del_gap_ext, minscore = p->minscore;
  int i, j;
  for (j = 0; j <= L; j++) {
    Hc[j] = 0;
    Ec[j] = NEG;
  }
  for (i = 1; i <= Q; i++) {
    int mrow[ORA_MAXA]; /* scores of every subject residue against this query residue */
    const uint64_t irow = (uint64_t)(uint32_t)i << 32;
    int hdiag = 0, hup = 0, F = NEG, c;
    uint64_t odiag = 0, oup = 0, oF = 0;
    for (c = 0; c < ORA_MAXA; c++) mrow[c] = m->score[c][q[i - 1]];
    for (j = 1; j <= L; j++) {
      /* written with selects instead of branches: in unrelated sequence every one of these comparisons is a coin flip */
      const int e_ext = Ec[j] + ie, e_open = Hc[j] + gi;
      const int f_ext = F + de, f_open = hup + gi;
      /* SELI/SELO(c, a, b) = c ? a : b as mask arithmetic (gcc turns ?: back into branches here) */
      const int e_opens = e_open > e_ext; /* extension preferred on ties */
      const int f_opens = f_open > f_ext;
      const int E = SELI(e_opens, e_open, e_ext);
      const uint64_t oe = SELO(e_opens, oH[j], oE[j]);
      const int d = hdiag + mrow[s[j - 1]];
      const uint64_t od = SELO(hdiag > 0, odiag, irow | (uint32_t)j);
      int h, c1, c2, pos;
      uint64_t oh;
      F = SELI(f_opens, f_open, f_ext);
      oF = SELO(f_opens, oup, oF);
      /* priority E > F > diagonal */
      c1 = F > E;
      h = SELI(c1, F, E);
      oh = SELO(c1, oF, oe);
      c2 = d > h;
      h = SELI(c2, d, h);
      oh = SELO(c2, od, oh);
      pos = h > 0;
      h = SELI(pos, h, 0);
      oh = SELO(pos, oh, (uint64_t)0);
      hdiag = Hc[j];
      odiag = oH[j];
      Hc[j] = h;
      oH[j] = oh;
      Ec[j] = E;
      oE[j] = oe;
      hup = h;
      oup = oh;
      if (h >= minscore) candmap_update(cm, oh, h, i, j);
    }
  }
  free(Hc);
  free(Ec);
  free(oH);
  free(oE);
}

/* ------------------------------------------------------------------------------------------ */
/* Banded global traceback between origin and end cell                                         */
/*   local coordinates u=1..n (query), v=1..mlen (subject); band = diagonals d=u-v in           */
/*   [min(0,n-m)-W, max(0,n-m)+W]; W = (max(bandwidth,1) + max(n,m)/32) * 2^k, first k whose     */
/*   banded score equals the hit's raw score (or whose band covers the whole rectangle).       */
/* ------------------------------------------------------------------------------------------ */

static int banded_traceback(const unsigned char *q, int n, const unsigned char *s, int mlen, const ora_matrix *m,
                            const ora_params *p, int W, int *score_out, char *qaln, char *saln, const char *qlet,
                            const char *slet) {
  const int gi = p->gap_init, ie = p->ins_gap_ext, de = p->del_gap_ext;
  int dlo = (n - mlen < 0 ? n - mlen : 0) - W, dhi = (n - mlen > 0 ? n - mlen : 0) + W;
  int bw = dhi - dlo + 1; /* band cells per query row u: v = u - d, d in [dlo,dhi] */
  /* per-row arrays indexed by k = d - dlo */
  int *Hp = (int *)malloc((size_t)(bw + 2) * sizeof(int)), *Hn = (int *)malloc((size_t)(bw + 2) * sizeof(int));
  int *Ep = (int *)malloc((size_t)(bw + 2) * sizeof(int)), *En = (int *)malloc((size_t)(bw + 2) * sizeof(int));
  unsigned char *dir = (unsigned char *)calloc((size_t)(n + 1) * (size_t)bw, 1);
  int u, k, len = -1;
  /* row u-1 = 0: only H(0,0)=0 is reachable */
  for (k = 0; k < bw + 2; k++) {
    Hp[k] = NEG;
    Ep[k] = NEG;
  }
  if (0 >= dlo && 0 <= dhi) Hp[0 - dlo] = 0; /* cell (0,0): d = 0 */
  for (u = 1; u <= n; u++) {
    int qi = q[u - 1];
    int F = NEG, hleft_v = NEG; /* H(u, v-1) while sweeping v upward */
    /* v increases as d decreases: sweep k from high d to low d so that v goes up */
    for (k = bw - 1; k >= 0; k--) {
      int d = dlo + k, v = u - d;
      int E, D, h, e_ext, e_open, f_ext, f_open, hd;
      unsigned char bits = 0;
      if (v < 1 || v > mlen) {
        Hn[k] = NEG;
        En[k] = NEG;
        if (v < 1) { /* below the first subject column: nothing to carry */
          F = NEG;
          hleft_v = NEG;
        }
        continue;
      }
      /* E: from (u-1, v): same v, d-1 -> k-1 */
      e_ext = (k - 1 >= 0 ? Ep[k - 1] : NEG) + ie;
      e_open = (k - 1 >= 0 ? Hp[k - 1] : NEG) + gi;
      if (e_ext >= e_open) {
        E = e_ext;
        bits |= 4;
      } else
        E = e_open;
      /* F: from (u, v-1): d+1 -> k+1, computed earlier in this sweep */
      f_ext = F + de;
      f_open = hleft_v + gi;
      if (f_ext >= f_open) {
        F = f_ext;
        bits |= 8;
      } else
        F = f_open;
      /* D: from (u-1, v-1): same d -> k */
      hd = Hp[k];
      D = hd > NEG / 2 ? hd + m->score[s[v - 1]][qi] : NEG;
      if (u == 1 && v == 1) { /* forced start: the alignment opens with a diagonal at the origin */
        E = NEG;
        F = NEG;
      }
      h = E;
      bits |= 1; /* 1=E 2=F 3=D */
      if (F > h) {
        h = F;
        bits = (unsigned char)((bits & ~3) | 2);
      }
      if (D > h) {
        h = D;
        bits = (unsigned char)((bits & ~3) | 3);
      }
      if (h < NEG / 2) h = NEG;
      if (E < NEG / 2) E = NEG;
      if (F < NEG / 2) F = NEG;
      Hn[k] = h;
      En[k] = E;
      hleft_v = h;
      dir[(size_t)u * bw + k] = bits;
    }
    {
      int *t = Hp;
      Hp = Hn;
      Hn = t;
      t = Ep;
      Ep = En;
      En = t;
    }
  }
  {
    int kend = (n - mlen) - dlo;
    *score_out = Hp[kend];
    if ((dir[(size_t)n * bw + kend] & 3) == 3 && Hp[kend] > NEG / 2) {
      /* walk back */
      int uu = n, vv = mlen, state = 0, pos = 0; /* state 0=H 1=E 2=F */
      char *qa = qaln, *sa = saln;
      while (uu > 0 || vv > 0) {
        int kk;
        unsigned char b;
        if (uu < 1 || vv < 1) { /* left the rectangle anywhere but at (0,0): band too narrow */
          pos = -1;
          break;
        }
        kk = (uu - vv) - dlo;
        if (kk < 0 || kk >= bw) {
          pos = -1;
          break;
        }
        b = dir[(size_t)uu * bw + kk];
        if (state == 0) {
          int c = b & 3;
          if (c == 3) {
            qa[pos] = qlet[uu - 1];
            sa[pos] = slet[vv - 1];
            pos++;
            uu--;
            vv--;
          } else if (c == 1)
            state = 1;
          else
            state = 2;
        } else if (state == 1) {
          qa[pos] = qlet[uu - 1];
          sa[pos] = '-';
          pos++;
          state = (b & 4) ? 1 : 0;
          uu--;
        } else {
          qa[pos] = '-';
          sa[pos] = slet[vv - 1];
          pos++;
          state = (b & 8) ? 2 : 0;
          vv--;
        }
      }
      if (pos > 0) {
        int a;
        for (a = 0; a < pos / 2; a++) {
          char t = qa[a];
          qa[a] = qa[pos - 1 - a];
          qa[pos - 1 - a] = t;
          t = sa[a];
          sa[a] = sa[pos - 1 - a];
          sa[pos - 1 - a] = t;
        }
        qa[pos] = 0;
        sa[pos] = 0;
      }
      len = pos;
    }
  }
  free(Hp);
  free(Hn);
  free(Ep);
  free(En);
  free(dir);
  return len;
}

int ora_global_traceback(const unsigned char *q, int n, const unsigned char *s, int mlen, const ora_matrix *m,
                         const ora_params *p, int want_score, char *qaln, char *saln, const char *qlet,
                         const char *slet) {
  int full = n > mlen ? n : mlen;
  int W = (p->bandwidth > 0 ? p->bandwidth : 1) + full / 32; /* W0 = max(bandwidth,1) + max(n,m)/32, then doubling */
  for (;;) {
    int sc = NEG;
    int len = banded_traceback(q, n, s, mlen, m, p, W, &sc, qaln, saln, qlet, slet);
    if (len > 0 && sc == want_score) return len;
    if (W >= full) return -1;
    W *= 2;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Complexity adjustment, SearchResult.pm:2111-2150                                            */
/* ------------------------------------------------------------------------------------------ */

int ora_complexity_adjust(int raw, const long counts[ORA_MAXA], const ora_matrix *m) {
  double t_factor = 0, t_sum = 0, t_counts = 0, adj;
  int c;
  for (c = 0; c < m->n; c++) {
    if (counts[c] > 0 && m->freq[c] > 0 && log(m->freq[c]) != 0) {
      double cnt = (double)counts[c];
      t_factor += cnt * log(cnt);
      t_sum += cnt * log(m->freq[c]);
      t_counts += cnt;
    }
  }
  if (t_counts != 0) t_factor -= t_counts * log(t_counts);
  t_sum -= t_factor;
  adj = (double)raw + t_sum / m->lambda + .999;
  c = (int)adj; /* perl int(): truncation toward zero */
  return c < 0 ? 0 : c;
}

/* Raw and adjusted score of a given gapped alignment (restates SearchResult.pm:1829-1917 without the
   CpG options): used by the tests to check the scoring model against t/SearchResult.t. */
int ora_rescore(const char *qaln, const char *saln, const ora_matrix *m, const ora_params *p, int *raw_out,
                int *adj_out) {
  long counts[ORA_MAXA];
  int raw = 0, i;
  memset(counts, 0, sizeof(counts));
  for (i = 0; qaln[i] && saln[i]; i++) {
    if (saln[i] == '-')
      raw += (i > 0 && saln[i - 1] == '-') ? p->ins_gap_ext : p->gap_init;
    else if (qaln[i] == '-')
      raw += (i > 0 && qaln[i - 1] == '-') ? p->del_gap_ext : p->gap_init;
    else {
      raw += m->score[m->code_of[(unsigned char)saln[i]]][m->code_of[(unsigned char)qaln[i]]];
      counts[m->code_of[(unsigned char)qaln[i]]]++;
    }
  }
  *raw_out = raw;
  *adj_out = ora_complexity_adjust(raw, counts, m);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Hit assembly                                                                                */
/* ------------------------------------------------------------------------------------------ */

static int base_class(char c) { /* 0 purine A/G, 1 pyrimidine C/T, -1 other */
  if (c == 'A' || c == 'G') return 0;
  if (c == 'C' || c == 'T') return 1;
  return -1;
}

static void fmt_pct(char *out, long num, long den) {
  /* cross_match prints percentages with two decimals; the quotient is formed in double exactly
     like printf("%.2f", 100.0 * num / den) */
  snprintf(out, 16, "%.2f", den > 0 ? 100.0 * (double)num / (double)den : 0.0);
}

static void result_push(ora_result *r, const ora_hit *h) {
  if (r->n == r->cap) {
    r->cap = r->cap ? r->cap * 2 : 64;
    r->hit = (ora_hit *)realloc(r->hit, (size_t)r->cap * sizeof(ora_hit));
  }
  r->hit[r->n++] = *h;
}

static int cmp_cand(const void *a, const void *b) {
  const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
  if (x->origin != y->origin) return x->origin < y->origin ? -1 : 1;
  return 0;
}

static int cmp_hit(const void *a, const void *b) {
  /* output order: query file order, ascending query start, then DESCENDING subject file order
     (what the 3,882-hit fixture t/seqs/masklevel/hum-1-sine-ml101.out and seq1-2-ml101.out show),
     then ascending subject start, query end, strand */
  const ora_hit *x = (const ora_hit *)a, *y = (const ora_hit *)b;
  if (x->qidx != y->qidx) return x->qidx - y->qidx;
  if (x->qstart != y->qstart) return x->qstart - y->qstart;
  if (x->sidx != y->sidx) return y->sidx - x->sidx;
  if (x->sstart != y->sstart) return x->sstart - y->sstart;
  if (x->qend != y->qend) return x->qend - y->qend;
  if (x->strand != y->strand) return x->strand - y->strand;
  return x->send - y->send;
}

static void apply_masklevel_and_star(ora_result *r, int masklevel) {
  /* '*': a higher-scoring match overlaps this match's query domain (CrossmatchSearchEngine.pm:99-100).
     masklevel (repeatmasker.help:714-728): a match is reported only if at least (100-masklevel)% of
     the bases of its query domain are outside the union of the domains of higher-scoring reported
     matches; greedy in descending score.  masklevel >= 101 reports everything. */
  int n = r->n, a, b, kept = 0;
  int *order = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
  char *drop = (char *)calloc((size_t)(n > 0 ? n : 1), 1);
  for (a = 0; a < n; a++) order[a] = a;
  /* stable insertion sort by (qidx, score desc) keeps output order among equals */
  for (a = 1; a < n; a++) {
    int t = order[a];
    b = a - 1;
    while (b >= 0 && (r->hit[order[b]].qidx > r->hit[t].qidx ||
                      (r->hit[order[b]].qidx == r->hit[t].qidx && r->hit[order[b]].score < r->hit[t].score))) {
      order[b + 1] = order[b];
      b--;
    }
    order[b + 1] = t;
  }
  for (a = 0; a < n; a++) {
    ora_hit *h = &r->hit[order[a]];
    h->star = 0;
    for (b = 0; b < n; b++) {
      const ora_hit *o = &r->hit[b];
      if (o->qidx == h->qidx && o->score > h->score && o->qstart <= h->qend && o->qend >= h->qstart) {
        h->star = 1;
        break;
      }
    }
  }
  if (masklevel < 101) {
    for (a = 0; a < n; a++) {
      ora_hit *h = &r->hit[order[a]];
      int len = h->qend - h->qstart + 1, pos, uncovered = 0;
      char *cov = (char *)calloc((size_t)len, 1);
      for (b = 0; b < a; b++) {
        const ora_hit *o = &r->hit[order[b]];
        int lo, hi;
        if (drop[order[b]] || o->qidx != h->qidx || !(o->score > h->score)) continue;
        lo = o->qstart > h->qstart ? o->qstart : h->qstart;
        hi = o->qend < h->qend ? o->qend : h->qend;
        for (pos = lo; pos <= hi; pos++) cov[pos - h->qstart] = 1;
      }
      for (pos = 0; pos < len; pos++) uncovered += !cov[pos];
      free(cov);
      if ((long)uncovered * 100 < (long)(100 - masklevel) * len) drop[order[a]] = 1;
    }
    for (a = 0; a < n; a++) {
      if (drop[a]) {
        free(r->hit[a].qaln);
        free(r->hit[a].saln);
      } else
        r->hit[kept++] = r->hit[a];
    }
    r->n = kept;
  }
  free(order);
  free(drop);
}

int ora_search(const ora_seqset *qs, const ora_seqset *ss, const ora_matrix *m, const ora_params *p, ora_result *out) {
  int qi, si, strand, k;
  int maxL = 0;
  unsigned char **scode = (unsigned char **)malloc((size_t)ss->n * sizeof(unsigned char *));
  int use_simd = 0;
  int *bylen = (int *)malloc((size_t)(ss->n > 0 ? ss->n : 1) * sizeof(int)); /* subjects by ascending length */
  char *passes = (char *)calloc((size_t)(ss->n > 0 ? ss->n : 1), 1);
  memset(out, 0, sizeof(*out));
#ifdef ORA_HAVE_AVX2
  use_simd = __builtin_cpu_supports("avx2") && !getenv("RM_ORACLE_SCALAR") && p->minscore <= ORA_CAP &&
             p->gap_init > -16000 && p->ins_gap_ext > -16000 && p->del_gap_ext > -16000;
#endif
  for (si = 0; si < ss->n; si++) bylen[si] = si;
  {
    int a2, b2; /* insertion sort is enough for the few thousand records of a library; ties keep file order */
    for (a2 = 1; a2 < ss->n; a2++) {
      int t2 = bylen[a2];
      for (b2 = a2 - 1; b2 >= 0 && ss->len[bylen[b2]] > ss->len[t2]; b2--) bylen[b2 + 1] = bylen[b2];
      bylen[b2 + 1] = t2;
    }
  }
  for (si = 0; si < ss->n; si++) {
    scode[si] = (unsigned char *)malloc((size_t)ss->len[si] + 1);
    for (k = 0; k < ss->len[si]; k++) scode[si][k] = (unsigned char)m->code_of[(unsigned char)ss->seq[si][k]];
    if (ss->len[si] > maxL) maxL = ss->len[si];
  }
  for (qi = 0; qi < qs->n; qi++) {
    int Q = qs->len[qi];
    char *qlet = (char *)malloc((size_t)Q + 1);
    unsigned char *qcode = (unsigned char *)malloc((size_t)Q + 1);
    for (strand = 0; strand < 2; strand++) {
      if (strand == 0)
        memcpy(qlet, qs->seq[qi], (size_t)Q + 1);
      else
        ora_revcomp(qs->seq[qi], Q, qlet);
      for (k = 0; k < Q; k++) qcode[k] = (unsigned char)m->code_of[(unsigned char)qlet[k]];
#ifdef ORA_HAVE_AVX2
      if (use_simd) { /* exact prefilter, 16 subjects (of similar length) per pass */
        for (k = 0; k < ss->n; k += ORA_LANES) {
          const unsigned char *sp[ORA_LANES];
          int sl[ORA_LANES], sb[ORA_LANES], nb = ss->n - k < ORA_LANES ? ss->n - k : ORA_LANES, b;
          for (b = 0; b < nb; b++) {
            sp[b] = scode[bylen[k + b]];
            sl[b] = ss->len[bylen[k + b]];
          }
          ora_sw_score_batch(qcode, Q, sp, sl, nb, m, p, sb);
          for (b = 0; b < nb; b++) passes[bylen[k + b]] = sb[b] >= p->minscore;
        }
      }
#endif
      for (si = 0; si < ss->n; si++) {
        int L = ss->len[si];
        candmap cm;
        cand_t *cands;
        size_t nc = 0, c;
        out->cells += (double)Q * (double)L;
        if (Q == 0 || L == 0) continue;
        if (use_simd ? !passes[si] : ora_sw_score(qcode, Q, scode[si], L, m, p) < p->minscore) continue;
        candmap_init(&cm);
        origin_pass(qcode, Q, scode[si], L, m, p, &cm);
        cands = (cand_t *)malloc((cm.n + 1) * sizeof(cand_t));
        for (c = 0; c < cm.cap; c++)
          if (cm.tab[c].origin) cands[nc++] = cm.tab[c];
        free(cm.tab);
        qsort(cands, nc, sizeof(cand_t), cmp_cand);
        for (c = 0; c < nc; c++) {
          int i0 = (int)(cands[c].origin >> 32), j0 = (int)(cands[c].origin & 0xffffffffu);
          int i1 = cands[c].i1, j1 = cands[c].j1;
          int n = i1 - i0 + 1, ml = j1 - j0 + 1, len, a;
          ora_hit h;
          long counts[ORA_MAXA];
          long nsub = 0, ndel = 0, nins = 0;
          char prev_q = 0, prev_s = 0;
          memset(&h, 0, sizeof(h));
          h.qaln = (char *)malloc((size_t)(n + ml + 2));
          h.saln = (char *)malloc((size_t)(n + ml + 2));
          len = ora_global_traceback(qcode + i0 - 1, n, scode[si] + j0 - 1, ml, m, p, cands[c].score, h.qaln, h.saln,
                                     qlet + i0 - 1, ss->seq[si] + j0 - 1);
          if (len < 0) {
            fprintf(stderr, "oracle: traceback failed q%d s%d strand%d origin(%d,%d) end(%d,%d) score %d\n", qi, si,
                    strand, i0, j0, i1, j1, cands[c].score);
            free(h.qaln);
            free(h.saln);
            continue;
          }
          memset(counts, 0, sizeof(counts));
          for (a = 0; a < len; a++) {
            char cq = h.qaln[a], cs = h.saln[a];
            if (cs == '-') {
              nins++;
              h.gap_bases++;
              if (prev_s != '-') h.gap_inits++;
            } else if (cq == '-') {
              ndel++;
              h.gap_bases++;
              if (prev_q != '-') h.gap_inits++;
            } else {
              counts[m->code_of[(unsigned char)cq]]++;
              if (cq != cs) {
                int bq = base_class(cq), bs = base_class(cs);
                nsub++;
                if (bq < 0 || bs < 0)
                  h.ambiguous++;
                else if (bq == bs)
                  h.transitions++;
                else
                  h.transversions++;
              }
            }
            prev_q = cq;
            prev_s = cs;
          }
          h.raw_score = cands[c].score;
          h.score = p->raw ? h.raw_score : ora_complexity_adjust(h.raw_score, counts, m);
          if (h.score < p->minscore) {
            free(h.qaln);
            free(h.saln);
            continue;
          }
          h.qidx = qi;
          h.sidx = si;
          h.strand = strand;
          if (strand == 0) {
            h.qstart = i0;
            h.qend = i1;
          } else {
            h.qstart = Q + 1 - i1;
            h.qend = Q + 1 - i0;
          }
          h.qleft = Q - h.qend;
          h.sstart = j0;
          h.send = j1;
          h.sleft = L - j1;
          fmt_pct(h.div, nsub, n);
          fmt_pct(h.del, ndel, n);
          fmt_pct(h.ins, nins, n);
          result_push(out, &h);
        }
        free(cands);
      }
    }
    free(qlet);
    free(qcode);
  }
  out->cells *= 1.0; /* both strands already counted in the loop */
  for (si = 0; si < ss->n; si++) free(scode[si]);
  free(scode);
  free(bylen);
  free(passes);
  qsort(out->hit, (size_t)out->n, sizeof(ora_hit), cmp_hit);
  apply_masklevel_and_star(out, p->masklevel);
  return 0;
}

void ora_result_free(ora_result *r) {
  int i;
  for (i = 0; i < r->n; i++) {
    free(r->hit[i].qaln);
    free(r->hit[i].saln);
  }
  free(r->hit);
  memset(r, 0, sizeof(*r));
}

/* ------------------------------------------------------------------------------------------ */
/* cross_match text (grammar CrossmatchSearchEngine.pm:58-139, parsed by :549-879)             */
/* ------------------------------------------------------------------------------------------ */

void ora_print_hit(const ora_hit *h, const ora_seqset *q, const ora_seqset *s, int alignments, void *fp) {
  FILE *f = (FILE *)fp;
  const char *qn = q->name[h->qidx], *sn = s->name[h->sidx];
  if (h->strand == 0)
    fprintf(f, "%4d %5s %4s %4s  %s %8d %5d (%d)    %s %8d %5d (%d) %s\n", h->score, h->div, h->del, h->ins, qn,
            h->qstart, h->qend, h->qleft, sn, h->sstart, h->send, h->sleft, h->star ? "*" : "");
  else
    fprintf(f, "%4d %5s %4s %4s  %s %8d %5d (%d)  C %s   (%d) %5d %5d %s\n", h->score, h->div, h->del, h->ins, qn,
            h->qstart, h->qend, h->qleft, sn, h->sleft, h->send, h->sstart, h->star ? "*" : "");
  if (alignments && h->qaln) {
    int len = (int)strlen(h->qaln), pos = 0;
    int qpos = h->strand ? h->qend : h->qstart, spos = h->sstart;
    int qstep = h->strand ? -1 : 1;
    fprintf(f, "\n");
    while (pos < len) {
      int w = len - pos < 50 ? len - pos : 50, a, nq = 0, ns = 0;
      char marks[64];
      for (a = 0; a < w; a++) {
        char cq = h->qaln[pos + a], cs = h->saln[pos + a];
        if (cq != '-') nq++;
        if (cs != '-') ns++;
        if (cq == '-' || cs == '-')
          marks[a] = '-';
        else if (cq == cs)
          marks[a] = ' ';
        else {
          int bq = base_class(cq), bs = base_class(cs);
          marks[a] = (bq < 0 || bs < 0) ? '?' : (bq == bs ? 'i' : 'v');
        }
      }
      marks[w] = 0;
      {
        int qlast = nq ? qpos + qstep * (nq - 1) : qpos, slast = ns ? spos + (ns - 1) : spos;
        fprintf(f, "%s %-15.15s %9d %.*s %d\n", h->strand ? "C" : " ", qn, qpos, w, h->qaln + pos, qlast);
        fprintf(f, "%27s%s\n", "", marks);
        fprintf(f, "  %-15.15s %9d %.*s %d\n\n", sn, spos, w, h->saln + pos, slast);
        if (nq) qpos = qlast + qstep;
        if (ns) spos = slast + 1;
      }
      pos += w;
    }
    fprintf(f, "Transitions / transversions = %.2f (%d / %d); Gap_init rate = %.2f (%d / %d), avg. gap size = %.2f (%d / %d)\n\n",
            h->transversions ? (double)h->transitions / h->transversions : 1.0, h->transitions, h->transversions,
            (double)h->gap_inits / (double)(h->qend - h->qstart + 1), h->gap_inits, h->qend - h->qstart + 1,
            h->gap_inits ? (double)h->gap_bases / h->gap_inits : 0.0, h->gap_bases, h->gap_inits);
  }
}
