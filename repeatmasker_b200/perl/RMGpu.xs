/*
 * RMGpu.xs -- XS glue between GpuSearchEngine.pm and the C ABI of librmgpu.so (include/rmgpu.h).
 * One call, RMGpu::search(...), stands where CrossmatchSearchEngine.pm:444 popen()s cross_match and
 * :549-879 parses its text: it returns (status, error text, array ref of hit hashes).
 */
#define PERL_NO_GET_CONTEXT
#include "EXTERN.h"
#include "perl.h"
#include "XSUB.h"

#include "rmgpu.h"

MODULE = RMGpu    PACKAGE = RMGpu

PROTOTYPES: DISABLE

const char *
version()
  CODE:
    RETVAL = rmgpu_version();
  OUTPUT:
    RETVAL

int
server_ping(socket_path)
    const char *socket_path
  CODE:
    RETVAL = rmgpu_server_ping(socket_path);
  OUTPUT:
    RETVAL

int
server_quit(socket_path)
    const char *socket_path
  CODE:
    RETVAL = rmgpu_server_quit(socket_path);
  OUTPUT:
    RETVAL

SV *
format_alignment(query_name, subject_name, qaln, saln, qstart, qend, sstart, send, strand, matrix_name, pct_kimura, cpg_sites, pct_raw_kimura)
    const char *query_name
    const char *subject_name
    const char *qaln
    const char *saln
    IV qstart
    IV qend
    IV sstart
    IV send
    int strand
    SV *matrix_name
    SV *pct_kimura
    SV *cpg_sites
    SV *pct_raw_kimura
  PREINIT:
    const char *mn, *pk, *cs, *rk;
    int64_t cap, need;
    SV *out;
  CODE:
    /* undef stays undef (NULL): the reference prints those lines only for defined fields (SearchResult.pm:2786-2800) */
    mn = SvOK(matrix_name) ? SvPV_nolen(matrix_name) : NULL;
    pk = SvOK(pct_kimura) ? SvPV_nolen(pct_kimura) : NULL;
    cs = SvOK(cpg_sites) ? SvPV_nolen(cpg_sites) : NULL;
    rk = SvOK(pct_raw_kimura) ? SvPV_nolen(pct_raw_kimura) : NULL;
    cap = (int64_t)strlen(qaln) * 6 + 1024 + (mn ? (int64_t)strlen(mn) : 0);
    out = newSV((STRLEN)cap);
    need = rmgpu_format_alignment(query_name, subject_name, qaln, saln, (int64_t)qstart, (int64_t)qend, (int64_t)sstart,
                                  (int64_t)send, strand, mn, pk, cs, rk, SvPVX(out), cap);
    if (need >= cap) { /* cannot happen with the bound above unless the optional texts are huge: size exactly */
      SvGROW(out, (STRLEN)need + 1);
      need = rmgpu_format_alignment(query_name, subject_name, qaln, saln, (int64_t)qstart, (int64_t)qend, (int64_t)sstart,
                                    (int64_t)send, strand, mn, pk, cs, rk, SvPVX(out), need + 1);
    }
    if (need < 0) {
      SvREFCNT_dec(out);
      XSRETURN_UNDEF;
    }
    SvPOK_only(out);
    SvCUR_set(out, (STRLEN)need);
    RETVAL = out;
  OUTPUT:
    RETVAL

void
search(query, subject, matrix, gap_init, ins_gap_ext, del_gap_ext, minscore, masklevel, minmatch, bandwidth, raw, word_raw, alignments, device)
    const char *query
    const char *subject
    const char *matrix
    int gap_init
    int ins_gap_ext
    int del_gap_ext
    int minscore
    int masklevel
    int minmatch
    int bandwidth
    int raw
    int word_raw
    int alignments
    int device
  PREINIT:
    rmgpu_params p;
    rmgpu_result *res = NULL;
    int rc, i;
    AV *hits;
  PPCODE:
    memset(&p, 0, sizeof(p));
    p.gap_init = gap_init;
    p.ins_gap_ext = ins_gap_ext;
    p.del_gap_ext = del_gap_ext;
    p.minscore = minscore;
    p.masklevel = masklevel;
    p.minmatch = minmatch;
    p.bandwidth = bandwidth;
    p.raw = raw;
    p.word_raw = word_raw;
    p.want_alignments = alignments;
    p.device = device;
    rc = rmgpu_search(&p, query, subject, matrix, &res);
    hits = newAV();
    if (rc == 0 && res) {
      for (i = 0; i < res->n_hits; i++) {
        const rmgpu_hit *h = &res->hits[i];
        HV *hv = newHV();
        (void)hv_stores(hv, "score", newSViv(h->score));
        (void)hv_stores(hv, "rawScore", newSViv(h->raw_score));
        (void)hv_stores(hv, "pctDiverge", newSVpv(h->pct_div, 0));
        (void)hv_stores(hv, "pctDelete", newSVpv(h->pct_del, 0));
        (void)hv_stores(hv, "pctInsert", newSVpv(h->pct_ins, 0));
        (void)hv_stores(hv, "queryName", newSVpv(h->query_name, 0));
        (void)hv_stores(hv, "queryStart", newSViv(h->qstart));
        (void)hv_stores(hv, "queryEnd", newSViv(h->qend));
        (void)hv_stores(hv, "queryRemaining", newSViv(h->qleft));
        (void)hv_stores(hv, "orientation", newSVpv(h->strand ? "C" : "", 0));
        (void)hv_stores(hv, "subjName", newSVpv(h->subject_name, 0));
        (void)hv_stores(hv, "subjStart", newSViv(h->sstart));
        (void)hv_stores(hv, "subjEnd", newSViv(h->send));
        (void)hv_stores(hv, "subjRemaining", newSViv(h->sleft));
        (void)hv_stores(hv, "star", newSViv(h->overlap_star));
        (void)hv_stores(hv, "transitions", newSViv(h->transitions));
        (void)hv_stores(hv, "transversions", newSViv(h->transversions));
        if (h->pct_kimura[0]) (void)hv_stores(hv, "pctKimuraDiverge", newSVpv(h->pct_kimura, 0));
        if (h->qaln) (void)hv_stores(hv, "queryString", newSVpv(h->qaln, 0));
        if (h->saln) (void)hv_stores(hv, "subjString", newSVpv(h->saln, 0));
        av_push(hits, newRV_noinc((SV *)hv));
      }
    }
    EXTEND(SP, 3);
    PUSHs(sv_2mortal(newSViv(rc)));
    PUSHs(sv_2mortal(newSVpv(rc ? rmgpu_last_error() : "", 0)));
    PUSHs(sv_2mortal(newRV_noinc((SV *)hits)));
    if (res) rmgpu_free(res);
