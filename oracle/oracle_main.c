/*
 * ORACLE / TEST INFRASTRUCTURE ONLY.  Command line of the plain-C restatement:
 *   crumble_oracle [crumble options] in out
 * Same option string as crumble (parsed by the product's host option parser, which the
 * test-suite checks against the verbatim reference's -v report).  ORACLE_COLUMN_DUMP=path
 * writes one binary oracle_column per counted column.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "htslib/sam.h"
#include "crumble_host.h"
#include "crumble_oracle.h"

static void dump_cb(void *d, const oracle_column *c) { fwrite(c, sizeof(*c), 1, (FILE *)d); }

int main(int argc, char **argv) {
    crumble_opts o; htsFormat ifmt, ofmt; int oi = 1;
    memset(&ifmt, 0, sizeof ifmt); memset(&ofmt, 0, sizeof ofmt);
    crumble_opts_default(&o);
    int r = crumble_parse_args(&o, argc, argv, &ifmt, &ofmt, &oi);
    if (r) { crumble_usage(r == 2 ? stdout : stderr); return r == 2 ? 0 : 1; }
    if (o.p.verbose) crumble_print_params(&o);
    if (oi + 1 >= argc) { fprintf(stderr, "usage: crumble_oracle [options] in out\n"); return 1; }
    const char *fnin = argv[oi], *fnout = argv[oi + 1];
    if (o.p.softclip || o.p.perfect_col || o.bed_fn) { fprintf(stderr, "crumble_oracle: option not restated\n"); return 1; }
    for (int i = 0; i < 256; i++) if (o.p.preserve_qual[i]) { fprintf(stderr, "crumble_oracle: -k/-K not restated\n"); return 1; }
    samFile *in = sam_open_format(fnin, "r", &ifmt);
    if (!in) { perror(fnin); return 1; }
    char mode[8] = "w"; sam_open_mode(mode + 1, fnout, NULL);
    samFile *out = sam_open_format(fnout, mode, &ofmt);
    if (!out) { perror(fnout); return 1; }
    bam_hdr_t *h = sam_hdr_read(in);
    if (!h) return 1;
    if (sam_hdr_write(out, h) != 0) return 1;
    hts_itr_t *it = NULL;
    oracle_params P; memset(&P, 0, sizeof P);
    const cg_params *p = &o.p;
    P.reduce_qual = p->reduce_qual; P.binary_qual = p->binary_qual; P.iSTR_add = p->iSTR_add; P.sSTR_add = p->sSTR_add;
    P.iSTR_mul = p->iSTR_mul; P.sSTR_mul = p->sSTR_mul; P.qlow = p->qlow; P.qcutoff = p->qcutoff; P.qhigh = p->qhigh; P.qcap = p->qcap;
    P.min_mqual = p->min_mqual; P.indel_fract = p->indel_fract; P.min_qual_A = p->min_qual_A; P.min_indel_A = p->min_indel_A;
    P.min_discrep_A = p->min_discrep_A; P.min_qual_B = p->min_qual_B; P.min_indel_B = p->min_indel_B; P.min_discrep_B = p->min_discrep_B;
    P.low_mqual_perc = p->low_mqual_perc; P.clip_perc = p->clip_perc; P.ins_len_perc = p->ins_len_perc; P.over_depth = p->over_depth;
    P.indel_ov_perc = p->indel_ov_perc; P.pblock = p->pblock; P.region_tid = -1; P.region_end = 0x7fffffff;
    if (o.region) {
        it = sam_itr_querys(NULL, h, o.region);
        if (!it) { fprintf(stderr, "bad region\n"); return 1; }
        P.region_tid = it->tid; P.region_beg = it->beg; P.region_end = it->end;
    }
    size_t n = 0, cap = 0; bam1_t **rv = NULL; orec *recs = NULL;
    bam1_t *b = bam_init1();
    for (;;) {
        int rr = it ? sam_itr_next(in, it, b) : sam_read1(in, h, b);
        if (rr < 0) break;
        if (n == cap) { cap = cap ? cap * 2 : 4096; rv = (bam1_t **)realloc(rv, cap * sizeof(*rv)); recs = (orec *)realloc(recs, cap * sizeof(*recs)); }
        rv[n] = bam_dup1(b);
        orec *q = &recs[n]; memset(q, 0, sizeof(*q));
        q->tid = rv[n]->core.tid; q->pos = rv[n]->core.pos; q->flag = rv[n]->core.flag; q->mapq = rv[n]->core.qual;
        q->l_qseq = rv[n]->core.l_qseq; q->n_cigar = (int)rv[n]->core.n_cigar;
        q->cigar = bam_get_cigar(rv[n]); q->seq = bam_get_seq(rv[n]); q->q_out = bam_get_qual(rv[n]);
        q->q_in = (uint8_t *)malloc((size_t)q->l_qseq + 1);
        n++;
    }
    bam_destroy1(b);
    FILE *dump = getenv("ORACLE_COLUMN_DUMP") ? fopen(getenv("ORACLE_COLUMN_DUMP"), "wb") : NULL;
    long long cnt[ORACLE_N_COUNTERS];
    hts_lite_io_span_reset();
    int rc = oracle_transcode(&P, recs, (long)n, o.bed_fp, h->target_name, cnt, dump ? dump_cb : NULL, dump);
    if (dump) fclose(dump);
    if (rc) { fprintf(stderr, "Error while reducing file (%d)\n", rc); return 1; }
    for (size_t i = 0; i < n; i++) { crumble_purge_tags(&o, rv[i]); if (sam_write1(out, h, rv[i]) < 0) return 1; }
    for (int i = 0; i < ORACLE_N_COUNTERS; i++) o.counters[i] = cnt[i];
    sam_close(in); sam_close(out);
    if (o.p.verbose) crumble_print_counters(&o);
    if (o.bed_fp) fclose(o.bed_fp);
    return 0;
}
