/*
 * transcode_gpu.c — host driver with the reference's transcode() contract
 * (snp_score.c:1336-2029): read records, hand them to the device path through the C ABI
 * of include/crumble_gpu.h, write them back in input order with rewritten qualities.
 *
 * What stays on the host, as in the reference: record I/O, the BED text lines
 * (snp_score.c:1496-1498 etc.), the -v counters, the read-count check (2021-2026).
 * Everything between "record decoded" and "record ready to write" runs on the GPU.
 *
 * The stream is processed in bounded memory as a CHAIN of device calls (cg_process_window): the host cuts the
 * coordinate-sorted input every CRUMBLE_BATCH_READS records (default 2 Mi), keeps the reads that straddle a cut
 * as the next call's halo, and writes records out in input order as they become final.  The reference does the same
 * thing with its streaming loop and the two RB-trees of in-flight reads (snp_score.c:1113-1153, 1926-1975); here the
 * unit is a region shard instead of a column.  With -r the region is one call (cg_process).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <inttypes.h>
#include <limits.h>
#include "htslib/sam.h"
#include "crumble_host.h"

/* one record in flight: read but not yet written, or written but still part of a later call's halo */
typedef struct {
    bam1_t  *b;             /* NULL once written and destroyed (then oq/meta below serve the halo) */
    uint8_t *oq;            /* original qualities, saved when the record turns final while later calls still need it */
    int32_t  end;           /* pos + reference span (exclusive) for pileup records */
    uint8_t  in_pileup, final, is_new;
    int64_t  slot;          /* index in the current batch, -1 when not part of it */
} live_rec;

typedef struct { live_rec *v; size_t n, cap, head; } live_q;

static int lq_push(live_q *q, bam1_t *b) {
    if (q->n == q->cap) {
        if (q->head > 0) {                                  /* drop the written prefix */
            memmove(q->v, q->v + q->head, (q->n - q->head) * sizeof(*q->v));
            q->n -= q->head; q->head = 0;
        }
        if (q->n == q->cap) {
            size_t nc = q->cap ? q->cap * 2 : 4096;
            live_rec *nv = (live_rec *)realloc(q->v, nc * sizeof(*nv));
            if (!nv) return -1;
            q->v = nv; q->cap = nc;
        }
    }
    live_rec *r = &q->v[q->n++];
    memset(r, 0, sizeof(*r));
    r->b = b; r->is_new = 1; r->slot = -1;
    /* pileup eligibility and reference span exactly as the device decides them (cg_prep_read; snp_score.c:1125-1149) */
    int span = 0, hasref = 0;
    const uint32_t *cig = bam_get_cigar(b);
    for (uint32_t k = 0; k < b->core.n_cigar; k++)
        if (bam_cigar_type(bam_cigar_op(cig[k])) & 2) { span += (int)bam_cigar_oplen(cig[k]); hasref = 1; }
    r->in_pileup = b->core.tid >= 0 && !(b->core.flag & BAM_FUNMAP) && hasref;
    if (r->in_pileup && span == 0) span = 1;
    r->end = b->core.pos + (r->in_pileup ? span : 0);
    return 0;
}

static void lq_free(live_q *q) {
    for (size_t i = q->head; i < q->n; i++) { if (q->v[i].b) bam_destroy1(q->v[i].b); free(q->v[i].oq); }
    free(q->v);
}

static void write_bed(crumble_opts *o, bam_hdr_t *header, const cg_result *res) {
    /* BED lines, in column order (snp_score.c:1496-1498,1676-1678,1768-1770,1802-1804,1810-1812) */
    static const char *tag[5] = { "VDEEP", "DEEP", "CLIP", "INDEL_LEN", "INDEL_COVERAGE" };
    if (!o->bed_fp) return;
    for (int64_t i = 0; i < res->n_events; i++) {
        const cg_bed_event *e = &res->events[i];
        int s = e->pos - 50; if (s < 0) s = 0;
        fprintf(o->bed_fp, "%s\t%d\t%d\t%s\n", header->target_name[e->tid], s, e->pos + 50, tag[e->tag]);
    }
}

static int grow_result(cg_result *res, int64_t *qcap, int64_t qual_bytes) {
    if (qual_bytes + 16 > *qcap) {
        int64_t nc = qual_bytes + (qual_bytes >> 2) + 4096;
        uint8_t *nq = (uint8_t *)realloc(res->qual_out, (size_t)nc);
        if (!nq) return -1;
        res->qual_out = nq; *qcap = nc;
    }
    return 0;
}

int transcode_gpu(crumble_opts *o, samFile *in, samFile *out, bam_hdr_t *header, hts_itr_t *h_iter) {
    int ret = -1, err = 0;
    live_q lq = {0};
    cg_batch_builder *bb = NULL;
    cg_ctx *ctx = NULL;
    cg_result res; memset(&res, 0, sizeof(res));
    int64_t qcap = 0;
    int64_t count_in = 0, count_out = 0;
    bam1_t *nx = NULL;                                      /* the record after the current cut */
    float device_ms = 0;

    cg_params p = o->p;
    if (h_iter) { p.region_tid = h_iter->tid; p.region_beg = h_iter->beg; p.region_end = h_iter->end; }
    int64_t batch_reads = 2 << 20;
    const char *ev = getenv("CRUMBLE_BATCH_READS");
    if (ev && atoll(ev) > 0) batch_reads = atoll(ev);
    if (h_iter) batch_reads = INT64_MAX;                    /* a -r region is one call: the region logic owns the column limits */

    if (!(bb = cgb_create(1))) goto done;
    ctx = cg_create(&p, o->device, &err);
    if (!ctx) { fprintf(stderr, "crumble: cannot create GPU context: %s\n", cg_strerror(err)); goto done; }
    res.events_cap = 1 << 16;
    res.events = (cg_bed_event *)malloc(sizeof(cg_bed_event) * (size_t)res.events_cap);
    if (!res.events) goto done;

    int eof = 0, first = 1;
    int32_t lo_tid = -1, lo_pos = 0, cnt_pos = 0;
    int64_t last_key = INT64_MIN; int seen_unplaced = 0;
    while (!eof || nx || lq.head < lq.n) {
        /* ---- gather the new records of this call ---- */
        int64_t n_new = 0;
        int32_t last_tid = -2;
        for (;;) {
            if (!nx && !eof) {
                nx = bam_init1();
                int r = h_iter ? sam_itr_next(in, h_iter, nx) : sam_read1(in, header, nx);
                if (r < -1) { fprintf(stderr, "Error reading input\n"); goto done; }
                if (r < 0) { eof = 1; bam_destroy1(nx); nx = NULL; }
                else {
                    count_in++;
                    if (nx->core.tid >= 0 && !(nx->core.flag & BAM_FUNMAP)) {       /* sortedness across calls (cgb_add checks inside one) */
                        int64_t key = ((int64_t)nx->core.tid << 32) | (uint32_t)nx->core.pos;
                        if (key < last_key || seen_unplaced) { fprintf(stderr, "crumble: %s\n", cg_strerror(CG_ERR_UNSORTED)); goto done; }
                        last_key = key;
                    } else if (nx->core.tid < 0) seen_unplaced = 1;
                }
            }
            if (!nx) break;
            if (n_new >= batch_reads) break;
            if (lq_push(&lq, nx) < 0) goto done;
            last_tid = nx->core.tid;
            nx = NULL; n_new++;
        }
        if (n_new == 0) break;

        /* ---- this call's column window ---- */
        cg_window win; memset(&win, 0, sizeof win);
        win.first = first; win.lo_tid = lo_tid; win.lo_pos = lo_pos; win.cnt_pos = cnt_pos; win.hi_tid = -1;
        if (nx && nx->core.tid >= 0 && nx->core.tid == last_tid) { win.hi_tid = nx->core.tid; win.hi_pos = nx->core.pos; }
        /* ---- the batch: halo (earlier records of lo_tid that reach beyond lo_pos) + new records, in input order ---- */
        cgb_reset(bb);
        int64_t nb = 0;
        int32_t S = win.hi_pos;
        for (size_t i = lq.head; i < lq.n; i++) {
            live_rec *r = &lq.v[i];
            r->slot = -1;
            if (!r->is_new && !(r->in_pileup && !first && r->b->core.tid == lo_tid && r->end > lo_pos)) continue;
            const bam1_t *b = r->b;
            int e = cgb_add(bb, b->core.tid, b->core.pos, b->core.flag, b->core.qual, b->core.l_qseq, b->core.n_cigar,
                            bam_get_cigar(b), bam_get_seq(b), r->oq ? r->oq : bam_get_qual(b));
            if (e) { fprintf(stderr, "crumble: %s\n", cg_strerror(e)); goto done; }
            r->slot = nb++;
            if (win.hi_tid >= 0 && r->in_pileup && b->core.tid == win.hi_tid && r->end > win.hi_pos && b->core.pos < S) S = b->core.pos;
        }
        win.next_lo_pos = S;
        cg_batch batch;
        if ((err = cgb_finish(bb, &batch)) != 0) { fprintf(stderr, "crumble: %s\n", cg_strerror(err)); goto done; }
        if (grow_result(&res, &qcap, batch.qual_bytes) < 0) goto done;

        /* ---- device ---- */
        if (h_iter) {
            for (;;) {
                err = cg_process(ctx, &batch, &res);
                if (err == CG_OK && res.n_events > res.events_cap) {            /* event buffer too small: grow and redo */
                    res.events_cap = res.n_events;
                    cg_bed_event *ne = (cg_bed_event *)realloc(res.events, sizeof(cg_bed_event) * (size_t)res.events_cap);
                    if (!ne) goto done;
                    res.events = ne;
                    continue;
                }
                break;
            }
        } else {
            err = cg_process_window(ctx, &batch, &win, &res);
            if (err == CG_OK && res.n_events > res.events_cap) {                /* the chain's state has moved on: fetch again, do not redo */
                res.events_cap = res.n_events;
                cg_bed_event *ne = (cg_bed_event *)realloc(res.events, sizeof(cg_bed_event) * (size_t)res.events_cap);
                if (!ne) goto done;
                res.events = ne;
                err = cg_download(ctx, &res);
            }
        }
        if (err) { fprintf(stderr, "crumble: GPU path failed: %s (%s)\n", cg_strerror(err), cg_last_error(ctx)); goto done; }
        device_ms += cg_last_ms(ctx, CG_T_TOTAL);
        write_bed(o, header, &res);
        for (int i = 0; i < CG_N_COUNTERS; i++) o->counters[i] += res.counters[i];

        /* ---- records that are final now: everything except pileup records reaching the first incomplete column ---- */
        for (size_t i = lq.head; i < lq.n; i++) {
            live_rec *r = &lq.v[i];
            r->is_new = 0;
            if (r->final || r->slot < 0) continue;
            bam1_t *b = r->b;
            if (win.hi_tid >= 0 && r->in_pileup && b->core.tid == win.hi_tid && r->end > win.hi_pos) continue;
            if (b->core.l_qseq) {
                if (win.hi_tid >= 0 && r->in_pileup && b->core.tid == win.hi_tid && r->end > S) {   /* part of the next halo */
                    r->oq = (uint8_t *)malloc((size_t)b->core.l_qseq);
                    if (!r->oq) goto done;
                    memcpy(r->oq, bam_get_qual(b), (size_t)b->core.l_qseq);
                }
                memcpy(bam_get_qual(b), res.qual_out + batch.off[r->slot], (size_t)b->core.l_qseq);
            }
            r->final = 1;
        }
        /* ---- write the final prefix in input order; forget what no later call needs ---- */
        {
            size_t i = lq.head;
            int blocked = 0;
            for (; i < lq.n; i++) {
                live_rec *r = &lq.v[i];
                if (!r->final) blocked = 1;
                if (!blocked && r->final != 2) {
                    crumble_purge_tags(o, r->b);                                /* snp_score.c:1088 */
                    count_out++;
                    if (sam_write1(out, header, r->b) < 0) goto done;
                    r->final = 2;                                               /* written */
                }
            }
            /* drop from the head every written record the next call does not need as halo */
            while (lq.head < lq.n) {
                live_rec *r = &lq.v[lq.head];
                if (r->final != 2) break;
                int needed = win.hi_tid >= 0 && r->in_pileup && r->b->core.tid == win.hi_tid && r->end > S;
                if (needed) break;
                bam_destroy1(r->b); free(r->oq); r->b = NULL; r->oq = NULL;
                lq.head++;
            }
        }
        /* ---- next call ---- */
        if (win.hi_tid >= 0) { first = 0; lo_tid = win.hi_tid; lo_pos = S; cnt_pos = win.hi_pos; }
        else { first = 1; lo_tid = -1; lo_pos = cnt_pos = 0; }
    }
    if (count_in != count_out) {                                         /* snp_score.c:2021-2026 */
        fprintf(stderr, "ERROR: lost a read?\nRead  %" PRId64 " reads\nWrote %" PRId64 " reads\n\n", count_in, count_out);
        ret = 1;
    } else ret = 0;
    o->last_device_ms = device_ms;

done:
    if (nx) bam_destroy1(nx);
    lq_free(&lq);
    free(res.qual_out); free(res.events);
    if (ctx) cg_destroy(ctx);
    if (bb) cgb_destroy(bb);
    return ret;
}
