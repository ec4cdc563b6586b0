/*
 * transcode_gpu.c — host driver with the reference's transcode() contract
 * (snp_score.c:1336-2029): read records, hand them to the device path through the C ABI
 * of include/crumble_gpu.h, write them back in input order with rewritten qualities.
 *
 * What stays on the host, as in the reference: record I/O, the BED text lines
 * (snp_score.c:1496-1498 etc.), the -v counters, the read-count check (2021-2026).
 * Everything between "record decoded" and "record ready to write" runs on the GPU.
 *
 * v1 limitation: the whole input (or -r region) is processed as one batch.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <inttypes.h>
#include "htslib/sam.h"
#include "crumble_host.h"

typedef struct { bam1_t **v; size_t n, cap; } recvec;

static int rv_push(recvec *rv, bam1_t *b) {
    if (rv->n == rv->cap) {
        size_t nc = rv->cap ? rv->cap * 2 : 4096;
        bam1_t **nv = (bam1_t **)realloc(rv->v, nc * sizeof(*nv));
        if (!nv) return -1;
        rv->v = nv; rv->cap = nc;
    }
    rv->v[rv->n++] = b;
    return 0;
}

int transcode_gpu(crumble_opts *o, samFile *in, samFile *out, bam_hdr_t *header, hts_itr_t *h_iter) {
    int ret = -1, err = 0;
    recvec rv = {0};
    cg_batch_builder *bb = NULL;
    cg_ctx *ctx = NULL;
    cg_result res; memset(&res, 0, sizeof(res));
    int64_t count_in = 0, count_out = 0;

    cg_params p = o->p;
    if (h_iter) { p.region_tid = h_iter->tid; p.region_beg = h_iter->beg; p.region_end = h_iter->end; }

    if (!(bb = cgb_create(1))) goto done;
    bam1_t *b = bam_init1();
    for (;;) {
        int r = h_iter ? sam_itr_next(in, h_iter, b) : sam_read1(in, header, b);
        if (r < -1) { fprintf(stderr, "Error reading input\n"); bam_destroy1(b); goto done; }
        if (r < 0) break;
        count_in++;
        int e = cgb_add(bb, b->core.tid, b->core.pos, b->core.flag, b->core.qual, b->core.l_qseq,
                        b->core.n_cigar, bam_get_cigar(b), bam_get_seq(b), bam_get_qual(b));
        if (e) { fprintf(stderr, "crumble: %s\n", cg_strerror(e)); bam_destroy1(b); goto done; }
        bam1_t *d = bam_dup1(b);
        if (!d || rv_push(&rv, d) < 0) { bam_destroy1(b); goto done; }
    }
    bam_destroy1(b);

    cg_batch batch;
    if ((err = cgb_finish(bb, &batch)) != 0) { fprintf(stderr, "crumble: %s\n", cg_strerror(err)); goto done; }

    ctx = cg_create(&p, o->device, &err);
    if (!ctx) { fprintf(stderr, "crumble: cannot create GPU context: %s\n", cg_strerror(err)); goto done; }

    res.qual_out = (uint8_t *)malloc((size_t)batch.qual_bytes + 16);
    res.events_cap = 1 << 16;
    res.events = (cg_bed_event *)malloc(sizeof(cg_bed_event) * (size_t)res.events_cap);
    if (!res.qual_out || !res.events) goto done;
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
    if (err) { fprintf(stderr, "crumble: GPU path failed: %s (%s)\n", cg_strerror(err), cg_last_error(ctx)); goto done; }

    /* BED lines, in column order (snp_score.c:1496-1498,1676-1678,1768-1770,1802-1804,1810-1812) */
    if (o->bed_fp) {
        static const char *tag[5] = { "VDEEP", "DEEP", "CLIP", "INDEL_LEN", "INDEL_COVERAGE" };
        for (int64_t i = 0; i < res.n_events; i++) {
            const cg_bed_event *e = &res.events[i];
            int s = e->pos - 50; if (s < 0) s = 0;
            fprintf(o->bed_fp, "%s\t%d\t%d\t%s\n", header->target_name[e->tid], s, e->pos + 50, tag[e->tag]);
        }
    }
    for (int i = 0; i < CG_N_COUNTERS; i++) o->counters[i] += res.counters[i];

    for (size_t i = 0; i < rv.n; i++) {
        bam1_t *r = rv.v[i];
        if (r->core.l_qseq) memcpy(bam_get_qual(r), res.qual_out + batch.off[i], (size_t)r->core.l_qseq);
        crumble_purge_tags(o, r);                                        /* snp_score.c:1088 */
        count_out++;
        if (sam_write1(out, header, r) < 0) goto done;
    }
    if (count_in != count_out) {                                         /* snp_score.c:2021-2026 */
        fprintf(stderr, "ERROR: lost a read?\nRead  %" PRId64 " reads\nWrote %" PRId64 " reads\n\n", count_in, count_out);
        ret = 1;
    } else ret = 0;
    o->last_device_ms = cg_last_ms(ctx, CG_T_TOTAL);

done:
    for (size_t i = 0; i < rv.n; i++) bam_destroy1(rv.v[i]);
    free(rv.v);
    free(res.qual_out); free(res.events);
    if (ctx) cg_destroy(ctx);
    if (bb) cgb_destroy(bb);
    return ret;
}
