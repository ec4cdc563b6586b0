/*
 * TEST INFRASTRUCTURE ONLY — never linked into libcrumble_gpu.so.
 *
 * Host emulation of the device pipeline: runs the very same work-item bodies
 * (crumble_b200/csrc/cg_pipeline.h) in plain loops, with sequential stand-ins for the
 * device scans/compactions.  It exists so the CPU-only test-suite can check the
 * *decomposition* (per-column facts -> sparse chain -> per-read replay) against the
 * verbatim reference where no GPU is available.  It implements the C ABI of
 * include/crumble_gpu.h so the product's host driver (transcode_gpu.c) links unchanged.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <vector>
#include <limits.h>
#include "../../crumble_b200/csrc/cg_host.h"

struct EmuCarry { CgWin w; int chain_tid; int64_t td, tc; int depth_tid; };
struct cg_ctx { cg_params p; CgTables T; char err[256]; int64_t n_cols; std::vector<cg_bed_reg> bed; std::vector<int64_t> bed_pm; EmuCarry carry; std::vector<cg_bed_event> events; };

extern "C" int cg_device_count(void) { return 0; }
extern "C" int cg_enable_pinned(void) { return 0; }
extern "C" cg_ctx *cg_create(const cg_params *p, int device, int *err) {
    (void)device;
    const char *why = NULL;
    int e = cg_params_check(p, &why);
    if (e) { if (err) *err = e; fprintf(stderr, "unsupported: %s\n", why); return NULL; }
    cg_ctx *c = new cg_ctx();
    c->p = *p; cg_tables_init(&c->T, p); c->err[0] = 0; c->n_cols = 0;
    if (p->nbed) { c->bed.assign(p->bed, p->bed + p->nbed); c->bed_pm.resize(p->nbed); cg_bed_prefix_max(c->bed.data(), p->nbed, c->bed_pm.data()); c->p.bed = c->bed.data(); }
    if (err) *err = 0;
    return c;
}
extern "C" void cg_destroy(cg_ctx *c) { delete c; }
extern "C" const char *cg_last_error(const cg_ctx *c) { return c->err; }
extern "C" float cg_last_ms(const cg_ctx *, int) { return 0; }
extern "C" int64_t cg_last_launches(const cg_ctx *) { return 0; }
extern "C" int64_t cg_last_h2d_bytes(const cg_ctx *) { return 0; }
extern "C" int64_t cg_n_columns(const cg_ctx *c) { return c->n_cols; }

static int emu_process(cg_ctx *ctx, const cg_batch *in, const cg_window *win, cg_result *out);
/* events of the last call again (the caller's buffer was too small); qualities are already in the caller's buffer */
extern "C" int cg_download(cg_ctx *ctx, cg_result *out) {
    out->n_events = (int64_t)ctx->events.size();
    for (int64_t i = 0; i < out->n_events && i < out->events_cap; i++) out->events[i] = ctx->events[i];
    return 0;
}
extern "C" int cg_process(cg_ctx *ctx, const cg_batch *in, cg_result *out) { return emu_process(ctx, in, NULL, out); }
extern "C" int cg_process_window(cg_ctx *ctx, const cg_batch *in, const cg_window *win, cg_result *out) { return emu_process(ctx, in, win, out); }

static int emu_process(cg_ctx *ctx, const cg_batch *in, const cg_window *win, cg_result *out) {
    CgDev D; memset(&D, 0, sizeof(D));
    const int64_t n = in->n_reads;
    D.n_reads = n; D.n_cigar_total = in->n_cigar_total;
    D.tid = in->tid; D.pos = in->pos; D.flag = in->flag; D.mapq = in->mapq; D.l_qseq = in->l_qseq;
    D.n_cigar = in->n_cigar; D.off = in->off; D.cigar_off = in->cigar_off; D.cigar = in->cigar; D.seq = in->seq; D.qual = in->qual;
    D.qual_out = out->qual_out;
    D.T = &ctx->T; cg_devparams_from(&D.P, &ctx->p);
    EmuCarry cy; cg_win_reset(&cy.w); cy.chain_tid = -2; cy.td = cy.tc = 0; cy.depth_tid = -2;
    if (win) {
        D.P.win_on = 1;
        D.P.win_lo_tid = win->first == 1 ? -1 : win->lo_tid; D.P.win_lo_pos = win->lo_pos; D.P.win_cnt_pos = win->cnt_pos;
        D.P.win_hi_tid = win->hi_tid; D.P.win_hi_pos = win->hi_pos;
        if (!win->first) cy = ctx->carry;
    }
    EmuCarry snap = cy; bool snapped_depth = false, snapped_chain = false;
    D.bed = ctx->bed.data(); D.bed_pm = ctx->bed_pm.data();
    int32_t err = 0, maxdepth = 0, beyond = 0; D.err = &err; D.maxdepth = &maxdepth; D.beyond = &beyond;
    unsigned long long counters[CG_N_COUNTERS] = {0}; D.counters = counters;
    std::vector<int32_t> jmap(n + 1), rspan(n + 1);
    D.jmap = jmap.data(); D.rspan = rspan.data();
    for (int64_t r = 0; r < n; r++) cg_prep_read(&D, r);
    int np = 0;
    for (int64_t r = 0; r < n; r++) { jmap[r] = np; if (rspan[r]) np++; }
    D.n_pile = np;
    std::vector<CgRead> rd(np + 1); std::vector<int64_t> ks(np + 1), ke(np + 1), gap(np + 1);
    std::vector<int32_t> pmaxcol(np + 1), orig(np + 1); std::vector<uint8_t> r_bf(np + 1);
    D.rd = rd.data(); D.ks = ks.data(); D.ke = ke.data(); D.gap = gap.data(); D.pmaxcol = pmaxcol.data(); D.orig = orig.data(); D.r_bf = r_bf.data();
    for (int64_t r = 0; r < n; r++) cg_prep_keys(&D, r);
    for (int j = 1; j < np; j++) if (ke[j] < ke[j - 1]) ke[j] = ke[j - 1];       /* inclusive prefix max */
    for (int j = 0; j < np; j++) gap[j] = cg_gap_of(&D, j);
    std::vector<CgIsland> isl;
    const int64_t K0 = np ? ks[0] : 0;
    {
        int64_t acc = 0;
        for (int j = 0; j < np; j++) {
            int64_t g = gap[j]; acc += g; gap[j] = acc;
            if (j == 0 || g > 0) { CgIsland I; I.col_start = (int32_t)(ks[j] - K0 - acc); I.tid = (int32_t)(ks[j] >> 32); I.pos_start = (int32_t)(ks[j] & 0xffffffff); I.pad = 0; isl.push_back(I); }
        }
    }
    for (int j = 0; j < np; j++) cg_finish_read(&D, j, K0);
    if (err) { snprintf(ctx->err, sizeof ctx->err, "unsorted"); return err; }
    D.n_cols = np ? pmaxcol[np - 1] : 0; D.n_tiles = (D.n_cols + 31) / 32;
    D.n_islands = (int)isl.size(); D.isl = isl.data();
    ctx->n_cols = D.n_cols;
    std::vector<int32_t> tile_lo(D.n_tiles + 2, np), tile_start(D.n_tiles + 2, np);
    D.tile_lo = tile_lo.data(); D.tile_start = tile_start.data();
    for (int j = 0; j < np; j++) cg_tile_index(&D, j);
    std::vector<uint8_t> cb(D.n_cols + 1); std::vector<uint16_t> ev(D.n_cols + 1); std::vector<uint32_t> depth(D.n_cols + 1);
    D.cb = cb.data(); D.ev = ev.data(); D.depth = depth.data();
    const int cS = (win && win->hi_tid >= 0) ? cg_find_col(&D, win->hi_tid, win->next_lo_pos) : D.n_cols;
    std::vector<cg_column> dump;
    D.want_dump = out->columns != NULL;
    if (D.want_dump) { dump.resize(D.n_cols + 1); D.coldump = dump.data(); }
    for (int c = 0; c < D.n_cols; c++) {
        CgColOut o = cg_column_body(&D, c);
        for (int i = 0; i < CG_N_COUNTERS; i++) if (o.cnt >> i & 1) counters[i]++;
        if (o.n_plp > maxdepth) maxdepth = o.n_plp;
    }
    /* flagged columns */
    std::vector<int32_t> fcol;
    for (int c = 0; c < D.n_cols; c++) if (ev[c] & CG_EV_FLAGGED) fcol.push_back(c);
    const int nf = (int)fcol.size();
    std::vector<CgTrig> trig(nf + 1); std::vector<CgWin> twin(nf + 1);
    D.fcol = fcol.data(); D.trig = trig.data(); D.twin = twin.data(); D.n_flagged = nf;
    CgFlagScratch *S = new CgFlagScratch();
    for (int k = 0; k < nf; k++) {
        uint32_t cnt = cg_flagged(&D, k, S);
        for (int i = 0; i < CG_N_COUNTERS; i++) if (cnt >> i & 1) counters[i]++;
    }
    delete S;
    if (err) { snprintf(ctx->err, sizeof ctx->err, "overflow"); return err; }
    /* depth average (sequential restatement of snp_score.c:1478-1491,1673-1687) */
    {
        int64_t td = cy.td, tc = cy.tc; int is = 0, last_tid = cy.depth_tid;
        for (int c = 0; c < D.n_cols; c++) {
            if (c == cS && !snapped_depth) { snap.td = td; snap.tc = tc; snap.depth_tid = last_tid; snapped_depth = true; }
            while (is + 1 < D.n_islands && isl[is + 1].col_start <= c) is++;
            if (!(ev[c] & CG_EV_COUNTED)) continue;
            if (isl[is].tid != last_tid) { td = 0; tc = 0; last_tid = isl[is].tid; }
            td += depth[c]; tc++;
            if (ev[c] & CG_EV_PROCESSED) {
                uint32_t cnt = cg_deep_test(&D, c, td, tc);
                if (cnt) counters[CG_CNT_OVER_DEPTH]++;
                if (tc > 1024 * 1024) { tc >>= 1; td >>= 1; }
            }
        }
        if (!snapped_depth) { snap.td = td; snap.tc = tc; snap.depth_tid = last_tid; }
    }
    /* keep-window chain, sequential as cg_chain, continuing the state the previous call of a chain left */
    {
        CgWin w = cy.w; int last_tid = cy.chain_tid;
        for (int k = 0; k < nf; k++) {
            if (fcol[k] >= cS && !snapped_chain) { snap.w = w; snap.chain_tid = last_tid; snapped_chain = true; }
            const CgTrig t = trig[k];
            if (t.hasI || t.hasS) {
                if (t.tid != last_tid) { cg_win_reset(&w); last_tid = t.tid; }
                cg_win_step(&w, &t, &D.P);
            }
            twin[k] = w;
        }
        if (!snapped_chain) { snap.w = w; snap.chain_tid = last_tid; }
    }
    if (win && !win->first && cy.chain_tid == win->lo_tid && cy.w.min_pos != INT_MAX && cy.w.max_pos2 >= win->lo_pos)
        cg_paint_range(&D, win->lo_tid, win->lo_pos, cy.w.max_pos2);
    ctx->carry = snap;
    for (int k = 0; k < nf; k++) cg_paint(&D, k, nf);
    for (int64_t r = 0; r < n; r++) cg_rewrite(&D, r, nf);
    /* events */
    out->n_events = 0;
    ctx->events.clear();
    {
        int is = 0;
        for (int c = 0; c < D.n_cols; c++) {
            while (is + 1 < D.n_islands && isl[is + 1].col_start <= c) is++;
            int bits = cg_event_bits(ev[c]);
            for (int t = 0; t < 5; t++) if (bits >> t & 1) {
                cg_bed_event e; e.tid = isl[is].tid; e.pos = isl[is].pos_start + (c - isl[is].col_start); e.tag = t;
                ctx->events.push_back(e);
                if (out->events && out->n_events < out->events_cap) out->events[out->n_events] = e;
                out->n_events++;
            }
        }
    }
    if (beyond) counters[CG_CNT_COLUMNS]++;      /* snp_score.c:1476 runs before the region break at 1516-1517 */
    for (int i = 0; i < CG_N_COUNTERS; i++) out->counters[i] = (int64_t)counters[i];
    out->n_columns = 0;
    if (out->columns) {
        for (int c = 0; c < D.n_cols; c++) {
            if (dump[c].tid < 0) continue;
            cg_column z = dump[c];
            if (cb[c] & CG_CB_ACTIVE) z.flags |= 4;
            if (cb[c] & CG_CB_KEEP) z.flags |= 2;
            if (ev[c] & CG_EV_TRIGGER) z.flags |= 16;
            if (ev[c] & CG_EV_HADINDEL) z.flags |= 32;
            z.flags |= (uint32_t)(ev[c] & CG_EV_BEDMASK) << 8;
            if (out->n_columns < out->columns_cap) out->columns[out->n_columns] = z;
            out->n_columns++;
        }
    }
    return 0;
}
