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
#include "../../crumble_b200/csrc/cg_column_lean.h"

struct EmuCarry { CgWin w; int chain_tid; int64_t td, tc; int depth_tid; };
struct cg_ctx { cg_params p; CgTables T; char err[256]; int64_t n_cols; std::vector<cg_bed_reg> bed; std::vector<int64_t> bed_pm; EmuCarry carry; std::vector<cg_bed_event> events;
                const cg_batch *sh_in; cg_window sh_win; cg_result *sh_out; int sh_state; };

/* EMU_DEVICES=n makes the host driver (transcode_gpu.c with CRUMBLE_GPUS) and the scheduler (cg_multi.c) believe in n devices: every "device" is
 * one emulated context, the scheduler's threads, cuts, halos, carries and gather run as they do on a box of GPUs */
extern "C" int cg_device_count(void) { const char *e = getenv("EMU_DEVICES"); return e ? atoi(e) : 0; }
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

/* the three phases of a region shard (include/crumble_gpu.h).  The emulation is sequential, so everything runs in the middle phase, where the
 * true state of the left neighbour is known; the protocol (call order, 128-byte state, halo bytes into the side buffer) is the device's */
extern "C" int cg_shard_begin(cg_ctx *ctx, const cg_batch *in, const cg_window *win, cg_result *out) {
    if (!in || !win) return CG_ERR_BAD_ARG;
    ctx->sh_in = in; ctx->sh_win = *win; ctx->sh_out = out; ctx->sh_state = 1;
    if (ctx->sh_win.first == 2) ctx->sh_win.first = 0;
    return 0;
}
extern "C" int cg_shard_carry(cg_ctx *ctx, const void *carry_in, void *carry_out) {
    static_assert(sizeof(EmuCarry) <= CG_CARRY_BYTES, "the emulated state must fit the carry blob");
    if (ctx->sh_state != 1) return CG_ERR_STATE;
    if (!ctx->sh_win.first && !carry_in) return CG_ERR_BAD_ARG;
    if (carry_in && !ctx->sh_win.first) memcpy(&ctx->carry, carry_in, sizeof(EmuCarry));
    cg_result *out = ctx->sh_out;
    const int64_t qb = ctx->sh_in->qual_bytes;
    std::vector<uint8_t> tmp((size_t)qb + 8);
    cg_result r = *out; r.qual_out = tmp.data(); r.qual_head = NULL; r.head_bytes = 0;
    int e = emu_process(ctx, ctx->sh_in, &ctx->sh_win, &r);
    if (e) return e;
    const int64_t hb = out->qual_head ? (out->head_bytes < qb ? out->head_bytes : qb) : 0;
    if (hb > 0) memcpy(out->qual_head, tmp.data(), (size_t)hb);
    if (out->qual_out && qb > hb) memcpy(out->qual_out + hb, tmp.data() + hb, (size_t)(qb - hb));
    out->n_events = r.n_events; out->n_columns = r.n_columns;
    memcpy(out->counters, r.counters, sizeof out->counters);
    if (carry_out) { memset(carry_out, 0, CG_CARRY_BYTES); memcpy(carry_out, &ctx->carry, sizeof(EmuCarry)); }
    ctx->sh_state = 2;
    return 0;
}
extern "C" int cg_shard_end(cg_ctx *ctx, cg_result *out) {
    (void)out;
    if (ctx->sh_state != 2) return CG_ERR_STATE;
    ctx->sh_state = 0;
    return 0;
}
extern "C" int cg_process_window(cg_ctx *ctx, const cg_batch *in, const cg_window *win, cg_result *out) { return emu_process(ctx, in, win, out); }

static int emu_process(cg_ctx *ctx, const cg_batch *in, const cg_window *win, cg_result *out) {
    if (getenv("EMU_NULL")) {                            /* host-pipeline profiling (tools/file_bench_host.sh): the device does nothing but hand the qualities back */
        if (out->qual_out && in->qual_bytes) memcpy(out->qual_out, in->qual, (size_t)in->qual_bytes);
        out->n_events = 0; out->n_columns = 0; memset(out->counters, 0, sizeof out->counters);
        ctx->events.clear();
        return 0;
    }
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
    D.want_dump = 1;                                     /* always: the tuned-path check below compares the consensus values themselves, not only the decisions */
    dump.resize(D.n_cols + 1); D.coldump = dump.data();
    for (int c = 0; c < D.n_cols; c++) {
        CgColOut o = cg_column_body(&D, c);
        for (int i = 0; i < CG_N_COUNTERS; i++) if (o.cnt >> i & 1) counters[i]++;
        if (o.n_plp > maxdepth) maxdepth = o.n_plp;
    }
    /* The tuned column path (k_cells + k_column of cg_device.cu), lane by lane, with the same bodies the kernels run (cg_cells.h,
     * cg_column_lean.h): cell matrix in groups of 8, staged rows of a tile in chunks, rank-space accumulation, lean finalisation.
     * Every per-column output must equal what the plain body just produced. */
    if (!cg_params_generic(&ctx->p) && np > 0) {
        const int doB = D.P.min_qual_B != 0;
        std::vector<CgCellRec> crec(np + 1);
        uint64_t ng = 0;
        for (int j = 0; j < np; j++) { CgCellRec r; r.cpos8 = (uint32_t)ng; r.col0 = rd[j].col0; r.span = rd[j].span; r.ngrp = cg_cell_ngroups(r.col0, r.span); crec[j] = r; ng += r.ngrp; }
        std::vector<uint16_t> cells(ng * 8 + 8, 0xdead);
        /* the kernels read whole aligned words around a read (CG_FRONT_PAD / tail padding on the device): give the host copies the same slack */
        std::vector<uint8_t> qpad(in->qual_bytes + 512, 0), spad(in->seq_bytes + 512, 0);
        memcpy(qpad.data() + 256, in->qual, in->qual_bytes); memcpy(spad.data() + 256, in->seq, in->seq_bytes);
        CgDev D2 = D; D2.qual = qpad.data() + 256; D2.seq = spad.data() + 256;
        for (int j = 0; j < np; j++)
            for (uint32_t k = 0; k < crec[j].ngrp; k++) { uint32_t o4[4]; cg_cells8(&D2, &rd[j], (int)k, doB, o4); memcpy(&cells[((size_t)crec[j].cpos8 + k) * 8], o4, 16); }
        ColTabRow tab[104];
        for (int i = 0; i < 104; i++) { int q = i > 100 ? 100 : i; tab[i].MM = ctx->T.MM[q]; tab[i].hM = ctx->T._M[q]; tab[i].om = ctx->T.omq2p[q]; tab[i].pad = 0; }
        std::vector<uint8_t> cb2(D.n_cols + 1, 0); std::vector<uint16_t> ev2(D.n_cols + 1, 0); std::vector<uint32_t> depth2(D.n_cols + 1, 0);
        std::vector<cg_column> dump2(D.n_cols + 1);
        unsigned long long counters2[CG_N_COUNTERS] = {0}; int32_t err2 = 0, maxdepth2 = 0, beyond2 = 0;
        D2 = D; D2.cb = cb2.data(); D2.ev = ev2.data(); D2.depth = depth2.data(); D2.coldump = dump2.data(); D2.counters = counters2; D2.err = &err2; D2.maxdepth = &maxdepth2; D2.beyond = &beyond2;
        const int R = getenv("CG_EMU_ROWS") ? atoi(getenv("CG_EMU_ROWS")) : 60;
        std::vector<uint16_t> buf((size_t)(R + 1) * 32);
        const int lean = doB && !D.P.min_qual_A;
        for (int t = 0; t < D.n_tiles; t++) {
            const int lo = tile_lo[t], hi = tile_start[t + 1], tile_c0 = t * 32;
            CgRankAcc acc[32]; double rare[32][9];
            for (int l = 0; l < 32; l++) cg_rank_init<1>(&acc[l], rare[l]);
            for (int j0 = lo; j0 < hi; j0 += R) {
                const int n = hi - j0 < R ? hi - j0 : R;
                for (int r = 0; r < n; r++) {
                    const CgCellRec cr = crec[j0 + r];
                    for (int p4 = 0; p4 < 4; p4++) {
                        const int gi = (tile_c0 + 8 * p4 - (cr.col0 & ~7)) >> 3;
                        if ((unsigned)gi < cr.ngrp) memcpy(&buf[(size_t)r * 32 + 8 * p4], &cells[((size_t)cr.cpos8 + gi) * 8], 16);
                        else memset(&buf[(size_t)r * 32 + 8 * p4], 0, 16);
                    }
                }
                for (int l = 0; l < 32; l++) { cg_rank_peek<32>(&acc[l], &buf[l], n); cg_rank_rows<32, 1>(&acc[l], rare[l], &buf[l], n, (const ColTabRow *)tab); }
            }
            for (int l = 0; l < 32; l++) {
                const int c = tile_c0 + l;
                if (c >= D.n_cols) continue;
                CgRankAcc &a = acc[l];
                double dmp[15]; int rk[5]; CgConsAcc A;
                cg_rank_dump_S<1, 1>(&a, rare[l], dmp); cg_rank_of_bases(a.pi, a.nseen, rk); cg_rank_unpermute_S<1>(dmp, rk, A.S);
                cg_rank_dump_C<1, 1>(&a, rare[l], dmp); cg_rank_unpermute_C<1>(dmp, rk, A.sumsC);
                CgColStats st; st.cp = 0; st.n_plp = a.n_plp; st.n_skip = a.n_skip; st.low_mq = a.low_mq; st.had_indel = a.indel_cnt > 0; st.indel_cnt = a.indel_cnt;
                st.clipped = a.clipped; st.n_overlap = a.n_overlap; st.ins_seen = a.ins_seen != 0;
                A.depth = a.n_plp - a.n_skip - a.n_none; A.nN = a.nN; A.sumsE = 0;
                CgColOut o;
                if (lean) { CgCons cB; cg_cons_finalize_lean(&ctx->T, &A, A.depth, A.nN, &cB); o = cg_column_finish(&D2, c, lo, hi, &st, NULL, &cB); }
                else o = cg_column_body(&D2, c);
                for (int i = 0; i < CG_N_COUNTERS; i++) if (o.cnt >> i & 1) counters2[i]++;
                if (o.n_plp > maxdepth2) maxdepth2 = o.n_plp;
            }
        }
        int bad = 0;
        for (int c = 0; c < D.n_cols && bad < 10; c++) {
            if (cb2[c] != cb[c] || ev2[c] != ev[c] || depth2[c] != depth[c]) { fprintf(stderr, "emu: tuned column path differs at column %d: cb %02x/%02x ev %04x/%04x depth %u/%u\n", c, cb2[c], cb[c], ev2[c], ev[c], depth2[c], depth[c]); bad++; }
            if (D.want_dump && memcmp(&dump2[c], &dump[c], sizeof(cg_column))) { fprintf(stderr, "emu: tuned column path: column dump differs at column %d (phred %d/%d het %d/%d)\n", c, dump2[c].phred, dump[c].phred, dump2[c].het_phred, dump[c].het_phred); bad++; }
        }
        for (int i = 0; i < CG_N_COUNTERS; i++) if (counters2[i] != counters[i]) { fprintf(stderr, "emu: tuned column path: counter %d differs %llu/%llu\n", i, counters2[i], counters[i]); bad++; }
        if (maxdepth2 != maxdepth || beyond2 != beyond) { fprintf(stderr, "emu: tuned column path: maxdepth/beyond differ\n"); bad++; }
        if (bad) { snprintf(ctx->err, sizeof ctx->err, "tuned column path differs from the plain body"); return CG_ERR_CUDA; }
        if (getenv("CG_EMU_VERBOSE")) fprintf(stderr, "emu: tuned column path verified on %d columns, %d tiles, %llu cell groups (rows per chunk %d, lean finalise %d)\n", D.n_cols, D.n_tiles, (unsigned long long)ng, R, lean);
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
