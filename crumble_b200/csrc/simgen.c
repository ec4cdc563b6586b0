/*
 * Synthetic aligned-read generator (see simgen.h).  Workload shapes follow
 * diploid sample with SNPs/indels/STR loci, Illumina-like
 * qualities, mapq mix, soft clips, and *planted* features so that every keep
 * heuristic of the reference fires (snp_score.c:1668,1673,1764,1798,1809,1627):
 * clip pile-ups, depth spikes, multi-length insertion sites, low-span insertion
 * sites, three-allele paralog patches and low-mapq majorities.
 *
 * Reads are sampled independently (mate fields are synthesised, no real mate is
 * emitted with matching name) so the stream can be produced in coordinate order by
 * independent chunks on many threads; crumble never looks at mate information.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include "simgen.h"

/* ---- rng: xoshiro256** seeded by splitmix64 --------------------------------- */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t a, uint64_t b) {
    uint64_t x = a * 0xD1342543DE82EF95ULL + b * 0x2545F4914F6CDD1DULL + 0x1234567;
    for (int i = 0; i < 4; i++) r->s[i] = splitmix(&x);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_u64(rng_t *r) {
    uint64_t *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static inline uint32_t rng_u32(rng_t *r) { return (uint32_t)(rng_u64(r) >> 32); }
static inline double rng_unit(rng_t *r) { return (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline int rng_int(rng_t *r, int lo, int hi) { return lo + (int)(rng_u32(r) % (uint32_t)(hi - lo + 1)); }
static inline double rng_norm(rng_t *r) {   /* Irwin-Hall(4), variance 1 */
    uint64_t a = rng_u64(r), b = rng_u64(r);
    double s = (double)(a & 0xffffffffu) + (double)(a >> 32) + (double)(b & 0xffffffffu) + (double)(b >> 32);
    return (s / 4294967296.0 - 2.0) * 1.7320508075688772;
}
static inline uint64_t hash64(uint64_t x) { return splitmix(&x); }

/* ---- contig model -------------------------------------------------------------- */
enum { V_SNP = 0, V_INS = 1, V_DEL = 2 };
typedef struct { int pos; uint8_t type, gt, alt; uint8_t len; uint8_t ins[30]; } variant;
enum { F_CLIP, F_SPIKE, F_MULTIINS, F_LOWSPAN, F_PARALOG, F_LOWMQ, F_REPEAT };
typedef struct { int type, start, end; uint64_t salt; } feature;

typedef struct {
    int tid; int64_t len;
    uint8_t *ref;                     /* 0..3 */
    variant *var; int nvar;
    feature *feat; int nfeat;
} contig;

static int cmp_var(const void *a, const void *b) { return ((const variant *)a)->pos - ((const variant *)b)->pos; }
static int cmp_feat(const void *a, const void *b) { return ((const feature *)a)->start - ((const feature *)b)->start; }

static void contig_build(contig *c, const simgen_cfg *cfg, int tid, int64_t len) {
    rng_t r; rng_seed(&r, cfg->seed, 1000 + (uint64_t)tid);
    c->tid = tid; c->len = len;
    c->ref = (uint8_t *)malloc((size_t)len + 64);
    /* i.i.d. bases, 41 % GC */
    for (int64_t i = 0; i < len; i++) {
        uint32_t u = rng_u32(&r);
        int gc = (u & 0xffff) < (uint32_t)(0.41 * 65536);
        c->ref[i] = (uint8_t)(gc ? ((u >> 16 & 1) ? 1 : 2) : ((u >> 16 & 1) ? 0 : 3));
    }
    /* STR loci every ~2 kb; remember them for STR-indels */
    int nstr_cap = (int)(len / 1500) + 8, nstr = 0;
    int *str_pos = (int *)malloc(sizeof(int) * (size_t)nstr_cap), *str_unit = (int *)malloc(sizeof(int) * (size_t)nstr_cap),
        *str_copies = (int *)malloc(sizeof(int) * (size_t)nstr_cap);
    for (int64_t p = 500 + rng_int(&r, 0, 1000); p + 200 < len && nstr < nstr_cap; p += 1000 + rng_int(&r, 0, 2000)) {
        int unit, copies;
        if (rng_u32(&r) & 1) { unit = 1; copies = rng_int(&r, 6, 20); }
        else { unit = rng_int(&r, 2, 6); copies = rng_int(&r, 4, 15); }
        uint8_t u[6];
        for (int k = 0; k < unit; k++) u[k] = (uint8_t)(rng_u32(&r) & 3);
        if (unit > 1 && u[0] == u[1]) u[1] = (uint8_t)((u[1] + 1) & 3);
        for (int k = 0; k < unit * copies; k++) c->ref[p + k] = u[k % unit];
        str_pos[nstr] = (int)p; str_unit[nstr] = unit; str_copies[nstr] = copies; nstr++;
    }
    /* variants: SNPs 1/1000 (2/3 het), indels 1/8000 (half inside STR loci) */
    int nsnp = (int)(len / 1000), nindel = (int)(len / 8000) + 1;
    c->var = (variant *)calloc((size_t)(nsnp + nindel + 8), sizeof(variant));
    int nv = 0;
    for (int i = 0; i < nsnp; i++) {
        variant *v = &c->var[nv++];
        v->pos = (int)(rng_u64(&r) % (uint64_t)len);
        v->type = V_SNP; v->len = 1;
        v->alt = (uint8_t)((c->ref[v->pos] + 1 + rng_u32(&r) % 3) & 3);
        uint32_t g = rng_u32(&r) % 3; v->gt = (uint8_t)(g == 0 ? 3 : g);   /* 1/3 hom, 1/3 hap0, 1/3 hap1 */
    }
    for (int i = 0; i < nindel; i++) {
        variant *v = &c->var[nv];
        int l = 1; while (l < 30 && rng_unit(&r) < 0.667) l++;                /* geometric, mean 3 */
        int is_ins = rng_u32(&r) & 1;
        if ((rng_u32(&r) & 1) && nstr) {
            int s = (int)(rng_u32(&r) % (uint32_t)nstr), k = rng_int(&r, 1, 3);
            if (k >= str_copies[s]) k = 1;
            l = k * str_unit[s]; if (l > 30) l = str_unit[s];
            v->pos = str_pos[s] + str_unit[s];          /* one unit into the locus */
            if (is_ins) for (int j = 0; j < l; j++) v->ins[j] = c->ref[str_pos[s] + j % str_unit[s]];
        } else {
            v->pos = 100 + (int)(rng_u64(&r) % (uint64_t)(len - 200));
            if (is_ins) for (int j = 0; j < l; j++) v->ins[j] = (uint8_t)(rng_u32(&r) & 3);
        }
        v->type = (uint8_t)(is_ins ? V_INS : V_DEL); v->len = (uint8_t)l;
        uint32_t g = rng_u32(&r) % 3; v->gt = (uint8_t)(g == 0 ? 3 : g);
        nv++;
    }
    qsort(c->var, (size_t)nv, sizeof(variant), cmp_var);
    /* one event per position, and nothing inside a preceding deletion */
    int w = 0, block_end = -1;
    for (int i = 0; i < nv; i++) {
        if (c->var[i].pos <= block_end) continue;
        c->var[w] = c->var[i];
        block_end = c->var[w].pos + (c->var[w].type == V_DEL ? c->var[w].len : 0);
        w++;
    }
    c->nvar = w;
    free(str_pos); free(str_unit); free(str_copies);

    /* planted features */
    int per_kind = (int)ceil(cfg->features_per_mb * (double)len / 1e6);
    if (cfg->features_per_mb > 0 && per_kind < 1) per_kind = 1;
    int nrep = (int)(0.02 * (double)len / 2000.0);
    c->feat = (feature *)calloc((size_t)(per_kind * 6 + nrep + 8), sizeof(feature));
    int nf = 0;
    static const int span[6] = { 1, 300, 1, 1, 500, 600 };
    for (int t = 0; t < 6; t++)
        for (int i = 0; i < per_kind; i++) {
            feature *f = &c->feat[nf++];
            f->type = t;
            f->start = 300 + (int)(rng_u64(&r) % (uint64_t)(len > 1200 ? len - 1200 : 1));
            f->end = f->start + span[t];
            f->salt = rng_u64(&r);
        }
    for (int i = 0; i < nrep; i++) {
        feature *f = &c->feat[nf++];
        f->type = F_REPEAT;
        f->start = (int)(rng_u64(&r) % (uint64_t)(len > 4000 ? len - 4000 : 1));
        f->end = f->start + 2000; f->salt = rng_u64(&r);
    }
    qsort(c->feat, (size_t)nf, sizeof(feature), cmp_feat);
    c->nfeat = nf;
}

static void contig_free(contig *c) { free(c->ref); free(c->var); free(c->feat); }

/* ---- output buffers -------------------------------------------------------------- */
typedef struct { uint8_t *p; size_t n, cap; int64_t reads, bases; } obuf;
static void ob_reserve(obuf *o, size_t extra) {
    if (o->n + extra <= o->cap) return;
    size_t nc = o->cap ? o->cap : (1u << 20);
    while (nc < o->n + extra) nc += nc >> 1;
    o->p = (uint8_t *)realloc(o->p, nc); o->cap = nc;
}
static inline void put32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }
static inline void put16(uint8_t *p, uint16_t v) { memcpy(p, &v, 2); }

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

/* read under construction */
typedef struct {
    uint8_t  base[512];      /* 0..3, 4 = N */
    uint8_t  qual[512];
    uint32_t cig[128]; int ncig;
    int l, reflen;
} rd_t;

static inline void cig_push(rd_t *r, int op, int len) {
    if (len <= 0) return;
    if (r->ncig && (int)(r->cig[r->ncig - 1] & 0xf) == op) { r->cig[r->ncig - 1] += (uint32_t)len << 4; return; }
    if (r->ncig < 128) r->cig[r->ncig++] = (uint32_t)len << 4 | (uint32_t)op;
}

static void emit_record(obuf *o, int tid, int pos, int mapq, int flag, const rd_t *r, int mtid, int mpos, int isize,
                        const char *name, int aligned) {
    static const uint8_t nt16[5] = { 1, 2, 4, 8, 15 };
    int ln = (int)strlen(name) + 1, l = r->l, nc = aligned ? r->ncig : 0;
    uint32_t bs = 32u + (uint32_t)ln + 4u * (uint32_t)nc + (uint32_t)((l + 1) >> 1) + (uint32_t)l;
    ob_reserve(o, bs + 4);
    uint8_t *p = o->p + o->n;
    put32(p, bs); put32(p + 4, (uint32_t)tid); put32(p + 8, (uint32_t)pos);
    p[12] = (uint8_t)ln; p[13] = (uint8_t)mapq;
    put16(p + 14, (uint16_t)reg2bin(pos, pos + (aligned && r->reflen ? r->reflen : 1)));
    put16(p + 16, (uint16_t)nc); put16(p + 18, (uint16_t)flag);
    put32(p + 20, (uint32_t)l); put32(p + 24, (uint32_t)mtid); put32(p + 28, (uint32_t)mpos); put32(p + 32, (uint32_t)isize);
    memcpy(p + 36, name, (size_t)ln);
    uint8_t *q = p + 36 + ln;
    memcpy(q, r->cig, 4u * (size_t)nc); q += 4 * nc;
    memset(q, 0, (size_t)((l + 1) >> 1));
    for (int i = 0; i < l; i++) q[i >> 1] |= (uint8_t)(nt16[r->base[i]] << ((~i & 1) << 2));
    q += (l + 1) >> 1;
    memcpy(q, r->qual, (size_t)l);
    o->n += bs + 4;
    o->reads++;
    if (aligned) o->bases += l;
}

/* quality model */
static uint32_t g_perr[64];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void init_tables(void) {
    for (int q = 0; q < 64; q++) g_perr[q] = (uint32_t)(pow(10.0, -q / 10.0) * 4294967295.0);
}
static inline int qual_draw(rng_t *r, int cycle, int L, int binned) {
    double x = (double)cycle / (double)L;
    double q = 39.0 - 14.0 * x * x + 2.5 * rng_norm(r);
    if ((rng_u32(r) & 63) == 0) q = 2 + (rng_u32(r) % 14);
    int qi = (int)(q + 0.5);
    if (qi < 2) qi = 2;
    if (qi > 41) qi = 41;
    if (binned) qi = qi < 7 ? 2 : qi < 18 ? 11 : qi < 31 ? 25 : 37;
    return qi;
}

typedef struct { int clip_at; int ins_at, ins_len; int paralog; int lowmq; int repeat; } rd_mods;

/* Sample one read starting at reference position p from haplotype h. Returns 0 if unusable. */
static int sample_read(const contig *c, const simgen_cfg *cfg, rng_t *r, int p, int h, int reverse,
                       const rd_mods *m, rd_t *rd) {
    const int L = cfg->read_len;
    rd->ncig = 0; rd->l = 0; rd->reflen = 0;
    int q = 0, rp = p;
    /* optional leading soft clip (2 % of reads get a clip at one end) */
    int sc_lead = 0, sc_tail = 0;
    uint32_t u = rng_u32(r) % 100;
    if (u == 0) sc_lead = rng_int(r, 5, 50); else if (u == 1) sc_tail = rng_int(r, 5, 50);
    for (int i = 0; i < sc_lead; i++) rd->base[q++] = (uint8_t)(rng_u32(r) & 3);
    cig_push(rd, 4, sc_lead);
    /* first variant at or after p */
    int lo = 0, hi = c->nvar;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (c->var[mid].pos < p) lo = mid + 1; else hi = mid; }
    int vi = lo, run = 0, last = -1;
    const int Lalign = L - sc_tail;
    while (q < Lalign) {
        if (rp >= c->len) break;
        if (m->clip_at >= 0 && rp >= m->clip_at && q > sc_lead) break;     /* breakpoint: rest is soft clipped */
        if (m->ins_at >= 0 && rp == m->ins_at && q > sc_lead && q + m->ins_len < Lalign) {
            for (int j = 0; j < m->ins_len; j++) rd->base[q++] = (uint8_t)(rng_u32(r) & 3);
            cig_push(rd, 1, m->ins_len);
        }
        while (vi < c->nvar && c->var[vi].pos < rp) vi++;
        if (vi < c->nvar && c->var[vi].pos == rp && (c->var[vi].gt >> h & 1)) {
            const variant *v = &c->var[vi++];
            if (v->type == V_SNP) { rd->base[q++] = v->alt; cig_push(rd, 0, 1); rp++; continue; }
            if (v->type == V_INS) {
                if (q > sc_lead && q + v->len < Lalign) {
                    for (int j = 0; j < v->len; j++) rd->base[q++] = v->ins[j];
                    cig_push(rd, 1, v->len);
                }
                /* fall through to emit the reference base at rp */
            } else {   /* deletion */
                if (q > sc_lead && rp + v->len < c->len) { cig_push(rd, 2, v->len); rp += v->len; continue; }
            }
        }
        int b = c->ref[rp];
        run = (b == last) ? run + 1 : 1; last = b;
        /* indel sequencing errors: 1e-5 per base, x20 inside homopolymers >= 8 */
        uint32_t e = rng_u32(r);
        uint32_t thr = run >= 8 ? 858993u : 42950u;
        if (e < thr && q > sc_lead && q + 2 < Lalign) {
            if (e & 1) { rd->base[q++] = (uint8_t)b; cig_push(rd, 1, 1); }
            else { cig_push(rd, 2, 1); rp++; continue; }
        }
        if (m->paralog) {
            uint64_t hsh = hash64((uint64_t)rp * 0x9E3779B1u ^ (uint64_t)c->tid << 40);
            if (hsh % 150 == 0) b = (b + m->paralog) & 3;                         /* three-allele site */
            else if (hash64(hsh + (uint64_t)m->paralog) % 50 == 0) b = (b + 2) & 3;   /* copy-specific divergence */
        }
        rd->base[q++] = (uint8_t)b; cig_push(rd, 0, 1); rp++;
    }
    /* a trailing insertion is reported as soft clip by real aligners */
    if (rd->ncig && (rd->cig[rd->ncig - 1] & 0xf) == 1) rd->cig[rd->ncig - 1] = (rd->cig[rd->ncig - 1] & ~0xfu) | 4u;
    int aligned_q = q;
    while (q < L) rd->base[q++] = (uint8_t)(rng_u32(r) & 3);
    cig_push(rd, 4, L - aligned_q);
    rd->l = L; rd->reflen = rp - p;
    if (rd->reflen <= 0) return 0;
    /* qualities + substitution errors (cycle order follows the strand) */
    for (int i = 0; i < L; i++) {
        int cyc = reverse ? L - 1 - i : i;
        int qv = qual_draw(r, cyc, L, cfg->qual_binned);
        rd->qual[i] = (uint8_t)qv;
        if (rng_u32(r) < g_perr[qv]) rd->base[i] = (uint8_t)((rd->base[i] + 1 + rng_u32(r) % 3) & 3);
    }
    if ((rng_u32(r) & 1023) == 0) rd->base[rng_u32(r) % (uint32_t)L] = 4;   /* occasional N */
    return 1;
}

static inline int poisson(rng_t *r, double lambda) {
    double l = exp(-lambda), p = 1.0; int k = 0;
    do { k++; p *= rng_unit(r); } while (p > l);
    return k - 1;
}

static int pick_mapq(rng_t *r, const rd_mods *m) {
    if (m->lowmq && rng_u32(r) % 10 < 7) return rng_int(r, 0, 5);
    if (m->paralog) return rng_int(r, 0, 40);
    if (m->repeat) return (rng_u32(r) & 1) ? rng_int(r, 0, 20) : 60;
    uint32_t u = rng_u32(r) % 100;
    if (u < 90) return 60;
    if (u < 93) return 0;
    return rng_int(r, 1, 59);
}

/* WGS-like chunk: all reads starting in [beg, end) of contig c */
static void gen_chunk(const contig *c, const simgen_cfg *cfg, int chunk_id, int beg, int end, obuf *o) {
    rng_t r; rng_seed(&r, cfg->seed ^ 0xC0FFEE, ((uint64_t)c->tid << 32) | (uint32_t)chunk_id);
    const int L = cfg->read_len;
    const double lam = cfg->depth / (double)L;
    int fi = 0;
    rd_t rd; char name[64]; uint32_t serial = 0;
    for (int p = beg; p < end; p++) {
        while (fi < c->nfeat && c->feat[fi].end + 1000 < p) fi++;
        double l = lam;
        for (int k = fi; k < c->nfeat && c->feat[k].start <= p; k++)
            if (c->feat[k].type == F_SPIKE && p >= c->feat[k].start - L / 2 && p < c->feat[k].end) l = lam * 4;
        int n = poisson(&r, l);
        for (int j = 0; j < n; j++) {
            rd_mods m = { -1, -1, 0, 0, 0, 0 };
            for (int k = fi; k < c->nfeat && c->feat[k].start < p + L + 40; k++) {
                const feature *f = &c->feat[k];
                if (f->end <= p) continue;
                switch (f->type) {
                case F_CLIP: if (f->start > p + 20 && (rng_u32(&r) & 1)) m.clip_at = f->start; break;
                case F_MULTIINS: if (f->start > p + 5 && rng_u32(&r) % 10 < 7) { m.ins_at = f->start; m.ins_len = 1 + (int)(rng_u32(&r) % 5); } break;
                case F_LOWSPAN:
                    if (f->start > p + 20) { if (rng_u32(&r) % 10 < 6) m.clip_at = f->start; else { m.ins_at = f->start; m.ins_len = 2; } }
                    break;
                case F_PARALOG: if (rng_u32(&r) % 10 < 6) m.paralog = 1 + (int)(rng_u32(&r) & 1); break;
                case F_LOWMQ: m.lowmq = 1; break;
                case F_REPEAT: m.repeat = 1; break;
                default: break;
                }
            }
            int h = (int)(rng_u32(&r) & 1), reverse = (int)(rng_u32(&r) & 1);
            if (!sample_read(c, cfg, &r, p, h, reverse, &m, &rd)) continue;
            int mapq = pick_mapq(&r, &m);
            int insert = (int)(350 + 50 * rng_norm(&r)); if (insert < L) insert = L;
            int first = (int)(rng_u32(&r) & 1);
            int flag = 1 | 2 | (reverse ? 16 : 32) | (first ? 64 : 128);
            if (rng_u32(&r) % 100 == 0) flag |= 1024;
            int mpos = reverse ? p - (insert - L) : p + (insert - L);
            if (mpos < 0) mpos = 0;
            snprintf(name, sizeof name, "s%d.%d.%u", c->tid, chunk_id, serial++);
            if (rng_u32(&r) % 200 == 0) {
                /* mate-unmapped style placed read: FUNMAP, no CIGAR, passes straight through */
                emit_record(o, c->tid, p, 0, 1 | 4 | (first ? 64 : 128), &rd, c->tid, p, 0, name, 0);
                continue;
            }
            emit_record(o, c->tid, p, mapq, flag, &rd, c->tid, mpos, reverse ? -insert : insert, name, 1);
        }
    }
}

/* amplicon panel: reads start at amplicon ends +-2 bp; 1000x */
typedef struct { int pos; int idx; } amp_read;
static int cmp_amp(const void *a, const void *b) {
    const amp_read *x = (const amp_read *)a, *y = (const amp_read *)b;
    return x->pos != y->pos ? x->pos - y->pos : x->idx - y->idx;
}
static void gen_amplicon(const contig *c, const simgen_cfg *cfg, int a, int start, obuf *o) {
    rng_t r; rng_seed(&r, cfg->seed ^ 0xA3591, ((uint64_t)c->tid << 32) | (uint32_t)a);
    const int L = cfg->read_len, AL = cfg->amplicon_len;
    int n = cfg->amplicon_depth * AL / L;
    amp_read *ar = (amp_read *)malloc(sizeof(amp_read) * (size_t)n);
    for (int i = 0; i < n; i++) {
        int fwd = i & 1;
        ar[i].pos = (fwd ? start : start + AL - L) + rng_int(&r, -2, 2);
        ar[i].idx = i;
    }
    qsort(ar, (size_t)n, sizeof(amp_read), cmp_amp);
    rd_t rd; char name[64];
    for (int i = 0; i < n; i++) {
        rd_mods m = { -1, -1, 0, 0, 0, 0 };
        int reverse = !(ar[i].idx & 1);
        int h = (int)(rng_u32(&r) & 1);
        if (!sample_read(c, cfg, &r, ar[i].pos, h, reverse, &m, &rd)) continue;
        int mapq = pick_mapq(&r, &m);
        snprintf(name, sizeof name, "a%d.%d.%d", c->tid, a, ar[i].idx);
        emit_record(o, c->tid, ar[i].pos, mapq, 1 | 2 | (reverse ? 16 : 32) | ((ar[i].idx & 2) ? 64 : 128), &rd,
                    c->tid, reverse ? start : start + AL - L, reverse ? -AL : AL, name, 1);
    }
    free(ar);
}

/* ---- driver ---------------------------------------------------------------------- */
typedef struct { const contig *c; const simgen_cfg *cfg; int id, beg, end; int amp; obuf out; } job_t;
typedef struct { job_t *jobs; int njobs; int next; pthread_mutex_t mu; } pool_t;

static void *worker(void *vp) {
    pool_t *pl = (pool_t *)vp;
    for (;;) {
        pthread_mutex_lock(&pl->mu);
        int j = pl->next < pl->njobs ? pl->next++ : -1;
        pthread_mutex_unlock(&pl->mu);
        if (j < 0) break;
        job_t *jb = &pl->jobs[j];
        if (jb->amp) gen_amplicon(jb->c, jb->cfg, jb->id, jb->beg, &jb->out);
        else gen_chunk(jb->c, jb->cfg, jb->id, jb->beg, jb->end, &jb->out);
    }
    return NULL;
}

int simgen_preset(simgen_cfg *cfg, const char *name, double scale, uint64_t seed) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->seed = seed; cfg->n_contigs = 1; cfg->depth = 30; cfg->read_len = 150;
    cfg->features_per_mb = 8; cfg->n_unmapped_tail = 16; cfg->threads = 0;
    if (scale <= 0) scale = 1;
    if (!strcmp(name, "C1")) { cfg->contig_len = (int64_t)(1000000 * scale); cfg->qual_binned = 0; }
    else if (!strcmp(name, "C2") || !strcmp(name, "C3")) { cfg->contig_len = (int64_t)(64000000 * scale); cfg->qual_binned = !strcmp(name, "C2"); }
    else if (!strcmp(name, "C4")) {
        cfg->amplicon = 1; cfg->n_amplicons = (int)(200 * scale); if (cfg->n_amplicons < 1) cfg->n_amplicons = 1;
        cfg->amplicon_len = 250; cfg->amplicon_depth = 1000;
        cfg->contig_len = (int64_t)cfg->n_amplicons * 10000 + 20000; cfg->features_per_mb = 0;
    }
    else if (!strcmp(name, "C5")) { cfg->n_contigs = 24; cfg->contig_len = (int64_t)(129000000 * scale); cfg->qual_binned = 1; cfg->human_like = 1; }
    else if (!strcmp(name, "tiny")) { cfg->contig_len = (int64_t)(50000 * scale); cfg->features_per_mb = 60; cfg->n_contigs = 2; }
    else return -1;
    if (cfg->contig_len < 2000) cfg->contig_len = 2000;
    return 0;
}

/* human-like contig lengths (Mb, chr1..22, X, Y: sum 3088), scaled to the configured total */
static int64_t contig_length(const simgen_cfg *cfg, int t) {
    static const int mb[24] = { 248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57 };
    if (!cfg->human_like) return cfg->contig_len;
    int64_t l = (int64_t)((double)cfg->contig_len * cfg->n_contigs * mb[t % 24] / 3088.0);
    return l < 2000 ? 2000 : l;
}

int simgen_generate(const simgen_cfg *cfg, uint8_t **out, size_t *out_len, int64_t *n_reads, int64_t *n_bases) {
    pthread_once(&g_once, init_tables);
    if (cfg->read_len > 400 || cfg->read_len < 30) return -1;
    int nthreads = cfg->threads > 0 ? cfg->threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    obuf all = {0};
    /* header */
    char text[8192]; int tl = snprintf(text, sizeof text, "@HD\tVN:1.6\tSO:coordinate\n");
    for (int t = 0; t < cfg->n_contigs; t++) {
        if (cfg->n_contigs == 1) tl += snprintf(text + tl, sizeof text - (size_t)tl, "@SQ\tSN:chr20\tLN:%lld\n", (long long)contig_length(cfg, t));
        else tl += snprintf(text + tl, sizeof text - (size_t)tl, "@SQ\tSN:chr%d\tLN:%lld\n", t + 1, (long long)contig_length(cfg, t));
    }
    ob_reserve(&all, (size_t)tl + 64 + 64 * (size_t)cfg->n_contigs);
    memcpy(all.p, "BAM\1", 4); put32(all.p + 4, (uint32_t)tl); memcpy(all.p + 8, text, (size_t)tl);
    all.n = 8 + (size_t)tl;
    put32(all.p + all.n, (uint32_t)cfg->n_contigs); all.n += 4;
    for (int t = 0; t < cfg->n_contigs; t++) {
        char nm[32]; int nl = cfg->n_contigs == 1 ? snprintf(nm, sizeof nm, "chr20") : snprintf(nm, sizeof nm, "chr%d", t + 1);
        put32(all.p + all.n, (uint32_t)nl + 1); all.n += 4;
        memcpy(all.p + all.n, nm, (size_t)nl + 1); all.n += (size_t)nl + 1;
        put32(all.p + all.n, (uint32_t)contig_length(cfg, t)); all.n += 4;
    }
    int tail = 1;
    for (int t = 0; t < cfg->n_contigs; t++) {
        const int64_t clen = contig_length(cfg, t);
        contig c; contig_build(&c, cfg, t, clen);
        pool_t pl; memset(&pl, 0, sizeof pl); pthread_mutex_init(&pl.mu, NULL);
        if (cfg->amplicon) {
            pl.njobs = cfg->n_amplicons;
            pl.jobs = (job_t *)calloc((size_t)pl.njobs, sizeof(job_t));
            for (int a = 0; a < pl.njobs; a++) { job_t *j = &pl.jobs[a]; j->c = &c; j->cfg = cfg; j->id = a; j->beg = 10000 + a * 10000; j->amp = 1; }
        } else {
            const int CH = SIMGEN_CHUNK;
            pl.njobs = (int)((clen + CH - 1) / CH);
            pl.jobs = (job_t *)calloc((size_t)pl.njobs, sizeof(job_t));
            for (int k = 0; k < pl.njobs; k++) {
                job_t *j = &pl.jobs[k]; j->c = &c; j->cfg = cfg; j->id = k; j->beg = k * CH;
                int64_t e = (int64_t)(k + 1) * CH, lim = clen - cfg->read_len - 60;
                if (e > lim) e = lim;
                j->end = (int)e; if (j->end < j->beg) j->end = j->beg;
                if (cfg->job_count > 0 && (k < cfg->job_first || k >= cfg->job_first + cfg->job_count)) j->end = j->beg;   /* not this caller's chunk */
            }
            if (cfg->job_count > 0 && (t + 1 < cfg->n_contigs || cfg->job_first + cfg->job_count < pl.njobs)) tail = 0;
        }
        int nt = nthreads < pl.njobs ? nthreads : pl.njobs;
        pthread_t th[256];
        for (int i = 0; i < nt; i++) pthread_create(&th[i], NULL, worker, &pl);
        for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
        size_t tot = 0;
        for (int k = 0; k < pl.njobs; k++) tot += pl.jobs[k].out.n;
        ob_reserve(&all, tot);
        for (int k = 0; k < pl.njobs; k++) {
            memcpy(all.p + all.n, pl.jobs[k].out.p, pl.jobs[k].out.n); all.n += pl.jobs[k].out.n;
            all.reads += pl.jobs[k].out.reads; all.bases += pl.jobs[k].out.bases;
            free(pl.jobs[k].out.p);
        }
        free(pl.jobs); pthread_mutex_destroy(&pl.mu);
        contig_free(&c);
    }
    /* trailing unmapped reads (tid = -1) */
    rng_t r; rng_seed(&r, cfg->seed, 77);
    for (int i = 0; tail && i < cfg->n_unmapped_tail; i++) {
        rd_t rd; rd.ncig = 0; rd.reflen = 0; rd.l = cfg->read_len;
        for (int k = 0; k < rd.l; k++) { rd.base[k] = (uint8_t)(rng_u32(&r) & 3); rd.qual[k] = (uint8_t)qual_draw(&r, k, rd.l, cfg->qual_binned); }
        char name[32]; snprintf(name, sizeof name, "u%d", i);
        emit_record(&all, -1, -1, 0, 4 | 1 | 8 | 64, &rd, -1, -1, 0, name, 0);
    }
    *out = all.p; *out_len = all.n;
    if (n_reads) *n_reads = all.reads;
    if (n_bases) *n_bases = all.bases;
    return 0;
}

void simgen_free(uint8_t *p) { free(p); }

#ifdef SIMGEN_MAIN
/* crumble_simgen <preset> <scale> <seed> <out.ubam> */
int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s <C1|C2|C3|C4|C5|tiny> <scale> <seed> <out.ubam> [threads]\n", argv[0]); return 1; }
    simgen_cfg cfg;
    if (simgen_preset(&cfg, argv[1], atof(argv[2]), strtoull(argv[3], NULL, 10)) < 0) { fprintf(stderr, "unknown preset\n"); return 1; }
    if (argc > 5) cfg.threads = atoi(argv[5]);
    uint8_t *buf; size_t len; int64_t nr, nb;
    if (simgen_generate(&cfg, &buf, &len, &nr, &nb) < 0) return 1;
    FILE *f = fopen(argv[4], "wb");
    if (!f) { perror(argv[4]); return 1; }
    fwrite(buf, 1, len, f); fclose(f);
    fprintf(stderr, "reads=%lld aligned_bases=%lld bytes=%zu\n", (long long)nr, (long long)nb, len);
    simgen_free(buf);
    return 0;
}
#endif
