/*
 * ORACLE / TEST INFRASTRUCTURE ONLY — never linked into, called by, or shipped with the
 * product path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may execute this code, and only as the checker.
 *
 * Plain-C, single-threaded CPU restatement of Crumble's hot path: the column loop of
 * transcode() (reference snp_score.c:1336-2029) with its own in-memory pileup, the gap5
 * heterozygous consensus (calculate_consensus_pileup, 533-797), the STR finder
 * (str_finder.c:34-189), mask_LC_regions (1230-1290), the per-base rewrite (1822-1920),
 * tail handling (1926-1975), flush strip + P-block (1090-1100, 803-834).
 *
 * Pinning: this restatement is checked against oracle/_ref (the reference's own sources
 * compiled verbatim against the htslib-shaped shim) by tests/test_oracle.py and against
 * the golden vectors under tests/golden/ that were generated from oracle/_ref.  The
 * pileup semantics themselves come from SURVEY.md §9.2, not from htslib source (absent):
 * parity at the htslib boundary is UNPINNED.
 *
 * Written for this repository in a different shape from the reference (records in
 * arrays, sliding window of active reads, no trees, no linked lists); every block cites
 * the reference lines it follows.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <float.h>
#include <math.h>
#include "crumble_oracle.h"

/* ------------------------------------------------------------------------------------ */
/* tables (consensus_init, snp_score.c:378-489; q2p / mqual_pow 564-574)                  */
static double T_lprior[15], T_pMM[101], T_p__[101], T_p_M[101], T_q2p[101], T_mqpow[256];
static double T_etab[1001], T_etab2[1001];
static int tables_ready = 0;

static void build_tables(void) {
    if (tables_ready) return;
    tables_ready = 1;
    for (int i = 0; i <= 1000; i++) { T_etab[i] = exp((double)(i - 500)); T_etab2[i] = exp((i - 500) / 10.); }
    const double p_het = 1e-6;
    const double hom = (1 - p_het) / 5, het = p_het / 20;
    /* 15 genotype slots: AA AC AG AT A* CC CG CT C* GG GT G* TT T* ** */
    int j = 0;
    for (int a = 0; a < 5; a++)
        for (int b = a; b < 5; b++)
            T_lprior[j++] = (a == b) ? log(hom) : log(het * 2);
    for (int q = 1; q <= 100; q++) {
        double prob = 1 - pow(10, -q / 10.0);
        T_pMM[q] = log(prob / 5);
        T_p__[q] = log((1 - prob) / 20);
        T_p_M[q] = log((exp(T_pMM[q]) + exp(T_p__[q])) / 2);
    }
    T_pMM[0] = T_pMM[1]; T_p__[0] = T_p__[1]; T_p_M[0] = T_p_M[1];
    for (int q = 0; q <= 100; q++) T_q2p[q] = pow(10, -q / 10.0);
    for (int i = 0; i < 255; i++) T_mqpow[i] = 1 - pow(10, -(i / 2 + .05) / 10.0);
    T_mqpow[255] = T_mqpow[10];
}

/* snp_score.c:491-518 */
static double f_exp(double y) {
    if (y >= -50 && y <= 50) return T_etab2[(int)(y * 10) + 500];
    if (y < -500) y = -500;
    if (y > 500) y = 500;
    return T_etab[(int)y + 500];
}
static double f_log2(double v) {
    union { double d; long long i; } u; u.d = v;
    long long x = u.i;
    int e = (int)((x >> 52) & 2047) - 1024;
    x &= ~(2047LL << 52);
    x += 1023LL << 52;
    u.i = x; v = u.d;
    v = ((-1.0f / 3) * v + 2) * v - 2.0f / 3;
    return v + e;
}
#define PHLOG(x) (-3.0103 * f_log2(x))

/* ------------------------------------------------------------------------------------ */
/* CIGAR helpers                                                                         */
#define OP(c)  ((int)((c) & 15))
#define LEN(c) ((int)((c) >> 4))
static int consumes_ref(int op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
static int consumes_qry(int op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }
static int is_match(int op) { return op == 0 || op == 7 || op == 8; }
static int nib(const orec *r, int i) { return (r->seq[i >> 1] >> ((~i & 1) << 2)) & 15; }

/* one pileup entry */
typedef struct { int ri, qpos, indel, is_del, is_refskip, is_head, is_tail; } pcell;

/* advance the per-read cigar cursor to reference position pos and describe the cell */
static void cursor_cell(orec *r, int pos, pcell *c) {
    if (r->k < 0) {
        r->x = r->pos; r->y = 0;
        int k = 0;
        for (; k < r->n_cigar; k++) {
            int op = OP(r->cigar[k]);
            if (consumes_ref(op)) break;
            if (op == 1 || op == 4) r->y += LEN(r->cigar[k]);
        }
        r->k = k;
    } else if (pos - r->x >= LEN(r->cigar[r->k])) {
        int l = LEN(r->cigar[r->k]);
        if (is_match(OP(r->cigar[r->k]))) r->y += l;
        r->x += l;
        int k = r->k + 1;
        for (; k < r->n_cigar; k++) {
            int op = OP(r->cigar[k]);
            if (consumes_ref(op)) break;
            if (op == 1 || op == 4) r->y += LEN(r->cigar[k]);
        }
        r->k = k;
    }
    int op = OP(r->cigar[r->k]), l = LEN(r->cigar[r->k]);
    c->indel = 0; c->is_del = 0; c->is_refskip = 0;
    if (r->x + l - 1 == pos && r->k + 1 < r->n_cigar) {
        int o2 = OP(r->cigar[r->k + 1]), l2 = LEN(r->cigar[r->k + 1]);
        if (o2 == 2) c->indel = -l2;
        else if (o2 == 1) c->indel = l2;
        else if (o2 == 6 && r->k + 2 < r->n_cigar) {
            int s = 0;
            for (int k = r->k + 2; k < r->n_cigar; k++) {
                int o3 = OP(r->cigar[k]);
                if (o3 == 1) s += LEN(r->cigar[k]);
                else if (consumes_ref(o3)) break;
            }
            if (s > 0) c->indel = s;
        }
    }
    if (is_match(op)) c->qpos = r->y + (pos - r->x);
    else { c->is_del = 1; c->qpos = r->y; c->is_refskip = (op == 3); }
    c->is_head = (pos == r->pos);
    c->is_tail = (pos == r->end - 1);
}

/* snp_score.c:1156-1179 */
static int ref_to_query(const orec *r, int pos) {
    int p = r->pos, q = 0;
    for (int i = 0; i < r->n_cigar; i++) {
        int op = OP(r->cigar[i]), l = LEN(r->cigar[i]);
        if (p + (consumes_ref(op) ? l : 0) < pos) {
            if (consumes_qry(op)) q += l;
            if (consumes_ref(op)) p += l;
            continue;
        }
        if (consumes_qry(op)) q += pos - p;
        return q >= 0 ? q : 0;
    }
    return q;
}

/* snp_score.c:1205-1219 */
static int query_to_ref(const orec *r, int qpos) {
    int rp = r->pos, aq = 0;
    for (int k = 0; k < r->n_cigar && aq < qpos; k++) {
        int op = OP(r->cigar[k]), l = LEN(r->cigar[k]);
        if (consumes_ref(op)) rp += (l <= qpos - aq) ? l : qpos - aq;
        if (consumes_qry(op)) aq += l;
    }
    return rp;
}

/* ------------------------------------------------------------------------------------ */
/* consensus for one column (snp_score.c:533-797)                                        */
typedef struct { int call, het_call, het_phred, phred, depth; float discrep; } ocons;

static void column_consensus(const oracle_params *P, int use_mapq, orec *recs, const pcell *cells, int n, ocons *out) {
    static const int code2base[16] = { 5, 0, 1, 5, 2, 5, 5, 5, 3, 5, 5, 5, 5, 5, 5, 5 };
    /* slot index of genotype (a,b), a<=b */
    static const int slot[5][5] = { { 0, 1, 2, 3, 4 }, { 1, 5, 6, 7, 8 }, { 2, 6, 9, 10, 11 }, { 3, 7, 10, 12, 13 }, { 4, 8, 11, 13, 14 } };
    double S[15] = { 0 }, sumsC[6] = { 0 }, sumsE = 0;
    int depth = 0, nN = 0;
    (void)P;
    for (int i = 0; i < n; i++) {
        if (cells[i].is_refskip) continue;
        const orec *r = &recs[cells[i].ri];
        if (!r->l_qseq) continue;
        int base = code2base[nib(r, cells[i].qpos)];
        if (cells[i].is_del) base = 4;
        unsigned char q = r->q_in[cells[i].qpos];
        if (use_mapq) {
            double _p = T_mqpow[q], _m = T_mqpow[r->mapq];
            q = (unsigned char)PHLOG(1 - (_m * _p + (1 - _m) / 4));
        }
        if (q < 1) q = 1;
        double u = T_p__[q], MM = T_pMM[q] - u, hM = T_p_M[q] - u;
        sumsE += T_q2p[q];
        sumsC[base] += 1 - T_q2p[q];
        if (base < 5) {
            for (int o = 0; o < 5; o++) S[slot[base][o]] += (o == base) ? MM : hM;
        } else {
            /* N: every genotype without a pad gets the match term, pad-containing ones the half term */
            for (int a = 0; a < 4; a++)
                for (int b = a; b < 5; b++) S[slot[a][b]] += (b == 4) ? hM : MM;
            nN++;
        }
        depth++;
    }
    double shift = -DBL_MAX, best = -DBL_MAX, best_het = -DBL_MAX, norm[15];
    int call = 0, het = 0;
    for (int j = 0; j < 15; j++) {
        S[j] += T_lprior[j];
        if (shift < S[j]) shift = S[j];
        int is_hom = (j == 0 || j == 5 || j == 9 || j == 12 || j == 14);
        if (is_hom) { if (best < S[j]) { best = S[j]; call = j; } }
        else if (best_het < S[j]) { best_het = S[j]; het = j; }
    }
    const double min_e = DBL_MIN_EXP * log(2) + 1;
    for (int j = 0; j < 15; j++) {
        S[j] -= shift;
        double e = f_exp(S[j]);
        S[j] = (S[j] > min_e) ? e : DBL_MIN;
        norm[j] = 0;
    }
    double t1 = 0, t2 = 0;
    for (int j = 0; j < 15; j++) { norm[j] += t1; norm[14 - j] += t2; t1 += S[j]; t2 += S[14 - j]; }
    if (depth && depth != nN) {
        static const int to_base[15] = { 0, 5, 5, 5, 5, 1, 5, 5, 5, 2, 5, 5, 3, 5, 4 };
        static const int to_pair[15] = { 0, 1, 2, 3, 4, 6, 7, 8, 9, 12, 13, 14, 18, 19, 24 };
        out->depth = depth;
        out->call = to_base[call];
        if (norm[call] == 0) norm[call] = DBL_MIN;
        int ph = PHLOG(norm[call]) + .5;
        out->phred = ph > 255 ? 255 : (ph < 0 ? 0 : ph);
        out->het_call = to_pair[het];
        if (norm[het] == 0) norm[het] = DBL_MIN;
        ph = 3.0103 * (f_log2(S[het]) - f_log2(norm[het])) + .5;
        out->het_phred = ph;
        double m = sumsC[0] + sumsC[1] + sumsC[2] + sumsC[3] + sumsC[4], c;
        if (out->het_phred > 0) c = sumsC[out->het_call % 5] + sumsC[out->het_call / 5];
        else c = sumsC[out->call];
        out->discrep = (m - c) / sqrt(m);
    } else {
        out->call = 5; out->het_call = 0; out->het_phred = 0; out->phred = 0; out->depth = 0; out->discrep = 0;
    }
}

/* ------------------------------------------------------------------------------------ */
/* STR finder on a character window (str_finder.c:34-189), array-backed list             */
typedef struct { int s, e; } rep;
typedef struct { rep v[1024]; int n; } replist;
static int code_of(char ch) { return ch == 'C' || ch == 'c' ? 1 : ch == 'G' || ch == 'g' ? 2 : (ch == 'T' || ch == 't' || ch == 'U' || ch == 'u') ? 3 : 0; }

static void rep_add(replist *L, const char *w, int wl, int at, int period) {
    if (L->n && L->v[L->n - 1].s <= at - 2 * period + 1 && L->v[L->n - 1].e >= at) return;
    int a = at - period + 1, b = at + 1;
    while (b < wl && code_of(w[a]) == code_of(w[b])) { a++; b++; }
    rep r; r.s = at + 1 - 2 * period; r.e = b - 1;
    /* forget earlier repeats that start inside the new one, walking back while they still reach it */
    int keep_to = L->n;
    while (keep_to > 0 && L->v[keep_to - 1].e >= r.s) keep_to--;
    int w2 = keep_to;
    for (int i = keep_to; i < L->n; i++) if (L->v[i].s < r.s) L->v[w2++] = L->v[i];
    L->v[w2++] = r;
    L->n = w2;
}

static void find_repeats(const char *w, int wl, replist *L) {
    unsigned int word = 0; int i = 0, seen = 0;
    L->n = 0;
    for (; i < wl && seen < 15; i++, seen++) {
        word = (word << 2) | (unsigned)code_of(w[i]);
        for (int p = 1; p <= 7; p++) {
            unsigned mask = (1u << (2 * p)) - 1;
            if (seen >= 2 * p - 1 && (word & mask) == ((word >> (2 * p)) & mask)) rep_add(L, w, wl, i, p);
        }
    }
    for (; i < wl; i++) {
        word = (word << 2) | (unsigned)code_of(w[i]);
        for (int p = 8; p >= 1; p--) {
            unsigned mask = p == 16 ? 0xffffffffu : ((1u << (2 * p)) - 1);
            if ((word & mask) == ((word >> (2 * p)) & mask)) { rep_add(L, w, wl, i, p); break; }
        }
    }
}

/* snp_score.c:1230-1290 */
static void widen_over_repeats(const oracle_params *P, int is_indel, const orec *r, int rpos, int *min_pos, int *max_pos) {
    static const char code2char[] = "=ACMGRSVTWYHKDBN";
    char w[512]; replist L;
    int start = rpos - 250 > 0 ? rpos - 250 : 0;
    int end = rpos + 250 < r->l_qseq ? rpos + 250 : r->l_qseq;
    int wl = end - start + 1;
    for (int i = start; i <= end; i++) {
        int code;
        if (i < r->l_qseq) code = nib(r, i);
        else code = (r->l_qseq & 1) ? (r->seq[r->l_qseq >> 1] & 15) : (r->q_in[0] >> 4);   /* one past the end: SURVEY §9.3(1) */
        w[i - start] = code2char[code];
    }
    find_repeats(w, wl, &L);
    int add = is_indel ? P->iSTR_add : P->sSTR_add;
    for (int k = 0; k < L.n; k++) {
        if (!(rpos + add >= L.v[k].s + start && rpos - add <= L.v[k].e + start)) continue;
        int s = query_to_ref(r, L.v[k].s + start), e = query_to_ref(r, L.v[k].e + start);
        if (*min_pos > s) *min_pos = s;
        if (*max_pos < e) *max_pos = e;
    }
}

/* snp_score.c:803-834 */
static void smooth_pblock(const oracle_params *P, unsigned char *q, int len, int level, int qcap) {
    int lo = INT_MAX, hi = INT_MIN, plo = 0, phi = 0, i, j;
    level *= 2;
    for (i = j = 0; i < len; i++) {
        if (lo > q[i]) lo = q[i];
        if (hi < q[i]) hi = q[i];
        if (hi - lo > level || P->preserve_qual[q[i]]) {
            int mid = (plo + phi) / 2;
            if (mid > qcap) mid = qcap;
            memset(q + j, mid, i - j);
            while (i < len && P->preserve_qual[q[i]]) i++;
            if (i < len) lo = hi = q[i];
            j = i;
        }
        plo = lo; phi = hi;
    }
    if (j < len) memset(q + j, (plo + phi) / 2, (i < len ? i : len) - j);
}

/* ------------------------------------------------------------------------------------ */
int oracle_transcode(const oracle_params *P, orec *recs, long n, FILE *bed_fp, char **names,
                     long long counters[ORACLE_N_COUNTERS], oracle_col_cb cb, void *cb_data) {
    build_tables();
    int bin2[256];
    for (int i = 0; i < 256; i++) { bin2[i] = i < P->qcutoff ? P->qlow : P->qhigh; if (P->preserve_qual[i] > 1) bin2[i] = i; }
    const int str_snp = (P->sSTR_add || P->sSTR_mul);
    memset(counters, 0, sizeof(long long) * ORACLE_N_COUNTERS);

    /* pileup eligibility (snp_score.c:1125-1149) and the pileup's capped copy (1325-1332) */
    long *pile = (long *)malloc(sizeof(long) * (size_t)(n + 1)), np = 0;
    for (long i = 0; i < n; i++) {
        orec *r = &recs[i];
        int hasref = 0, span = 0;
        for (int k = 0; k < r->n_cigar; k++) if (consumes_ref(OP(r->cigar[k]))) { hasref = 1; span += LEN(r->cigar[k]); }
        r->in_pileup = r->tid >= 0 && !(r->flag & 4) && hasref;
        r->end = r->pos + (span ? span : 1);
        r->k = -1; r->keep = 0; r->nopblock = 0;
        for (int x = 0; x < r->l_qseq; x++) {
            unsigned char q = r->q_out[x];
            r->q_in[x] = (q > P->qcap && !P->preserve_qual[q]) ? (unsigned char)P->qcap : q;
        }
        if (r->in_pileup) {
            if (np && (recs[pile[np - 1]].tid > r->tid || (recs[pile[np - 1]].tid == r->tid && recs[pile[np - 1]].pos > r->pos))) { free(pile); return -5; }
            pile[np++] = i;
        }
    }

    pcell *cells = NULL; int cells_cap = 0;
    long head = 0, next = 0;            /* window [head,next) of pileup reads that may still be active */
    int tid = -1, pos = 0, last_tid = -2;
    int min_pos = INT_MAX, max_pos = 0, min_pos2 = INT_MAX, max_pos2 = 0;
    long long total_depth = 0, total_col = 0;
    int stop = 0;

    while (!stop) {
        /* choose the next covered column */
        while (head < next && (recs[pile[head]].tid < tid || (recs[pile[head]].tid == tid && recs[pile[head]].end <= pos))) head++;
        int have = 0;
        for (long a = head; a < next; a++) if (recs[pile[a]].tid == tid && recs[pile[a]].end > pos) { have = 1; break; }
        if (!have) {
            if (next >= np) break;
            head = next;
            tid = recs[pile[next]].tid; pos = recs[pile[next]].pos;
        }
        while (next < np && recs[pile[next]].tid == tid && recs[pile[next]].pos <= pos) next++;
        int n_plp = 0;
        for (long a = head; a < next; a++) {
            orec *r = &recs[pile[a]];
            if (r->tid != tid || r->end <= pos) continue;
            if (n_plp == cells_cap) { cells_cap = cells_cap ? cells_cap * 2 : 256; cells = (pcell *)realloc(cells, sizeof(pcell) * (size_t)cells_cap); }
            cells[n_plp].ri = (int)pile[a];
            cursor_cell(r, pos, &cells[n_plp]);
            n_plp++;
        }
        const int col = pos++;           /* pos now names the next candidate column */
        if (!n_plp) continue;

        int preserve = 0, indel = 0, keep_qual = 0, processed = 0;
        unsigned bedbits = 0;
        ocons cA, cB; memset(&cA, 0, sizeof cA); memset(&cB, 0, sizeof cB);
        int nskip = 0;
        for (int i = 0; i < n_plp; i++) nskip += cells[i].is_refskip;
        if (nskip == n_plp) continue;                                           /* 1466-1472 */
        counters[OC_COLUMNS]++;
        if (tid != last_tid) {                                                  /* 1478-1488 */
            last_tid = tid; min_pos = min_pos2 = INT_MAX; max_pos = max_pos2 = 0; total_depth = total_col = 0;
        }
        total_depth += n_plp; total_col++;
        if (n_plp > 20000) { bedbits |= 1; goto finish_column; }                /* 1493-1500 */
        if (col > max_pos2) { min_pos = min_pos2 = INT_MAX; max_pos = max_pos2 = 0; }   /* 1508-1511 */
        if (P->region_tid >= 0) {                                               /* 1513-1518 */
            if (col < P->region_beg) continue;
            if (col >= P->region_end) { stop = 1; break; }
        }
        processed = 1;
        int call1 = 0, call2 = 0, hA = 0, sA = 0, hB = 0, sB = 0;
        if (P->min_qual_A) {
            column_consensus(P, 0, recs, cells, n_plp, &cA);
            if (cA.het_phred > 0) { call1 = 1 << (cA.het_call / 5); call2 = 1 << (cA.het_call % 5); } else call1 = call2 = 1 << cA.call;
            hA = cA.het_phred > 0 ? cA.het_call : cA.call * 6; sA = cA.het_phred > 0 ? cA.het_phred : cA.phred;
        }
        if (P->min_qual_B) {
            column_consensus(P, 1, recs, cells, n_plp, &cB);
            if (cB.het_phred > 0) { call1 = 1 << (cB.het_call / 5); call2 = 1 << (cB.het_call % 5); } else call1 = call2 = 1 << cB.call;
            hB = cB.het_phred > 0 ? cB.het_call : cB.call * 6; sB = cB.het_phred > 0 ? cB.het_phred : cB.phred;
        }
        if (P->min_qual_A && P->min_qual_B && hA != hB) { counters[OC_DIFF]++; preserve = 1; }
        if (P->min_qual_A) {
            if (cA.het_phred > 0) { counters[OC_HET_A]++; if (sA < P->min_qual_A) counters[OC_HET_QUAL_A]++; }
            else { counters[OC_HOM_A]++; if (sA < P->min_qual_A) counters[OC_HOM_QUAL_A]++; }
            if (cA.discrep >= P->min_discrep_A) { counters[OC_DISCREP_A]++; preserve = 1; }
            if (sA < P->min_qual_A) preserve = 1;
        }
        if (P->min_qual_B) {
            if (cB.het_phred > 0) { counters[OC_HET_B]++; if (sB < P->min_qual_B) counters[OC_HET_QUAL_B]++; }
            else { counters[OC_HOM_B]++; if (sB < P->min_qual_B) counters[OC_HOM_QUAL_B]++; }
            if (cB.discrep >= P->min_discrep_B) { counters[OC_DISCREP_B]++; preserve = 1; }
            if (sB < P->min_qual_B) preserve = 1;
        }
        /* read-set heuristics (1658-1688) */
        int had_indel = 0, had_indel_Q = 0, low_mq = 0, indel_cnt = 0;
        for (int i = 0; i < n_plp; i++) {
            low_mq += recs[cells[i].ri].mapq <= P->min_mqual;
            if (cells[i].indel || cells[i].is_del) { had_indel = 1; indel_cnt++; }
        }
        keep_qual = low_mq > P->low_mqual_perc * (n_plp + .01);
        counters[OC_LOW_MQUAL_PERC] += keep_qual;
        if (n_plp * (total_col + 1) > P->over_depth * (total_depth + 1)) { bedbits |= 2; keep_qual = 1; counters[OC_OVER_DEPTH]++; }
        if (total_col > 1024 * 1024) { total_col >>= 1; total_depth >>= 1; }
        /* indels, STR windows (1690-1762) */
        int indel_sz = 0, hist[101], clipped = 0, n_overlap = 0;
        hist[0] = 0;
        const int lowscore = (P->min_qual_A && sA < P->min_indel_A) || (P->min_qual_B && sB < P->min_indel_B);
        for (int i = 0; i < n_plp; i++) {
            const pcell *c = &cells[i];
            orec *r = &recs[c->ri];
            if (c->is_refskip) continue;
            int is_indel = (c->indel || c->is_del);
            if ((c->is_head && c->qpos > 0) || (c->is_tail && c->qpos + 1 < r->l_qseq)) clipped++;
            if (!c->is_tail && !c->is_head) n_overlap++;
            if (!c->is_head && !c->is_tail && (c->indel > 0 || had_indel)) {
                while (indel_sz < c->indel && indel_sz < 100) hist[++indel_sz] = 0;
                if (c->indel >= 0) hist[c->indel < 99 ? c->indel : 99]++;
            }
            if ((is_indel || (str_snp && preserve)) && lowscore) {
                if (is_indel) {
                    had_indel_Q++;
                    int v = abs(c->indel) + c->is_del;
                    if (indel < v) indel = v;
                } else indel = 1;
                if (indel_cnt >= n_plp * P->indel_fract && r->l_qseq > 0)
                    widen_over_repeats(P, is_indel, r, c->qpos + 1, &min_pos, &max_pos);   /* called twice in the reference, idempotent */
                if (min_pos > col) min_pos = col;
                if (max_pos < col) max_pos = col;
                double mul = is_indel ? P->iSTR_mul : P->sSTR_mul; int add = is_indel ? P->iSTR_add : P->sSTR_add;
                double lo = col - (col - min_pos) * mul - add, hi = col + (max_pos - col) * mul + add;
                min_pos2 = (min_pos2 < lo) ? min_pos2 : lo;      /* double -> int truncation, as the MIN/MAX macros do */
                max_pos2 = (max_pos2 > hi) ? max_pos2 : hi;
            }
        }
        if (had_indel) counters[OC_INDEL]++;
        if (had_indel_Q) counters[OC_INDEL_QUAL]++;
        if ((clipped - 1.0) >= P->clip_perc * n_overlap) { bedbits |= 4; keep_qual = 1; counters[OC_CLIP_PERC]++; }   /* 1764-1773 */
        if (indel_sz) {                                                         /* 1777-1819 */
            int q1 = 0, q2 = 0, ov = 0;
            for (int i = 0; i <= indel_sz && i < 100; i++) {
                if (!hist[i]) continue;
                ov += hist[i];
                if (q1 < hist[i]) { q2 = q1; q1 = hist[i]; } else if (q2 < hist[i]) q2 = hist[i];
            }
            if ((ov - q1 - q2) > P->ins_len_perc * (ov + .1)) { bedbits |= 8; keep_qual = 1; counters[OC_INS_LEN_PERC]++; }
            if ((double)ov < P->indel_ov_perc * n_plp) { bedbits |= 16; keep_qual = 1; counters[OC_INDEL_OV_PERC]++; }
        }
        /* per-base rewrite (1822-1920) */
        for (int i = 0; i < n_plp; i++) {
            const pcell *c = &cells[i];
            orec *r = &recs[c->ri];
            if (keep_qual) r->keep = 1;
            if (c->is_head && r->mapq <= P->min_mqual) for (int x = 0; x < r->l_qseq; x++) r->q_out[x] |= 0x80;
            if (!r->l_qseq) continue;
            unsigned char *q = &r->q_out[c->qpos];
            int base = nib(r, c->qpos);
            if (indel) for (int x = ref_to_query(r, min_pos2); x <= c->qpos; x++) r->q_out[x] = r->q_in[x] | 0x80;
            if (min_pos != INT_MAX) *q = r->q_in[c->qpos] | 0x80;
            if (preserve) *q |= 0x80;
            if (!(*q & 0x80)) {
                if (base == call1 || base == call2) *q = (unsigned char)P->qhigh;
                else if (P->reduce_qual) *q = (unsigned char)(P->binary_qual ? bin2[*q] : P->qlow);
            }
        }
finish_column:
        if (bed_fp)
            for (int t = 0; t < 5; t++) if (bedbits >> t & 1) {
                static const char *tag[5] = { "VDEEP", "DEEP", "CLIP", "INDEL_LEN", "INDEL_COVERAGE" };
                fprintf(bed_fp, "%s\t%d\t%d\t%s\n", names[tid], col - 50 > 0 ? col - 50 : 0, col + 50, tag[t]);
            }
        if (cb) {
            oracle_column oc; const ocons *cc = P->min_qual_B ? &cB : &cA;
            oc.tid = tid; oc.pos = col; oc.n_plp = n_plp; oc.call = cc->call; oc.het_call = cc->het_call; oc.het_phred = cc->het_phred;
            oc.phred = cc->phred; oc.discrep = cc->discrep;
            oc.flags = (preserve ? 1u : 0) | (keep_qual ? 2u : 0) | ((processed && min_pos != INT_MAX) ? 4u : 0) | (processed ? 8u : 0) | (bedbits << 8);
            cb(cb_data, &oc);
        }
        /* reads ending here (1926-1975): whole-read keep restores the pileup's copy */
        for (int i = 0; i < n_plp; i++) {
            orec *r = &recs[cells[i].ri];
            if (cells[i].is_tail && r->keep) memcpy(r->q_out, r->q_in, (size_t)r->l_qseq);
        }
    }
    /* flush (1090-1100, 1998-2015): strip the marker bit, P-block */
    for (long i = 0; i < n; i++) {
        orec *r = &recs[i];
        for (int x = 0; x < r->l_qseq; x++) r->q_out[x] &= 0x7f;
        if (P->pblock && !r->nopblock && r->l_qseq) smooth_pblock(P, r->q_out, r->l_qseq, P->pblock, P->qcap);
    }
    free(cells); free(pile);
    return 0;
}
