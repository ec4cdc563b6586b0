/*
 * cg_column_lean.h — the per-lane bodies of the tuned column kernel (k_column in cg_device.cu), written
 * __host__ __device__ so that tests/emu/ runs the very same code on the CPU against the plain bodies of cg_pipeline.h.
 *
 *   cg_rank_rows        one lane walks its column down the staged cell rows: flag counters and the in-order FP64
 *                       accumulation of calculate_consensus_pileup (snp_score.c:588-686) in RANK space;
 *   cg_cons_from_ranks  "and speculate" (snp_score.c:690-794) straight from the rank-space sums.
 *
 * Rank space: the bases of a column are numbered in order of first appearance; the 15 genotype sums and the 5 discrepancy
 * sums are kept as H[rank], C[rank], P[rank pair].  A read of rank k adds to H[k], C[k] and the four pairs holding k -
 * exactly the six adds of the reference's switch (656-683), in the same order per slot, so every slot sees the same IEEE
 * add sequence.  Nearly all cells of a column carry its first base, so the fast path is taken by whole warps whatever the
 * reference base under each lane is.  sumsE of the reference is dead (never read after the loop) and is not computed.
 * An N base adds to 14 slots; it first hands a rank to every base that has none yet.
 */
#ifndef CG_COLUMN_LEAN_H
#define CG_COLUMN_LEAN_H

#include "cg_cells.h"

typedef struct
#ifdef __CUDACC__
__align__(16)
#endif
ColTabRow { double MM, hM, om, pad; } ColTabRow;       /* per effective quality: pMM-p__, p_M-p__, 1-q2p (snp_score.c:644-651); row 0 is zero */

/* sums of ranks 0 and 1 (and their pairs) live in registers; H2 H3 H4 C2 C3 C4 P23 P24 P34 in `rare` (shared memory on the device) */
#ifndef COL_UNROLL
#define COL_UNROLL 1            /* rows of the column loop per iteration: the kernel is instruction-fetch bound, the smaller loop wins (A/B: 6.96 / 7.29 / 8.73 ms at 1 / 2 / 4) */
#endif
static const int cg_col_unroll = COL_UNROLL;
#define CG_NO_BASE 0x1000u           /* a base without the valid bit: no cell looks like this (an uncovered cell is all zero) */
typedef struct CgRankAcc {
    double H0, H1, C0, C1, P01, P02, P03, P04, P12, P13, P14;
    uint32_t pi, nseen;               /* base -> rank, 4 bits per base, 15 = not seen yet */
    uint32_t b0s, b1s;                /* first and second base of the column as cell bits (CELL_VALID | base << CELL_BASE_SH), CG_NO_BASE = none yet */
    int n_plp, n_skip, n_none, nN, low_mq, n_overlap, indel_cnt, clipped;
    uint32_t ins_seen;
} CgRankAcc;

/* index of a rank-space genotype sum in the 15-entry dump: H[r] = r, P[u<v] = 5 + u(9-u)/2 + v-u-1 */
#define CG_RK_H(r)     (r)
#define CG_RK_P(u, v)  (5 + (((u) * (9 - (u))) >> 1) + ((v) - (u) - 1))
/* where the nine rare sums sit in `rare`: H2 H3 H4 | C2 C3 C4 | P23 P24 P34 */

template <int RS>
CG_HD void cg_rank_init(CgRankAcc *a, double *rare) {
    a->H0 = a->H1 = a->C0 = a->C1 = 0;
    a->P01 = a->P02 = a->P03 = a->P04 = a->P12 = a->P13 = a->P14 = 0;
    a->pi = 0xfffffu; a->nseen = 0; a->b0s = CG_NO_BASE; a->b1s = CG_NO_BASE;
    a->n_plp = a->n_skip = a->n_none = a->nN = a->low_mq = a->n_overlap = a->indel_cnt = a->clipped = 0; a->ins_seen = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < 9; i++) rare[i * RS] = 0;
}

#ifdef __CUDA_ARCH__
/* shared-window address arithmetic: a generic pointer costs several instructions per row to rebuild */
__device__ __forceinline__ void cg_tab_load(uint32_t tab, uint32_t cell, double &mm, double &hm, double &om) {
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(mm), "=d"(hm) : "r"(tab + (cell & CELL_E_M)));
    asm("ld.shared.f64 %0, [%1+16];" : "=d"(om) : "r"(tab + (cell & CELL_E_M)));
}
#endif
CG_HD void cg_tab_load(const ColTabRow *tab, uint32_t cell, double &mm, double &hm, double &om) {
    const ColTabRow *r = (const ColTabRow *)((const char *)tab + (cell & CELL_E_M));
    mm = r->MM; hm = r->hM; om = r->om;
}

/* N: MM to every genotype without a pad, _M to the pad-containing hets, nothing to ** and to the discrepancy sums
 * (snp_score.c:677-682).  Every base needs a rank for that: the unseen ones get theirs now, in base order - ranks are
 * arbitrary labels, the un-permute at the end of the column does not care when they were handed out. */
template <int RS>
CG_HDN void cg_rank_add_N(CgRankAcc *a, double *rare, double mm, double hm) {
    a->nN++;
    int rk[5];
    for (int b = 0; b < 5; b++) {
        uint32_t r = (a->pi >> (4 * b)) & 0xfu;
        if (r == 15u) {
            r = a->nseen++; a->pi = (a->pi & ~(0xfu << (4 * b))) | (r << (4 * b));
            if (r == 0) a->b0s = CELL_VALID | ((uint32_t)b << CELL_BASE_SH);
            if (r == 1) a->b1s = CELL_VALID | ((uint32_t)b << CELL_BASE_SH);
        }
        rk[b] = (int)r;
    }
    for (int x = 0; x < 4; x++) {
        for (int y = x; y < 5; y++) {
            const double v = y == 4 ? hm : mm;
            const int u = rk[x] < rk[y] ? rk[x] : rk[y], w = rk[x] < rk[y] ? rk[y] : rk[x];
            if (x == y) { if (u == 0) a->H0 += v; else if (u == 1) a->H1 += v; else rare[(u - 2) * RS] += v; }
            else switch (CG_RK_P(u, w) - 5) {
                case 0: a->P01 += v; break; case 1: a->P02 += v; break; case 2: a->P03 += v; break; case 3: a->P04 += v; break;
                case 4: a->P12 += v; break; case 5: a->P13 += v; break; case 6: a->P14 += v; break;
                default: rare[(CG_RK_P(u, w) - 5 - 7 + 6) * RS] += v; break;     /* P23 P24 P34 -> rare 6 7 8 */
            }
        }
    }
}

/* third and later bases of a column, N, ref-skip, no-contribution cells: out of line of the hot loop */
template <int RS>
CG_HD void cg_rank_slow(CgRankAcc *a, double *rare, uint32_t cell, double mm, double hm, double om) {
    const uint32_t base = (cell >> CELL_BASE_SH) & 7u;
    if (base < 5u) {
        uint32_t rank = (a->pi >> (base << 2)) & 0xfu;
        if (rank == 15u) {
            rank = a->nseen++; a->pi = (a->pi & ~(0xfu << (base << 2))) | (rank << (base << 2));
            if (rank == 0) a->b0s = cell & (CELL_VALID | CELL_BASE_M);
            if (rank == 1) a->b1s = cell & (CELL_VALID | CELL_BASE_M);
        }
        if (rank == 0)      { a->H0 += mm; a->P01 += hm; a->P02 += hm; a->P03 += hm; a->P04 += hm; a->C0 += om; }
        else if (rank == 1) { a->P01 += hm; a->H1 += mm; a->P12 += hm; a->P13 += hm; a->P14 += hm; a->C1 += om; }
        else {
            rare[(rank - 2) * RS] += mm; rare[(rank + 1) * RS] += om;
            rare[(rank == 4 ? 7 : 6) * RS] += hm; rare[(rank == 2 ? 7 : 8) * RS] += hm;
            if (rank == 2)      { a->P02 += hm; a->P12 += hm; }
            else if (rank == 3) { a->P03 += hm; a->P13 += hm; }
            else                { a->P04 += hm; a->P14 += hm; }
        }
    } else if (base == 5u) {
        /* rare and bulky: out of line, through a copy, so that the accumulators of the hot loop stay in registers */
        CgRankAcc tmp = *a;
        cg_rank_add_N<RS>(&tmp, rare, mm, hm);
        *a = tmp;
    }
    else if (base == 6u) a->n_skip++;
    else a->n_none++;
}

/* the column's first base is (nearly always) the base of its first covering read: look it up before the loop so that the first
 * cell of every column does not take the rank-assignment path */
template <int CS>
CG_HD void cg_rank_peek(CgRankAcc *a, const uint16_t *col, int n) {
    if (a->nseen) return;
    int r = 0; uint32_t cell = 0;
    while (r < n && !((cell = col[r * CS]) & CELL_VALID)) r++;
    if (r < n && ((cell >> CELL_BASE_SH) & 7u) < 5u) {
        const uint32_t base = (cell >> CELL_BASE_SH) & 7u;
        a->b0s = cell & (CELL_VALID | CELL_BASE_M); a->nseen = 1; a->pi = (a->pi & ~(0xfu << (base << 2)));
    }
}

/* rows [0, n) of one column: col[r * CS] is the cell of row r (row n must be readable: the loop looks one row ahead).
 * Flag counters: five 6-bit fields in one word (ins | clip | indel | mid | lowmq), flushed every 32 rows.
 * An uncovered cell is all zero: it counts nothing, carries no flags, and matches neither b0x nor b1x (which include the valid
 * bit), so the common row - every lane on its column's first base, or off the read - runs without a divergent branch. */
template <int CS, int RS, class Tab>
CG_HD void cg_rank_rows(CgRankAcc *a, double *rare, const uint16_t *col, int n, Tab tab) {
    for (int rb = 0; rb < n; rb += 32) {
        const int re = n - rb < 32 ? n - rb : 32;
        uint32_t pk = 0;
        const uint16_t *cp = col + rb * CS;
        uint32_t nxt = cp[0];
#ifdef __CUDA_ARCH__
#pragma unroll cg_col_unroll
#endif
        for (int r = 0; r < re; r++) {
            const uint32_t cell = nxt;
            nxt = cp[(r + 1) * CS];                      /* next row's cell: its latency overlaps this row's arithmetic */
            a->n_plp += (int)(cell >> 15);
            pk += ((cell & 0x1fu) * 0x00108421u) & 0x01041041u;
            double mm, hm, om;
            cg_tab_load(tab, cell, mm, hm, om);
            if (((cell ^ a->b0s) & (CELL_VALID | CELL_BASE_M)) == 0) {
                a->H0 += mm; a->P01 += hm; a->P02 += hm; a->P03 += hm; a->P04 += hm; a->C0 += om;
            } else if (cell & CELL_VALID) {
                if (((cell ^ a->b1s) & (CELL_VALID | CELL_BASE_M)) == 0) { a->P01 += hm; a->H1 += mm; a->P12 += hm; a->P13 += hm; a->P14 += hm; a->C1 += om; }
                else cg_rank_slow<RS>(a, rare, cell, mm, hm, om);
            }
        }
        a->low_mq += pk & 0x3f; a->n_overlap += (pk >> 6) & 0x3f; a->indel_cnt += (pk >> 12) & 0x3f; a->clipped += (pk >> 18) & 0x3f; a->ins_seen |= pk >> 24;
    }
}

/* The un-permute scratch is lane-private (one column of DS-strided doubles) and used twice: first the 15 genotype sums
 * (the three rare H and the three rare P are copied over), then, once those are back in registers, the 5 discrepancy sums. */
template <int DS, int RS>
CG_HD void cg_rank_dump_S(const CgRankAcc *a, const double *rare, double *dump) {
    dump[0 * DS] = a->H0; dump[1 * DS] = a->H1; dump[2 * DS] = rare[0 * RS]; dump[3 * DS] = rare[1 * RS]; dump[4 * DS] = rare[2 * RS];
    dump[5 * DS] = a->P01; dump[6 * DS] = a->P02; dump[7 * DS] = a->P03; dump[8 * DS] = a->P04; dump[9 * DS] = a->P12;
    dump[10 * DS] = a->P13; dump[11 * DS] = a->P14; dump[12 * DS] = rare[6 * RS]; dump[13 * DS] = rare[7 * RS]; dump[14 * DS] = rare[8 * RS];
}
template <int DS, int RS>
CG_HD void cg_rank_dump_C(const CgRankAcc *a, const double *rare, double *dump) {
    dump[0 * DS] = a->C0; dump[1 * DS] = a->C1; dump[2 * DS] = rare[3 * RS]; dump[3 * DS] = rare[4 * RS]; dump[4 * DS] = rare[5 * RS];
}

/* base -> rank for all five bases; unseen bases take the unused ranks (their sums are zero / pure) */
CG_HD void cg_rank_of_bases(uint32_t pi, uint32_t nseen, int rk[5]) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int b = 0; b < 5; b++) { uint32_t r = (pi >> (4 * b)) & 0xfu; if (r == 15u) r = nseen++; rk[b] = (int)r; }
}

/* undo the rank permutation: the accumulators of the plain body (slot order AA AC AG AT A* CC CG CT C* GG GT G* TT T* **) */
template <int DS>
CG_HD void cg_rank_unpermute_S(const double *dump, const int rk[5], double S[15]) {
    int s = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int a = 0; a < 5; a++) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int b = a; b < 5; b++, s++) {
            if (a == b) S[s] = dump[CG_RK_H(rk[a]) * DS];
            else {
                const int u = rk[a] < rk[b] ? rk[a] : rk[b], v = rk[a] < rk[b] ? rk[b] : rk[a];
                S[s] = dump[CG_RK_P(u, v) * DS];
            }
        }
    }
}
template <int DS>
CG_HD void cg_rank_unpermute_C(const double *dump, const int rk[5], double C[5]) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int a = 0; a < 5; a++) C[a] = dump[rk[a] * DS];
}

/* exact recomputation of a column's accumulators with the plain code: used for the rare columns holding an N base */
CG_HDN void cg_col_gather_generic(const CgDev *D, int c, int lo, int hi, CgConsAcc *A) {
    const CgTables *T = D->T;
    cg_cons_init(A);
    for (int j = lo; j < hi; j++) {
        const CgRead q = D->rd[j];
        CgCell cell;
        if (!cg_cell(D, &q, c, &cell)) continue;
        if (cell.is_refskip || !q.l_qseq) continue;
        int nib = cg_seq_nib(D, &q, cell.qpos);
        int base = cell.is_del ? 4 : cg_nt16_to_base(nib);
        uint8_t qv = cg_cap_qual(D->qual[CG_OFF(&q) + cell.qpos], &D->P, T);
        cg_cons_add(T, A, base, T->effB[((int)q.mapq << 8) | qv]);
    }
}

/* calculate_consensus_pileup's tail (snp_score.c:690-794) from the un-permuted sums.
 * Same IEEE operation sequence as cg_cons_finalize (cg_core.h), arranged so that nothing is indexed dynamically:
 *   - shift = max over all 15 = max(hom max, het max); both arg-max chains keep the first maximum (strict <);
 *   - norm[j] = (sum of S[0..j-1], left to right) + (sum of S[14..j+1], right to left): only norm[call] and norm[het_call]
 *     are ever used, so the two running sums are captured as they pass those slots instead of being stored for all 15;
 *   - call is one of the hom slots 0 5 9 12 14, het_call one of the other ten. */
CG_HD void cg_cons_finalize_lean(const CgTables *T, const CgConsAcc *a, int depth, int nN, CgCons *o) {
    if (!depth || depth == nN) {                            /* snp_score.c:751: nothing but N bases counts as no depth */ o->call = 5; o->het_call = 0; o->het_phred = 0; o->phred = 0; o->depth = 0; o->discrep = 0; return; }
    double S[15];
    double mx = -DBL_MAX, mx_het = -DBL_MAX;
    int call = 0, het_call = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 0; j < 15; j++) {
        S[j] = a->S[j] + T->lprior15[j];
        if (j != 0 && j != 5 && j != 9 && j != 12 && j != 14) { if (mx_het < S[j]) { mx_het = S[j]; het_call = j; } }
        else { if (mx < S[j]) { mx = S[j]; call = j; } }
    }
    const double shift = mx < mx_het ? mx_het : mx;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 0; j < 15; j++) {
        const double y = S[j] - shift;
        const double e = cg_fast_exp_neg(T, y);         /* y <= 0: S[j] <= shift */
        S[j] = (y > T->min_e_exp) ? e : DBL_MIN;
    }
    double tot = 0, pre_c = 0, pre_h = 0, suf_c = 0, suf_h = 0, shet = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 0; j < 15; j++) {
        if (j == 0 || j == 5 || j == 9 || j == 12 || j == 14) { if (j == call) pre_c = tot; }
        else if (j == het_call) { pre_h = tot; shet = S[j]; }
        tot += S[j];
    }
    tot = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 14; j >= 0; j--) {
        if (j == 0 || j == 5 || j == 9 || j == 12 || j == 14) { if (j == call) suf_c = tot; }
        else if (j == het_call) suf_h = tot;
        tot += S[j];
    }
    /* norm[j] starts at 0 and receives its two parts in loop order (snp_score.c:743-749): 0 + x is exact and the final add commutes */
    double ncall = pre_c + suf_c, nhet = pre_h + suf_h;
    const int cs = call == 0 ? 0 : call == 5 ? 1 : call == 9 ? 2 : call == 12 ? 3 : 4;                     /* map_sing */
    const int ch = het_call < 5 ? het_call : het_call < 9 ? het_call + 1 : het_call < 12 ? het_call + 3 : het_call + 6;   /* map_het: 1 2 3 4 | 6->7.. */
    o->depth = depth;
    o->call = cs;
    if (ncall == 0) ncall = DBL_MIN;
    int ph = (int)(cg_ph_log(T, ncall) + .5);
    o->phred = ph > 255 ? 255 : (ph < 0 ? 0 : ph);
    o->het_call = ch;
    if (nhet == 0) nhet = DBL_MIN;
    ph = (int)(CG_TENLOG2OVERLOG10 * (cg_fast_log2(T, shet) - cg_fast_log2(T, nhet)) + .5);
    o->het_phred = ph;
    const double m = a->sumsC[0] + a->sumsC[1] + a->sumsC[2] + a->sumsC[3] + a->sumsC[4];
    double c;
    if (ph > 0) {
        double c1 = 0, c2 = 0; const int h1 = ch % 5, h2 = ch / 5;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int j = 0; j < 5; j++) { if (j == h1) c1 = a->sumsC[j]; if (j == h2) c2 = a->sumsC[j]; }
        c = c1 + c2;
    } else {
        c = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int j = 0; j < 5; j++) if (j == cs) c = a->sumsC[j];
    }
    o->discrep = (float)((m - c) / sqrt(m));
}

#endif /* CG_COLUMN_LEAN_H */
