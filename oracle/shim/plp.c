/*
 * ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the htslib pileup iterator semantics that the reference's
 * transcode() relies on (call sites snp_score.c:1427-1437, 2017).  Restated from
 * SURVEY.md §9.2; "parity unpinned" at this boundary (htslib absent, version
 * unpinned by the reference, and the reference ships no tests).
 *
 * Semantics kept:
 *  - records with tid < 0 or BAM_FUNMAP never enter the buffer; nothing else is
 *    filtered (dup/secondary/qcfail stay); maxcnt is honoured only as "no limit";
 *  - the constructor hook is applied to the buffer's private copy of the record;
 *  - a column (tid,pos) is emitted once a record starting beyond it (or EOF) has
 *    been seen; zero-coverage positions are skipped; reads are listed in arrival order;
 *  - per (read, column): qpos / is_del / is_refskip / indel / is_head / is_tail as
 *    in the state machine of SURVEY.md §9.2.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <limits.h>
#include <htslib/sam.h>

typedef struct node {
    bam1_t b;
    int beg, end;            /* [beg, end) on the reference */
    int k, x, y;             /* current ref-consuming op, its ref start, its query start */
    int last;                /* end - 1 */
    bam_pileup_cd cd;
    struct node *next;
} node;

struct oracle_plp {
    bam_plp_auto_f func; void *data;
    int (*construct)(void *, const bam1_t *, bam_pileup_cd *);
    node *head, *tail;       /* arrival-ordered singly linked list */
    node *freelist;          /* nodes dropped at the previous call (their bam data stays valid until now) */
    bam1_t *rb;              /* read buffer handed to func */
    int tid, pos;            /* next column to consider */
    int max_tid, max_pos;    /* start of the last record pushed */
    int is_eof, error;
    bam_pileup1_t *plp; int m_plp;
    int maxcnt;
};

bam_plp_t bam_plp_init(bam_plp_auto_f func, void *data) {
    bam_plp_t it = (bam_plp_t)calloc(1, sizeof(*it));
    it->func = func; it->data = data;
    it->rb = bam_init1();
    it->max_tid = it->max_pos = -1;
    it->maxcnt = 8000;
    return it;
}
void bam_plp_set_maxcnt(bam_plp_t it, int maxcnt) { it->maxcnt = maxcnt; }
void bam_plp_constructor(bam_plp_t it, int (*func)(void *, const bam1_t *, bam_pileup_cd *)) { it->construct = func; }

static void node_free(node *n) { free(n->b.data); free(n); }

void bam_plp_destroy(bam_plp_t it) {
    if (!it) return;
    for (node *n = it->head; n; ) { node *x = n->next; node_free(n); n = x; }
    for (node *n = it->freelist; n; ) { node *x = n->next; node_free(n); n = x; }
    bam_destroy1(it->rb);
    free(it->plp); free(it);
}

static int is_refop(int op) {
    return op == BAM_CMATCH || op == BAM_CDEL || op == BAM_CREF_SKIP || op == BAM_CEQUAL || op == BAM_CDIFF;
}
static int is_mop(int op) { return op == BAM_CMATCH || op == BAM_CEQUAL || op == BAM_CDIFF; }

static int push(bam_plp_t it, const bam1_t *b) {
    if (b->core.tid < 0) return 0;
    if (b->core.flag & BAM_FUNMAP) return 0;
    int end = bam_endpos(b);
    if (b->core.tid < it->max_tid || (b->core.tid == it->max_tid && b->core.pos < it->max_pos)) {
        fprintf(stderr, "[oracle plp] the input is not sorted\n");
        it->error = 1; return -1;
    }
    it->max_tid = b->core.tid; it->max_pos = b->core.pos;
    if (!(end > it->pos || b->core.tid > it->tid)) return 0;
    node *n = (node *)calloc(1, sizeof(*n));
    bam_copy1(&n->b, b);
    n->beg = b->core.pos; n->end = end; n->last = end - 1; n->k = -1;
    if (it->construct) it->construct(it->data, &n->b, &n->cd);   /* return value is garbage in the reference: ignored */
    if (it->tail) it->tail->next = n; else it->head = n;
    it->tail = n;
    return 0;
}

static void resolve(bam_pileup1_t *p, node *s, int pos) {
    bam1_t *b = &s->b;
    const uint32_t *cg = bam_get_cigar(b);
    int nc = (int)b->core.n_cigar, k;
    if (s->k == -1) {
        s->x = b->core.pos; s->y = 0;
        for (k = 0; k < nc; k++) {
            int op = bam_cigar_op(cg[k]), l = (int)bam_cigar_oplen(cg[k]);
            if (is_refop(op)) break;
            if (op == BAM_CINS || op == BAM_CSOFT_CLIP) s->y += l;
        }
        s->k = k;
    } else {
        int l = (int)bam_cigar_oplen(cg[s->k]);
        if (pos - s->x >= l) {
            if (is_mop(bam_cigar_op(cg[s->k]))) s->y += l;
            s->x += l;
            for (k = s->k + 1; k < nc; k++) {
                int op = bam_cigar_op(cg[k]); l = (int)bam_cigar_oplen(cg[k]);
                if (is_refop(op)) break;
                if (op == BAM_CINS || op == BAM_CSOFT_CLIP) s->y += l;
            }
            s->k = k;
        }
    }
    int op = bam_cigar_op(cg[s->k]), l = (int)bam_cigar_oplen(cg[s->k]);
    p->is_del = p->is_refskip = 0; p->indel = 0;
    if (s->x + l - 1 == pos && s->k + 1 < nc) {
        int op2 = bam_cigar_op(cg[s->k + 1]), l2 = (int)bam_cigar_oplen(cg[s->k + 1]);
        if (op2 == BAM_CDEL) p->indel = -l2;
        else if (op2 == BAM_CINS) p->indel = l2;
        else if (op2 == BAM_CPAD && s->k + 2 < nc) {
            int l3 = 0;
            for (k = s->k + 2; k < nc; k++) {
                op2 = bam_cigar_op(cg[k]); l2 = (int)bam_cigar_oplen(cg[k]);
                if (op2 == BAM_CINS) l3 += l2;
                else if (is_refop(op2)) break;
            }
            if (l3 > 0) p->indel = l3;
        }
    }
    if (is_mop(op)) p->qpos = s->y + (pos - s->x);
    else { p->is_del = 1; p->qpos = s->y; p->is_refskip = (op == BAM_CREF_SKIP); }
    p->is_head = (pos == b->core.pos);
    p->is_tail = (pos == s->last);
}

static const bam_pileup1_t *plp_next(bam_plp_t it, int *_tid, int *_pos, int *_n) {
    *_n = 0;
    if (it->error) { *_n = -1; return NULL; }
    /* nodes dropped by the previous call are only now safe to free */
    for (node *n = it->freelist; n; ) { node *x = n->next; node_free(n); n = x; }
    it->freelist = NULL;
    if (it->is_eof && !it->head) return NULL;
    while (it->is_eof || it->max_tid > it->tid || (it->max_tid == it->tid && it->max_pos > it->pos)) {
        int n_plp = 0;
        node **pp = &it->head, *prev = NULL;
        while (*pp) {
            node *p = *pp;
            if (p->b.core.tid < it->tid || (p->b.core.tid == it->tid && p->end <= it->pos)) {
                *pp = p->next;
                if (it->tail == p) it->tail = prev;
                p->next = it->freelist; it->freelist = p;
            } else {
                if (p->b.core.tid == it->tid && p->beg <= it->pos) {
                    if (n_plp == it->m_plp) {
                        it->m_plp = it->m_plp ? it->m_plp * 2 : 256;
                        it->plp = (bam_pileup1_t *)realloc(it->plp, sizeof(bam_pileup1_t) * (size_t)it->m_plp);
                    }
                    memset(&it->plp[n_plp], 0, sizeof(bam_pileup1_t));
                    it->plp[n_plp].b = &p->b;
                    it->plp[n_plp].cd = p->cd;
                    resolve(&it->plp[n_plp], p, it->pos);
                    n_plp++;
                }
                prev = p;
                pp = &p->next;
            }
        }
        *_n = n_plp; *_tid = it->tid; *_pos = it->pos;
        if (it->head) {
            if (it->tid > it->head->b.core.tid) { it->error = 1; *_n = -1; return NULL; }
            if (it->tid < it->head->b.core.tid) { it->tid = it->head->b.core.tid; it->pos = it->head->beg; }
            else if (it->pos < it->head->beg) it->pos = it->head->beg;
            else ++it->pos;
        } else ++it->pos;
        if (n_plp) return it->plp;
        if (it->is_eof && !it->head) break;
    }
    return NULL;
}

const bam_pileup1_t *bam_plp_auto(bam_plp_t it, int *_tid, int *_pos, int *_n_plp) {
    const bam_pileup1_t *plp;
    if (!it->func || it->error) { *_n_plp = -1; return NULL; }
    for (;;) {
        if ((plp = plp_next(it, _tid, _pos, _n_plp)) != NULL) return plp;
        *_n_plp = 0;
        if (it->is_eof || it->error) return NULL;
        int ret = it->func(it->data, it->rb);
        if (ret >= 0) { if (push(it, it->rb) < 0) { *_n_plp = -1; return NULL; } }
        else { it->is_eof = 1; if (ret < -1) { it->error = ret; *_n_plp = -1; return NULL; } }
    }
}
