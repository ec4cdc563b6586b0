/*
 * ORACLE / TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * <htslib/sam.h> as seen by the reference sources when they are compiled verbatim
 * from /root/reference into oracle/_ref/.  It is the product's hts_lite record/file
 * API plus the pileup iterator (bam_plp_*) that only the CPU reference needs: the
 * GPU path builds its pileup on the device and never calls these.
 *
 * The pileup semantics are restated in plp.c from the description in SURVEY.md §9.2
 * (htslib's source is not available in this environment; htslib version is unpinned
 * by the reference — ax_with_htslib.m4:126,141 — so parity at this boundary is
 * "unpinned").
 */
#ifndef ORACLE_SHIM_SAM_H
#define ORACLE_SHIM_SAM_H
#include "../../../crumble_b200/csrc/hts_lite/htslib/sam.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef union { void *p; int64_t i; double f; } bam_pileup_cd;

typedef struct {
    bam1_t *b;
    int32_t qpos;
    int indel, level;
    uint32_t is_del:1, is_head:1, is_tail:1, is_refskip:1, aux:28;
    bam_pileup_cd cd;
} bam_pileup1_t;

typedef int (*bam_plp_auto_f)(void *data, bam1_t *b);
struct oracle_plp;
typedef struct oracle_plp *bam_plp_t;

bam_plp_t bam_plp_init(bam_plp_auto_f func, void *data);
void bam_plp_destroy(bam_plp_t iter);
void bam_plp_set_maxcnt(bam_plp_t iter, int maxcnt);
void bam_plp_constructor(bam_plp_t iter, int (*func)(void *data, const bam1_t *b, bam_pileup_cd *cd));
const bam_pileup1_t *bam_plp_auto(bam_plp_t iter, int *_tid, int *_pos, int *_n_plp);

#ifdef __cplusplus
}
#endif
#endif
