/*
 * cg_params.c — option defaults, level presets and error strings of the C ABI
 * (include/crumble_gpu.h).  Plain C so that the command line, the device library and the
 * CPU oracle's command line all share one definition of the reference's option values.
 */
#include <string.h>
#include <limits.h>
#include "../../include/crumble_gpu.h"

/* ---- parameters ------------------------------------------------------------------------ */
void cg_params_default(cg_params *p) {          /* snp_score.c:91-147, 2152-2192 */
    memset(p, 0, sizeof(*p));
    p->reduce_qual = 1; p->binary_qual = 0;
    p->iSTR_mul = 1.0; p->iSTR_add = 2; p->sSTR_mul = 0.0; p->sSTR_add = 0;
    p->qlow = 5; p->qcutoff = 25; p->qhigh = 40; p->qcap = 60;
    p->min_mqual = 0;
    p->min_qual_A = 0; p->min_indel_A = 50; p->min_discrep_A = 2.0;
    p->min_qual_B = 70; p->min_indel_B = 125; p->min_discrep_B = 1.5;
    p->indel_fract = 0.0;
    p->clip_perc = 0.2; p->low_mqual_perc = 1.0; p->ins_len_perc = 1.0; p->over_depth = 999.0; p->indel_ov_perc = 0.0;
    p->pblock = 8;
    p->region_tid = -1; p->region_beg = 0; p->region_end = INT_MAX;
}

int cg_params_level(cg_params *p, int level) {  /* snp_score.c:2380-2482 */
    switch (level) {
    case 9: case 8:
        p->pblock = level == 9 ? 8 : 0;
        p->min_qual_B = 70; p->min_indel_B = 125; p->min_discrep_B = 1.5;
        p->low_mqual_perc = 1.0; p->ins_len_perc = 1.0; p->indel_ov_perc = 0.0; p->over_depth = 999.0;
        p->sSTR_mul = 0.0; p->sSTR_add = 0; p->iSTR_mul = 1.0; p->iSTR_add = 2; p->min_mqual = 0;
        return 0;
    case 7:
        p->pblock = 0;
        p->min_qual_B = 75; p->min_indel_B = 150; p->min_discrep_B = 1.0;
        p->low_mqual_perc = 1.0; p->ins_len_perc = 1.0; p->indel_ov_perc = 0.0; p->over_depth = 999.0;
        p->sSTR_mul = 0.0; p->sSTR_add = 0; p->iSTR_mul = 1.1; p->iSTR_add = 2; p->min_mqual = 0;
        return 0;
    case 5: case 3: case 1:
        p->pblock = 0;
        p->min_qual_B = 75; p->min_indel_B = 150; p->min_discrep_B = 1.0;
        p->low_mqual_perc = 0.5; p->ins_len_perc = 0.1; p->indel_ov_perc = 0.5; p->over_depth = 3.0;
        p->sSTR_mul = level == 5 ? 0.0 : 1.0; p->sSTR_add = level == 1 ? 5 : 0;
        p->iSTR_mul = level == 1 ? 2.0 : 1.1; p->iSTR_add = level == 1 ? 1 : 2;
        p->min_mqual = level == 1 ? 5 : 0;
        return 0;
    default:
        return CG_ERR_BAD_ARG;
    }
}

const char *cg_strerror(int code) {
    switch (code) {
    case CG_OK: return "ok";
    case CG_ERR_NO_DEVICE: return "no usable CUDA device (the GPU path is mandatory; there is no CPU fallback)";
    case CG_ERR_CUDA: return "CUDA error";
    case CG_ERR_NOMEM: return "out of memory";
    case CG_ERR_BAD_ARG: return "bad argument";
    case CG_ERR_UNSORTED: return "input is not coordinate sorted";
    case CG_ERR_UNSUPPORTED: return "option not supported by the device path";
    case CG_ERR_OVERFLOW: return "internal device list overflow";
    case CG_ERR_STATE: return "calls made out of order";
    default: return "unknown error";
    }
}

int cg_abi_version(void) { return CG_ABI_VERSION; }

