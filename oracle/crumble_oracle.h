/*
 * ORACLE / TEST INFRASTRUCTURE ONLY (see crumble_oracle.c).
 * Options not restated here (the device path rejects them too): -k/-K/-N/-y pbccs, -S, -R.
 */
#ifndef CRUMBLE_ORACLE_H
#define CRUMBLE_ORACLE_H
#include <stdio.h>
#include <stdint.h>

typedef struct {
    int reduce_qual, binary_qual, iSTR_add, sSTR_add; double iSTR_mul, sSTR_mul;
    int qlow, qcutoff, qhigh, qcap, min_mqual; double indel_fract;
    int min_qual_A, min_indel_A; double min_discrep_A;
    int min_qual_B, min_indel_B; double min_discrep_B;
    double low_mqual_perc, clip_perc, ins_len_perc, over_depth, indel_ov_perc;
    int pblock, region_tid, region_beg, region_end;
    unsigned char preserve_qual[256];
} oracle_params;

typedef struct {
    int tid, pos, end, flag, mapq, l_qseq, n_cigar;
    const uint32_t *cigar; const uint8_t *seq;
    uint8_t *q_in;        /* scratch, l_qseq bytes: the pileup's capped copy */
    uint8_t *q_out;       /* in: original qualities; out: rewritten qualities */
    int in_pileup, keep, nopblock, k, x, y;
} orec;

typedef struct { int tid, pos, n_plp, call, het_call, het_phred, phred; float discrep; unsigned flags; } oracle_column;
typedef void (*oracle_col_cb)(void *data, const oracle_column *c);

enum { OC_DIFF = 0, OC_INDEL_QUAL, OC_INDEL, OC_HET_QUAL_A, OC_HET_A, OC_HOM_QUAL_A, OC_HOM_A, OC_DISCREP_A,
       OC_HET_QUAL_B, OC_HET_B, OC_HOM_QUAL_B, OC_HOM_B, OC_DISCREP_B, OC_COLUMNS, OC_LOW_MQUAL_PERC, OC_CLIP_PERC,
       OC_INS_LEN_PERC, OC_INDEL_OV_PERC, OC_OVER_DEPTH, ORACLE_N_COUNTERS };

int oracle_transcode(const oracle_params *P, orec *recs, long n, FILE *bed_fp, char **names,
                     long long counters[ORACLE_N_COUNTERS], oracle_col_cb cb, void *cb_data);
#endif
