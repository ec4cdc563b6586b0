/*
 * crumble_opts.c — option surface of the crumble command line: same getopt string, presets,
 * defaults and -v reports as the reference main() (snp_score.c:2056-2540, 2650-2666), plus the
 * host-side aux-tag filter (purge_tags, 989-1054).  No device code is referenced here.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <getopt.h>
#include <unistd.h>
#include <inttypes.h>
#include "htslib/sam.h"
#include "htslib/cram.h"
#include "crumble_host.h"

#define CRUMBLE_VERSION "0.9.1"

void crumble_opts_default(crumble_opts *o) {
    memset(o, 0, sizeof(*o));
    cg_params_default(&o->p);
}

static int tag_list(uint8_t **bm, const char *arg) {            /* parse_aux_list, snp_score.c:2031-2054 */
    if (!*bm) *bm = (uint8_t *)calloc(8192, 1);
    while (strlen(arg) >= 2) {
        unsigned x = (unsigned)(uint8_t)arg[0] << 8 | (uint8_t)arg[1];
        (*bm)[x >> 3] |= (uint8_t)(1u << (x & 7));
        arg += 2;
        if (*arg == ',') arg++;
        else if (*arg != 0) break;
    }
    if (strlen(arg) != 0) {
        fprintf(stderr, "main_samview: Error parsing option, auxiliary tags should be exactly two characters long.\n");
        return -1;
    }
    return 0;
}

void crumble_usage(FILE *fp) {
    if (fp == stderr) { fprintf(fp, "\nSee \"crumble -h\" for usage.\n"); return; }
    fprintf(fp, "Crumble version %s (B200 device path)\n\n", CRUMBLE_VERSION);
    fprintf(fp, "Usage: crumble [options] in-file out-file\n\nOptions:\n"
            "-I fmt(,opt...)   Input format and format-options [auto].\n"
            "-O fmt(,opt...)   Output format and format-options [SAM].\n"
            "-v                Increase verbosity\n"
            "-z                Do not add an @PG SAM header line\n"
            "-c qual_cutoff    In highly confident regions, quality values above/below\n"
            "-l qual_lower         'qual_cutoff' [25] are quantised to 'qual_lower' [5]\n"
            "-u qual_upper         and 'qual_upper' [40] based on agreement to consensus.\n"
            "-U qual_max       The maximum quality cap used in all bases (even if kept [60])\n"
            "-S                Quantise qualities (with -[clu] options) in soft-clips too.\n"
            "-m min_mqual      Keep qualities for seqs with mapping quality <= mqual [0].\n"
            "-L bool           Whether mismatching bases can have qualities lowered [1]\n"
            "-B                If set, replace quals in good regions with low/high [unset]\n"
            "-i STR_mul,add    Adjust indel size by (STR_size+add)*mul [1.0,2]\n"
            "-s STR_mul,add    Adjust SNP size by (STR_size+add)*mul [0.0,0]\n"
            "-r region         Limit input to region chr:pos(-pos) []\n"
            "-R keep.bed       Keep quality in regions contained in the supplied bed []\n"
            "-t tag_list       Comma separated list of aux tags to keep []\n"
            "-T tag_list       Comma separated list of aux tags to discard []\n"
            "-b out.bed        Output suspicious regions to out.bed []\n"
            "-P float          Keep qual if local depth >= [999.0] times deeper than expected\n"
            "-Y float          Fraction of reads with indel to trigger STR analysis [0.00]\n"
            "-C float          Keep if >= [0.20] reads have soft-clipping\n"
            "-M float          Keep if >= [1.00] reads have low mapping quality\n"
            "-Z float          Keep if >= [1.00] indel sizes do not fit bi-modal dist.\n"
            "-V float          Keep if <  [0.00] reads span indel\n"
            "-q/-d/-x          Calling while ignoring mapping quality: min SNP conf [0], min indel conf [50], min discrepancy [2.0]\n"
            "-Q/-D/-X          Calling with mapping quality: min SNP conf [70], min indel conf [125], min discrepancy [1.5]\n"
            "-p int            P-block algorithm; quality values +/- 'int' [8]\n"
            "-e/-f/-g, -E/-F/-G  BD / BI aux tag binary-binning (lower, cutoff, upper)\n"
            "-k qual / -K qual Preserve quality value if any diffs present / regardless of diffs\n"
            "-N                Store entire column when preserved qualities are present\n"
            "-y machine        illumina | pbccs\n"
            "-1,-3,-5,-7,-8,-9 Compression level presets (use as 1st option); -9 is the default.\n");
}

int crumble_parse_args(crumble_opts *o, int argc, char **argv, htsFormat *in_fmt, htsFormat *out_fmt, int *optind_out) {
    int opt;
    cg_params *p = &o->p;
    optind = 1;
    while ((opt = getopt(argc, argv, "I:O:q:d:x:Q:D:X:m:l:u:U:c:i:L:Bs:t:T:hr:b:vC:M:Z:P:V:p:e:f:g:E:F:G:S135789zR:Y:y:k:K:N")) != -1) {
        switch (opt) {
        case 'I': hts_parse_format(in_fmt, optarg); break;
        case 'O': hts_parse_format(out_fmt, optarg); break;
        case 'q': p->min_qual_A = atoi(optarg); break;
        case 'd': p->min_indel_A = atoi(optarg); break;
        case 'x': p->min_discrep_A = atof(optarg); break;
        case 'Q': p->min_qual_B = atoi(optarg); break;
        case 'D': p->min_indel_B = atoi(optarg); break;
        case 'X': p->min_discrep_B = atof(optarg); break;
        case 'm': p->min_mqual = atoi(optarg); break;
        case 'l': p->qlow = atoi(optarg); break;
        case 'u': p->qhigh = atoi(optarg); break;
        case 'c': p->qcutoff = atoi(optarg); break;
        case 'U': p->qcap = atoi(optarg); break;
        case 'i': p->iSTR_mul = atof(optarg); if (strchr(optarg, ',')) p->iSTR_add = atoi(strchr(optarg, ',') + 1); break;
        case 's': p->sSTR_mul = atof(optarg); if (strchr(optarg, ',')) p->sSTR_add = atoi(strchr(optarg, ',') + 1); break;
        case 'L': p->reduce_qual = atoi(optarg); break;
        case 'B': p->binary_qual = 1; break;
        case 'r': o->region = optarg; break;
        case 'R': o->bed_fn = optarg; break;
        case 't': if (tag_list(&o->aux_whitelist, optarg)) return 1; break;
        case 'T': if (tag_list(&o->aux_blacklist, optarg)) return 1; break;
        case 'b': if (!(o->bed_fp = fopen(optarg, "w"))) { perror(optarg); return 3; } break;
        case 'C': p->clip_perc = atof(optarg); break;
        case 'M': p->low_mqual_perc = atof(optarg); break;
        case 'Z': p->ins_len_perc = atof(optarg); break;
        case 'P': p->over_depth = atof(optarg); break;
        case 'Y': p->indel_fract = atof(optarg); break;
        case 'y':
            if (!strcasecmp(optarg, "illumina")) { }
            else if (!strcasecmp(optarg, "pbccs")) {                    /* snp_score.c:2318-2327 */
                fprintf(stderr, "Using -X0.8 -Y0.1 -m40 -u60 -p16 -k93 -N\n");
                p->indel_fract = 0.1; p->min_discrep_B = 0.8; p->qcutoff = 40; p->qhigh = 60;
                p->pblock = 16; p->perfect_col = 1; p->preserve_qual[93] = 1;
            }
            break;
        case 'V': p->indel_ov_perc = atof(optarg); break;
        case 'p': p->pblock = atoi(optarg); break;
        case 'e': o->BD_low = atoi(optarg) + 33; break;
        case 'f': o->BD_mid = atoi(optarg) + 33; break;
        case 'g': o->BD_high = atoi(optarg) + 33; break;
        case 'E': o->BI_low = atoi(optarg) + 33; break;
        case 'F': o->BI_mid = atoi(optarg) + 33; break;
        case 'G': o->BI_high = atoi(optarg) + 33; break;
        case 'k': case 'K': {                                           /* snp_score.c:2362-2375 */
            char *endp = optarg;
            do {
                long q1 = strtol(endp, &endp, 10), q2 = q1;
                if (*endp == '-') q2 = strtol(endp + 1, &endp, 10);
                do { long q = q1 < 0 ? 0 : (q1 > 255 ? 255 : q1); p->preserve_qual[q] = (uint8_t)(1 + (opt == 'K')); } while (++q1 <= q2);
            } while (*endp++ == ',');
            break;
        }
        case 'N': p->perfect_col = 1; break;
        case '9': case '8': case '7': case '5': case '3': case '1': cg_params_level(p, opt - '0'); break;
        case 'S': p->softclip = 1; break;
        case 'z': p->noPG = 1; break;
        case 'v': p->verbose++; break;
        case 'h': return 2;
        default: return 1;
        }
    }
    *optind_out = optind;
    return 0;
}

void crumble_print_params(const crumble_opts *o) {
    const cg_params *p = &o->p;
    printf("--- Crumble v%s: parameters ---\n", CRUMBLE_VERSION);
    printf("reduce qual:   %s\n", p->reduce_qual ? "yes" : "no");
    printf("indel STR mul: %.2f\n", p->iSTR_mul);
    printf("indel STR add: %d\n", p->iSTR_add);
    printf("SNP   STR mul: %.2f\n", p->sSTR_mul);
    printf("SNP   STR add: %d\n", p->sSTR_add);
    if (p->binary_qual) {
        printf("Qual low  1..%d -> %d\n", p->qcutoff - 1, p->qlow);
        printf("Qual high %d..  -> %d\n", p->qcutoff, p->qhigh);
    } else {
        printf("Qual low  %d, used for discrepant bases in high conf call\n", p->qlow);
        printf("Qual high %d, used for matching bases in high conf call\n", p->qhigh);
    }
    printf("Keep if mqual <= %d\n", p->min_mqual);
    if (p->min_qual_A) {
        printf("Calls without mqual, keep qual if:\n");
        printf("  SNP < %d,  indel < %d,  discrep > %.2f\n", p->min_qual_A, p->min_indel_A, p->min_discrep_A);
    } else printf("Calls without mqual: disabled.\n");
    if (p->min_qual_B) {
        printf("Calls with mqual, keep qual if:\n");
        printf("  SNP < %d,  indel < %d,  discrep > %.2f\n", p->min_qual_B, p->min_indel_B, p->min_discrep_B);
    } else printf("Calls with mqual: disabled.\n");
    fprintf(stderr, "Low mqual perc   = %f\n", p->low_mqual_perc);
    fprintf(stderr, "Ins length perc  = %f\n", p->ins_len_perc);
    fprintf(stderr, "indel ov perc    = %f\n", p->indel_ov_perc);
    fprintf(stderr, "overdepth factor = %f\n", p->over_depth);
    fprintf(stderr, "P-block level    = %d\n", p->pblock);
}

void crumble_print_counters(const crumble_opts *o) {
    const int64_t *c = o->counters;
    fprintf(stderr, "\n\n: Counts of positions preserved by option\n");
    fprintf(stderr, "A/B Diff         = %d\n", (int)c[CG_CNT_DIFF]);
    fprintf(stderr, "A/B Indel        = %d / %d\n", (int)c[CG_CNT_INDEL_QUAL], (int)c[CG_CNT_INDEL]);
    fprintf(stderr, "A:  Het          = %d / %d\n", (int)c[CG_CNT_HET_QUAL_A], (int)c[CG_CNT_HET_A]);
    fprintf(stderr, "A:  Hom          = %d / %d\n", (int)c[CG_CNT_HOM_QUAL_A], (int)c[CG_CNT_HOM_A]);
    fprintf(stderr, "A:  Discrep      = %d\n", (int)c[CG_CNT_DISCREP_A]);
    fprintf(stderr, "B:  Het          = %d / %d\n", (int)c[CG_CNT_HET_QUAL_B], (int)c[CG_CNT_HET_B]);
    fprintf(stderr, "B:  Hom          = %d / %d\n", (int)c[CG_CNT_HOM_QUAL_B], (int)c[CG_CNT_HOM_B]);
    fprintf(stderr, "B:  Discrep      = %d\n\n", (int)c[CG_CNT_DISCREP_B]);
    fprintf(stderr, "Columns          = %" PRId64 "\n", c[CG_CNT_COLUMNS]);
    fprintf(stderr, "Low_mqual_perc   = %" PRId64 "\n", c[CG_CNT_LOW_MQUAL_PERC]);
    fprintf(stderr, "Clip_perc        = %" PRId64 "\n", c[CG_CNT_CLIP_PERC]);
    fprintf(stderr, "Ins_len_perc     = %" PRId64 "\n", c[CG_CNT_INS_LEN_PERC]);
    fprintf(stderr, "indel_ov_perc    = %" PRId64 "\n", c[CG_CNT_INDEL_OV_PERC]);
    fprintf(stderr, "count_over_depth = %" PRId64 "\n", c[CG_CNT_OVER_DEPTH]);
}

/* ---- aux tag filtering / BD,BI binarisation (purge_tags, snp_score.c:989-1054) -------------- */
static const uint8_t *aux_skip(const uint8_t *s, const uint8_t *end) {     /* s at the type byte */
    if (s >= end) return NULL;
    uint8_t t = *s++;
    switch (t) {
    case 'A': case 'c': case 'C': return s + 1;
    case 's': case 'S': return s + 2;
    case 'i': case 'I': case 'f': return s + 4;
    case 'd': return s + 8;
    case 'Z': case 'H': while (s < end && *s) s++; return s + 1;
    case 'B': {
        if (s + 5 > end) return NULL;
        uint8_t st = *s++; uint32_t n; memcpy(&n, s, 4); s += 4;
        int sz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
        if (!sz) return NULL;
        return s + (size_t)sz * n;
    }
    default: return NULL;
    }
}

void crumble_purge_tags(const crumble_opts *o, bam1_t *b) {
    const uint8_t *bm = o->aux_whitelist ? o->aux_whitelist : o->aux_blacklist;
    if (bm) {
        int white = o->aux_whitelist != NULL;
        uint8_t *from = bam_get_aux(b), *to = from, *end = b->data + b->l_data;
        while (from + 3 <= end) {
            unsigned x = (unsigned)from[0] << 8 | from[1];
            const uint8_t *nx = aux_skip(from + 2, end);
            if (!nx || nx > end) abort();                                /* snp_score.c:982 */
            int listed = (bm[x >> 3] >> (x & 7)) & 1;
            if (listed == white) { if (to != from) memmove(to, from, (size_t)(nx - from)); to += nx - from; }
            from = (uint8_t *)nx;
        }
        b->l_data = (int)(to - b->data);
    }
    for (int which = 0; which < 2; which++) {
        int lo = which ? o->BI_low : o->BD_low, mid = which ? o->BI_mid : o->BD_mid, hi = which ? o->BI_high : o->BD_high;
        if (!(lo || mid || hi)) continue;
        uint8_t *t = bam_get_aux(b), *end = b->data + b->l_data;
        while (t + 3 <= end) {
            if (t[0] == 'B' && t[1] == (which ? 'I' : 'D')) {
                uint8_t *c = t + 2;                                      /* starts on the type byte, as the reference does */
                while (*++c) *c = (uint8_t)((*c >= mid) ? hi : lo);
            }
            const uint8_t *nx = aux_skip(t + 2, end);
            if (!nx) abort();
            t = (uint8_t *)nx;
        }
    }
}

