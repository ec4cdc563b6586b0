/*
 * crumble_bed.c — -R keep.bed: load, sort and collapse the regions whose qualities are kept verbatim
 * (reference bed.c:9-103).  The list is handed to the device path as cg_params.bed / nbed.
 *
 * Collapse rule as the reference applies it (bed.c:20-40): after sorting by (tid, start) a region starts a new
 * entry when its tid is larger than the previous REGION's tid or its start lies beyond the previous REGION's end
 * (not the merged entry's end), otherwise it can only extend the current entry's end.  The reference then copies
 * one element from beyond the array (bed.c:37) and counts it; that element is indeterminate memory which never
 * matches a column in practice, and is not reproduced here.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "crumble_host.h"

static int bed_cmp(const void *a, const void *b) {
    const cg_bed_reg *x = (const cg_bed_reg *)a, *y = (const cg_bed_reg *)b;
    if (x->tid != y->tid) return x->tid - y->tid;
    return x->start - y->start;
}

cg_bed_reg *crumble_bed_load(const char *fn, bam_hdr_t *header, int *nreg) {
    FILE *fp = fopen(fn, "r");
    cg_bed_reg *reg = NULL;
    int cap = 0, n = 0;
    char line[8192], chr[8192];
    if (!fp) { perror(fn); return NULL; }
    while (fgets(line, sizeof line, fp)) {
        int start, end, tid;
        if (line[0] == '#' || !strncmp(line, "track", 5) || !strncmp(line, "browser", 7) || line[0] == '\n') continue;   /* bed.c:56-60 */
        if (sscanf(line, "%s %d %d", chr, &start, &end) != 3) { fprintf(stderr, "Malformed bed line: %s", line); goto err; }
        if ((tid = bam_name2id(header, chr)) < 0) { fprintf(stderr, "Unknown reference name: %s\n", chr); goto err; }
        if (n >= cap) {
            cap = cap ? cap * 2 : 1024;
            cg_bed_reg *t = (cg_bed_reg *)realloc(reg, (size_t)cap * sizeof(*reg));
            if (!t) goto err;
            reg = t;
        }
        reg[n].tid = tid; reg[n].start = start; reg[n].end = end; n++;
    }
    fclose(fp);
    if (n) {
        qsort(reg, (size_t)n, sizeof(*reg), bed_cmp);
        int j = 0, last_tid = -1, last_end = -1;
        for (int i = 0; i < n; i++) {
            if (reg[i].tid > last_tid || reg[i].start > last_end) reg[j++] = reg[i];
            else if (reg[i].end > reg[j - 1].end) reg[j - 1].end = reg[i].end;
            last_tid = reg[i].tid; last_end = reg[i].end;
        }
        n = j;
    }
    if (!reg) reg = (cg_bed_reg *)calloc(1, sizeof(*reg));
    *nreg = n;
    return reg;
err:
    free(reg);
    fclose(fp);
    return NULL;
}
