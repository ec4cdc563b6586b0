/*
 * ORACLE / TEST INFRASTRUCTURE ONLY.
 * Entry point of oracle/_ref/crumble_ref: the reference's own main() (snp_score.c
 * compiled verbatim with -Dmain=crumble_ref_main) wrapped so that the time spent
 * between the first record read and the last record written — i.e. transcode(),
 * snp_score.c:2626 — can be reported without file-open/close costs.
 *
 *   CRUMBLE_REF_TIMING=1  print "transcode_seconds=<s>" on stderr at exit.
 *   CRUMBLE_REF_PRELOAD=1 input is slurped by hts_lite at open time in any case;
 *                         output is buffered in memory when the output name starts
 *                         with "mem:" (discarding writer).
 */
#include <stdio.h>
#include <stdlib.h>
#include <htslib/sam.h>

int crumble_ref_main(int argc, char **argv);

int main(int argc, char **argv) {
    int r = crumble_ref_main(argc, argv);
    if (getenv("CRUMBLE_REF_TIMING"))
        fprintf(stderr, "transcode_seconds=%.6f\n", hts_lite_io_span_seconds());
    return r;
}
