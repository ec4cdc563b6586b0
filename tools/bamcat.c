/* tools/bamcat.c — copy records through hts_lite (sam_read1 -> sam_write1): times the I/O layer alone.
 * gcc -O2 -Icrumble_b200/csrc/hts_lite tools/bamcat.c crumble_b200/csrc/hts_lite/sam.c crumble_b200/csrc/hts_lite/sam_hdr.c -lz -lpthread -o /tmp/bamcat
 * usage: bamcat IN OUT [in-format] [out-format]      e.g.  bamcat a.bam b.bam bam,nthreads=4 bam,nthreads=16 */
#include <stdio.h>
#include <string.h>
#include "htslib/sam.h"
int main(int argc, char **argv) {
    if (argc < 3) return 2;
    htsFormat fi, fo; memset(&fi, 0, sizeof fi); memset(&fo, 0, sizeof fo);
    if (argc > 3) hts_parse_format(&fi, argv[3]);
    hts_parse_format(&fo, argc > 4 ? argv[4] : "bam");
    samFile *in = sam_open_format(argv[1], "r", &fi);
    char mode[8] = "w"; sam_open_mode(mode + 1, argv[2], NULL);
    samFile *out = sam_open_format(argv[2], mode, &fo);
    if (!in || !out) return 1;
    bam_hdr_t *h = sam_hdr_read(in);
    if (!h || sam_hdr_write(out, h) != 0) return 1;
    bam1_t *b = bam_init1(); long n = 0;
    while (sam_read1(in, h, b) >= 0) { if (sam_write1(out, h, b) < 0) return 1; n++; }
    bam_destroy1(b);
    if (sam_close(in) != 0 || sam_close(out) != 0) return 1;
    fprintf(stderr, "%ld records\n", n);
    return 0;
}
