/*
 * crumble_main.c — main() of the crumble command line on top of the GPU path: open files,
 * @PG line, optional region iterator, transcode_gpu(), exit codes as the reference
 * (snp_score.c:2544-2677); file I/O through hts_lite (or a real htslib).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "htslib/sam.h"
#include "htslib/cram.h"
#include "crumble_host.h"

#define CRUMBLE_VERSION "0.9.1"

int crumble_main(int argc, char **argv) {
    samFile *in, *out = NULL;
    htsFormat in_fmt, out_fmt;
    bam_hdr_t *header;
    hts_itr_t *h_iter = NULL;
    crumble_opts o;
    int oi = 1;
    memset(&in_fmt, 0, sizeof in_fmt); memset(&out_fmt, 0, sizeof out_fmt);
    crumble_opts_default(&o);
    const char *dev = getenv("CRUMBLE_DEVICE");
    if (dev) o.device = atoi(dev);
    int r = crumble_parse_args(&o, argc, argv, &in_fmt, &out_fmt, &oi);
    if (r == 2) { crumble_usage(stdout); return 0; }
    if (r == 3) return 1;
    if (r) { crumble_usage(stderr); return 1; }
    if (o.p.verbose) crumble_print_params(&o);

    const char *fnin = NULL;
    if (oi >= argc) {
        if (!isatty(STDIN_FILENO)) fnin = "-";
        else if (argc != 1) { fprintf(stderr, "Missing input filename.\n"); crumble_usage(stderr); return 1; }
        else { crumble_usage(stdout); return 0; }
    } else fnin = argv[oi++];
    if (!(in = sam_open_format(fnin, "r", &in_fmt))) { perror(fnin); return 1; }
    char mode[8] = "w";
    const char *fnout = oi < argc ? argv[oi++] : "-";
    sam_open_mode(mode + 1, fnout, NULL);
    if (!(out = sam_open_format(fnout, mode, &out_fmt))) { perror("(stdout)"); return 1; }
    if (!(header = sam_hdr_read(in))) { fprintf(stderr, "Failed to read file header\n"); return 1; }
    cg_bed_reg *keep_bed = NULL;
    if (o.bed_fn) {                                                      /* snp_score.c:2580-2586 */
        int nb = 0;
        if (!(keep_bed = crumble_bed_load(o.bed_fn, header, &nb))) return 1;
        o.p.bed = keep_bed; o.p.nbed = nb;
    }
    if (!o.p.noPG) {                                                     /* snp_score.c:2588-2609 */
        SAM_hdr *sh = sam_hdr_parse_(header->text, (int)header->l_text);
        if (!sh) return 1;
        char *arg_list = stringify_argv(argc, argv);
        if (sam_hdr_add_PG(sh, "crumble", "VN", CRUMBLE_VERSION, arg_list ? "CL" : NULL, arg_list ? arg_list : NULL, NULL) != 0) return 1;
        free(header->text);
        if (!(header->text = strdup(sam_hdr_str(sh)))) return 1;
        header->l_text = (uint32_t)sam_hdr_length(sh);
        sam_hdr_free(sh); free(arg_list);
    }
    if (out && sam_hdr_write(out, header) != 0) { fprintf(stderr, "Failed to write file header\n"); return 1; }
    if (o.region) {
        hts_idx_t *idx = sam_index_load(in, fnin);
        h_iter = idx ? sam_itr_querys(idx, header, o.region) : NULL;
        if (!h_iter || !idx) { fprintf(stderr, "Failed to load index and/or parse iterator.\n"); return 1; }
        hts_idx_destroy(idx);
    }
    if (transcode_gpu(&o, in, out, header, h_iter) != 0) { fprintf(stderr, "Error while reducing file\n"); return 1; }
    bam_hdr_destroy(header);
    if (h_iter) hts_itr_destroy(h_iter);
    if (sam_close(in) != 0) { fprintf(stderr, "Error while closing input fd\n"); return 1; }
    if (out && sam_close(out) != 0) { fprintf(stderr, "Error while closing output fd\n"); return 1; }
    if (o.p.verbose) crumble_print_counters(&o);
    if (o.bed_fp) fclose(o.bed_fp);
    free(o.aux_whitelist); free(o.aux_blacklist); free(keep_bed);
    return 0;
}

#ifdef CRUMBLE_CLI_MAIN
int main(int argc, char **argv) { return crumble_main(argc, argv); }
#endif
