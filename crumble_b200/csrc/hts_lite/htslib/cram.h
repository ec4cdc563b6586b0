/*
 * hts_lite: the handful of header-editing entry points crumble borrows from
 * htslib's (pre-1.10, internal) CRAM header API to add its @PG line
 * (reference call sites snp_score.c:2590-2608).  Implemented in sam_hdr.c.
 */
#ifndef HTS_LITE_CRAM_H
#define HTS_LITE_CRAM_H
#include "sam.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { char *text; size_t len, cap; } SAM_hdr;
SAM_hdr *sam_hdr_parse_(const char *hdr, int len);
/* variadic list of key,value C strings terminated by a NULL key */
int      sam_hdr_add_PG(SAM_hdr *sh, const char *name, ...);
const char *sam_hdr_str(SAM_hdr *sh);
int      sam_hdr_length(SAM_hdr *sh);
void     sam_hdr_free(SAM_hdr *sh);
char    *stringify_argv(int argc, char *argv[]);
#ifdef __cplusplus
}
#endif
#endif
