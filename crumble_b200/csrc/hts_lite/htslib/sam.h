/*
 * hts_lite: a minimal, from-scratch stand-in for the slice of the htslib API that
 * crumble's host code touches.  htslib itself is absent from the
 * build image and from the GPU box, so the product carries this instead and links
 * the real htslib only when a user has one (same function names and struct shapes).
 *
 * Supported containers: SAM text, BAM (BGZF via zlib, or "raw" uncompressed BAM
 * stream beginning with the BAM\1 magic).  CRAM is not supported here.
 *
 * This header is written for this repository; it is NOT a copy of htslib's sam.h.
 * Only the names/semantics that the reference calls are mirrored:
 *   reference call sites: snp_score.c:597,926,1113-1153,1156-1219,1242,2561-2642.
 */
#ifndef HTS_LITE_SAM_H
#define HTS_LITE_SAM_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- CIGAR ---------------------------------------------------------------- */
#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8
#define BAM_CBACK       9

#define BAM_CIGAR_STR   "MIDNSHP=XB"
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf
/* bit0: consumes query, bit1: consumes reference; ops MIDNSHP=XB */
#define BAM_CIGAR_TYPE  0x3C1A7

#define bam_cigar_op(c)     ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c)  ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_type(o)   (BAM_CIGAR_TYPE >> ((o) << 1) & 3)
#define bam_cigar_gen(l, o) ((l) << BAM_CIGAR_SHIFT | (o))

/* ---- flags ---------------------------------------------------------------- */
#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

/* ---- records -------------------------------------------------------------- */
typedef struct {
    int32_t  tid;
    int32_t  pos;
    uint16_t bin;
    uint8_t  qual;        /* mapping quality */
    uint8_t  l_qname;     /* includes NUL and padding */
    uint16_t flag;
    uint8_t  unused1;
    uint8_t  l_extranul;
    uint32_t n_cigar;
    int32_t  l_qseq;
    int32_t  mtid;
    int32_t  mpos;
    int32_t  isize;
} bam1_core_t;

typedef struct {
    bam1_core_t core;
    int      l_data;
    uint32_t m_data;
    uint8_t *data;        /* qname | cigar | seq (4-bit) | qual | aux */
    uint64_t id;
} bam1_t;

#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i)   ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

extern const char          seq_nt16_str[];     /* "=ACMGRSVTWYHKDBN" */
extern const unsigned char seq_nt16_table[256];

typedef struct {
    int32_t   n_targets, ignore_sam_err;
    uint32_t  l_text;
    uint32_t *target_len;
    int8_t   *cigar_tab;
    char    **target_name;
    char     *text;
    void     *sdict;
} bam_hdr_t;

/* ---- files ---------------------------------------------------------------- */
enum htsExactFormat { unknown_format = 0, binary_format, text_format, sam, bam, bai, cram, crai, vcf, bcf };

typedef struct {
    int category;
    enum htsExactFormat format;
    struct { short major, minor; } version;
    int compression;
    short compression_level;
    void *specific;
    int  nthreads;        /* hts_lite extension: parsed from "nthreads=N" */
    int  level;           /* hts_lite extension: parsed from "level=N"; -1 unset */
    int  raw;             /* hts_lite extension: "raw" => uncompressed BAM stream */
} htsFormat;

struct hts_lite_file;
typedef struct hts_lite_file htsFile;
typedef htsFile samFile;

typedef struct { int dummy; } hts_idx_t;
typedef struct {
    int     tid;
    int     beg, end;     /* 0-based half open; reference reads ->beg/->end (snp_score.c:1514-1516) */
    int     finished;
} hts_itr_t;

int      hts_parse_format(htsFormat *opt, const char *str);
int      sam_open_mode(char *mode, const char *fn, const char *format);
samFile *sam_open_format(const char *fn, const char *mode, const htsFormat *fmt);
samFile *sam_open(const char *fn, const char *mode);
int      sam_close(samFile *fp);

bam_hdr_t *sam_hdr_read(samFile *fp);
int        sam_hdr_write(samFile *fp, const bam_hdr_t *h);
void       bam_hdr_destroy(bam_hdr_t *h);
bam_hdr_t *bam_hdr_init(void);
int        bam_name2id(bam_hdr_t *h, const char *ref);

int sam_read1(samFile *fp, bam_hdr_t *h, bam1_t *b);
int sam_write1(samFile *fp, const bam_hdr_t *h, const bam1_t *b);

hts_idx_t *sam_index_load(samFile *fp, const char *fn);
void       hts_idx_destroy(hts_idx_t *idx);
hts_itr_t *sam_itr_querys(const hts_idx_t *idx, bam_hdr_t *hdr, const char *region);
int        sam_itr_next(samFile *fp, hts_itr_t *itr, bam1_t *b);
void       hts_itr_destroy(hts_itr_t *itr);

bam1_t *bam_init1(void);
void    bam_destroy1(bam1_t *b);
bam1_t *bam_dup1(const bam1_t *b);
bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src);
int32_t bam_endpos(const bam1_t *b);

/* ---- hts_lite extensions (not in htslib) ---------------------------------- */
/* Memory-backed files: fn "mem:" + name.  Readers take their bytes from a buffer
 * registered with hts_lite_mem_put(); writers deposit theirs at sam_close() time and
 * they can be fetched with hts_lite_mem_get().  Used to time transcode without I/O. */
int    hts_lite_mem_put(const char *name, const void *buf, size_t len, int take_copy);
const void *hts_lite_mem_get(const char *name, size_t *len);
void   hts_lite_mem_drop(const char *name);
/* wall-clock seconds between the first sam_read1/sam_itr_next on any file and the
 * most recent sam_write1/sam_read1 call (brackets transcode()). */
double hts_lite_io_span_seconds(void);
void   hts_lite_io_span_reset(void);
/* BGZF worker threads per open file (default: online cores, at most 16; readers use at most 4); also -I/-O nthreads=N */
void   hts_lite_set_threads(int n);
/* parse one SAM text line (NUL terminated, no newline) into b; <0 on error */
int    sam_parse_line(char *line, size_t len, bam_hdr_t *h, bam1_t *b);
/* format b as a SAM text line into *buf (realloc'ed), returns length */
size_t sam_format_line(const bam_hdr_t *h, const bam1_t *b, char **buf, size_t *cap);

#ifdef __cplusplus
}
#endif
#endif
