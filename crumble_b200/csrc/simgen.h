/*
 * Synthetic aligned-read generator for the benchmark configurations of
 * BASELINE.json / SURVEY.md §8(d): simulated diploid reference + Illumina-like
 * reads, deterministic for a given (config, seed), no network, no files needed.
 *
 * Output is an uncompressed BAM stream (BAM\1 header + records, no BGZF) held in
 * memory; hts_lite reads it directly (so the CPU reference can consume it) and the
 * GPU host batcher packs it into SoA buffers (cg_batch_from_bam()).
 */
#ifndef CRUMBLE_SIMGEN_H
#define CRUMBLE_SIMGEN_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SIMGEN_CHUNK (1 << 18)     /* reads are generated per chunk of this many reference positions, each with its own random stream */
typedef struct {
    uint64_t seed;
    int      n_contigs;          /* contigs are named chr1..chrN (or "chr20" when n_contigs==1 and wgs) */
    int64_t  contig_len;         /* length of every contig */
    double   depth;              /* mean coverage */
    int      read_len;           /* 150 */
    int      qual_binned;        /* 1: NovaSeq-like {2,11,25,37}; 0: continuous 2..41 with 3' decay */
    double   features_per_mb;    /* planted heuristic-triggering features of each kind per Mb (>=5 asked by SURVEY) */
    int      amplicon;           /* 1: amplicon panel mode (C4) */
    int      n_amplicons;        /* 200 */
    int      amplicon_len;       /* 250 */
    int      amplicon_depth;     /* 1000 */
    int      n_unmapped_tail;    /* trailing tid=-1 reads */
    int      threads;            /* generator threads (<=0: all cores) */
    /* whole-genome mode only: generate the reads starting in 256 kb chunks [job_first, job_first + job_count) of every contig
     * (job_count <= 0: all).  The chunks of a contig concatenate to exactly the full stream, so a rank of a multi-GPU run can make
     * just its own region shard; the trailing unmapped reads come with the last chunk only. */
    int      job_first, job_count;
    int      human_like;         /* 1: contig t gets the length of human chromosome t+1 (1..22, X, Y) scaled so that they sum to n_contigs * contig_len */
} simgen_cfg;

/* named presets: "C1" (1 Mb 30x continuous quals), "C2" (64 Mb 30x binned), "C4" (amplicon),
 * "tiny" (50 kb); scale>0 multiplies contig length (e.g. C2 with scale 1/64 is a 1 Mb C2-like) */
int simgen_preset(simgen_cfg *cfg, const char *name, double scale, uint64_t seed);

/* Generates the whole data set as one raw BAM stream. Caller frees *out with simgen_free(). */
int  simgen_generate(const simgen_cfg *cfg, uint8_t **out, size_t *out_len, int64_t *n_reads, int64_t *n_aligned_bases);
void simgen_free(uint8_t *p);

#ifdef __cplusplus
}
#endif
#endif
