/*
 * hts_lite: minimal SAM / BAM reader+writer with an htslib-shaped API.
 * See htslib/sam.h in this directory for scope.  Written from the SAM/BAM format
 * specification (SAMv1.pdf §1.4, §4.2); not derived from htslib source.
 *
 * Design: a file opened for reading is slurped (and BGZF-inflated) into memory in
 * sam_open_format(); a file opened for writing accumulates into a memory buffer
 * that is flushed in large chunks.  That keeps record parsing out of syscalls and
 * lets "mem:" files time transcode() without any I/O.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <ctype.h>
#include <time.h>
#include <zlib.h>
#include <pthread.h>
#include <unistd.h>
#include "htslib/sam.h"

const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
const unsigned char seq_nt16_table[256] = {
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
     1, 2, 4, 8, 15,15,15,15, 15,15,15,15, 15, 0,15,15,
    15, 1,14, 2, 13,15,15, 4, 11,15,15,12, 15, 3,15,15,
    15,15, 5, 6,  8,15, 7, 9, 15,10,15,15, 15,15,15,15,
    15, 1,14, 2, 13,15,15, 4, 11,15,15,12, 15, 3,15,15,
    15,15, 5, 6,  8,15, 7, 9, 15,10,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15
};

/* ------------------------------------------------------------------------- */
/* growable byte buffer                                                       */
typedef struct { uint8_t *p; size_t n, cap; } bbuf;

static int bb_reserve(bbuf *b, size_t extra) {
    if (b->n + extra <= b->cap) return 0;
    size_t nc = b->cap ? b->cap : 4096;
    while (nc < b->n + extra) nc += nc >> 1;
    uint8_t *np = (uint8_t *)realloc(b->p, nc);
    if (!np) return -1;
    b->p = np; b->cap = nc;
    return 0;
}
static int bb_put(bbuf *b, const void *src, size_t len) {
    if (bb_reserve(b, len) < 0) return -1;
    memcpy(b->p + b->n, src, len); b->n += len;
    return 0;
}
static int bb_putc(bbuf *b, int c) {
    if (bb_reserve(b, 1) < 0) return -1;
    b->p[b->n++] = (uint8_t)c;
    return 0;
}
static int bb_puts(bbuf *b, const char *s) { return bb_put(b, s, strlen(s)); }
static int bb_putint(bbuf *b, long long v) {
    char tmp[32]; int n = 0; unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) tmp[n++] = '-';
    if (bb_reserve(b, (size_t)n) < 0) return -1;
    while (n) b->p[b->n++] = (uint8_t)tmp[--n];
    return 0;
}

/* ------------------------------------------------------------------------- */
/* memory file registry                                                       */
typedef struct memfile { char *name; uint8_t *buf; size_t len; int owned; struct memfile *next; } memfile;
static memfile *g_mem = NULL;

static memfile *mem_find(const char *name) {
    for (memfile *m = g_mem; m; m = m->next) if (!strcmp(m->name, name)) return m;
    return NULL;
}
void hts_lite_mem_drop(const char *name) {
    memfile **pp = &g_mem;
    while (*pp) {
        if (!strcmp((*pp)->name, name)) {
            memfile *m = *pp; *pp = m->next;
            if (m->owned) free(m->buf);
            free(m->name); free(m);
            return;
        }
        pp = &(*pp)->next;
    }
}
int hts_lite_mem_put(const char *name, const void *buf, size_t len, int take_copy) {
    hts_lite_mem_drop(name);
    memfile *m = (memfile *)calloc(1, sizeof(*m));
    if (!m) return -1;
    m->name = strdup(name);
    if (take_copy) {
        m->buf = (uint8_t *)malloc(len ? len : 1);
        if (!m->buf) { free(m->name); free(m); return -1; }
        memcpy(m->buf, buf, len);
        m->owned = 1;
    } else {
        m->buf = (uint8_t *)buf;   /* caller keeps it alive, or hands ownership via owned=2 */
        m->owned = 0;
    }
    m->len = len;
    m->next = g_mem; g_mem = m;
    return 0;
}
const void *hts_lite_mem_get(const char *name, size_t *len) {
    memfile *m = mem_find(name);
    if (!m) return NULL;
    if (len) *len = m->len;
    return m->buf;
}

/* ------------------------------------------------------------------------- */
/* fork-join worker pool: BGZF blocks are independent, so a reader inflates and a writer
 * deflates a run of blocks in parallel (one pool per open file; the caller takes part). */
typedef struct ppool {
    pthread_t *th; int nth;
    pthread_mutex_t mu; pthread_cond_t cv_go, cv_done;
    void (*fn)(void *, int); void *arg; int n, next, pending, gen, stop;
} ppool;

static void *ppool_main(void *v) {
    ppool *p = (ppool *)v;
    int seen = 0;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (!p->stop && p->gen == seen) pthread_cond_wait(&p->cv_go, &p->mu);
        if (p->stop) break;
        seen = p->gen;
        while (p->next < p->n) {
            int i = p->next++;
            pthread_mutex_unlock(&p->mu);
            p->fn(p->arg, i);
            pthread_mutex_lock(&p->mu);
            if (--p->pending == 0) pthread_cond_signal(&p->cv_done);
        }
    }
    pthread_mutex_unlock(&p->mu);
    return NULL;
}
static ppool *ppool_create(int nth) {
    ppool *p = (ppool *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    pthread_mutex_init(&p->mu, NULL); pthread_cond_init(&p->cv_go, NULL); pthread_cond_init(&p->cv_done, NULL);
    p->th = (pthread_t *)calloc((size_t)(nth > 0 ? nth : 1), sizeof(pthread_t));
    for (int i = 0; i < nth; i++) { if (pthread_create(&p->th[p->nth], NULL, ppool_main, p) == 0) p->nth++; }
    return p;
}
static void ppool_destroy(ppool *p) {
    if (!p) return;
    pthread_mutex_lock(&p->mu); p->stop = 1; pthread_cond_broadcast(&p->cv_go); pthread_mutex_unlock(&p->mu);
    for (int i = 0; i < p->nth; i++) pthread_join(p->th[i], NULL);
    pthread_mutex_destroy(&p->mu); pthread_cond_destroy(&p->cv_go); pthread_cond_destroy(&p->cv_done);
    free(p->th); free(p);
}
/* fn(arg, i) for i in [0, n); returns when all are done */
static void ppool_run(ppool *p, int n, void (*fn)(void *, int), void *arg) {
    if (!p || p->nth == 0 || n < 2) { for (int i = 0; i < n; i++) fn(arg, i); return; }
    pthread_mutex_lock(&p->mu);
    p->fn = fn; p->arg = arg; p->n = n; p->next = 0; p->pending = n; p->gen++;
    pthread_cond_broadcast(&p->cv_go);
    while (p->next < p->n) {
        int i = p->next++;
        pthread_mutex_unlock(&p->mu);
        fn(arg, i);
        pthread_mutex_lock(&p->mu);
        --p->pending;
    }
    while (p->pending > 0) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
}
static int g_threads = -1;
void hts_lite_set_threads(int n) { g_threads = n; }
static int default_threads(void) {
    if (g_threads >= 0) return g_threads;
    const char *e = getenv("HTS_LITE_THREADS");
    if (e) return atoi(e) > 0 ? atoi(e) : 0;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return (int)(n < 1 ? 1 : (n > 16 ? 16 : n));
}

/* ------------------------------------------------------------------------- */
/* timing span                                                                */
static double g_t_first = 0, g_t_last = 0;
static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline void span_touch(void) {
    /* cheap: only the first call and every 4096th call read the clock; sam_close
       finalises.  (clock_gettime per record would perturb the measurement.) */
    static unsigned n = 0;
    if (g_t_first == 0) { g_t_first = g_t_last = now_s(); return; }
    if ((++n & 4095u) == 0) g_t_last = now_s();
}
double hts_lite_io_span_seconds(void) { return g_t_last - g_t_first; }
void hts_lite_io_span_reset(void) { g_t_first = g_t_last = 0; }

/* ------------------------------------------------------------------------- */
struct hts_lite_file {
    int is_write;
    enum htsExactFormat format;   /* sam or bam */
    int raw;                      /* bam without BGZF framing */
    int level;                    /* deflate level for BGZF */
    /* reading: rbuf[rpos, rlen) is decoded input not yet consumed; a file-backed reader refills it chunk by chunk */
    uint8_t *rbuf; size_t rlen, rpos, rcap; int rbuf_owned;
    FILE *in; int in_eof, in_bgzf; bbuf cbuf;  /* cbuf: compressed bytes read ahead (may end inside a block) */
    int nthreads; ppool *pool;
    /* writing */
    FILE *out; char *mem_name; bbuf wbuf;     /* formatted bytes not yet compressed/flushed */
    bbuf zbuf;                                /* BGZF output staging */
    char *line; size_t line_cap;
    bam_hdr_t *hdr;               /* borrowed: last header read from this file */
};

int hts_parse_format(htsFormat *opt, const char *str) {
    memset(opt, 0, sizeof(*opt));
    opt->level = -1;
    if (!str) return -1;
    const char *comma = strchr(str, ',');
    size_t n = comma ? (size_t)(comma - str) : strlen(str);
    if (n == 3 && !strncasecmp(str, "sam", 3)) opt->format = sam;
    else if (n == 3 && !strncasecmp(str, "bam", 3)) opt->format = bam;
    else if (n == 4 && !strncasecmp(str, "cram", 4)) opt->format = cram;
    else { fprintf(stderr, "hts_lite: unknown format '%.*s'\n", (int)n, str); return -1; }
    while (comma) {
        const char *o = comma + 1;
        comma = strchr(o, ',');
        size_t l = comma ? (size_t)(comma - o) : strlen(o);
        if (l > 9 && !strncmp(o, "nthreads=", 9)) opt->nthreads = atoi(o + 9);
        else if (l > 6 && !strncmp(o, "level=", 6)) opt->level = atoi(o + 6);
        else if (l == 3 && !strncmp(o, "raw", 3)) opt->raw = 1;
        /* other htslib/CRAM options are accepted and ignored */
    }
    return 0;
}

int sam_open_mode(char *mode, const char *fn, const char *format) {
    /* choose the mode suffix from the filename extension, as samtools does */
    const char *ext = format;
    if (!ext) {
        if (!fn) return -1;
        ext = strrchr(fn, '.');
        if (!ext || strchr(ext, '/')) { mode[0] = 0; return -1; }
        ext++;
    }
    if (!strcmp(ext, "bam")) strcpy(mode, "b");
    else if (!strcmp(ext, "cram")) strcpy(mode, "c");
    else if (!strcmp(ext, "ubam") || !strcmp(ext, "rawbam")) strcpy(mode, "bu");
    else if (!strcmp(ext, "sam")) strcpy(mode, "");
    else { mode[0] = 0; return -1; }
    return 0;
}

/* ---- BGZF ------------------------------------------------------------------ */
typedef struct { const uint8_t *c; uint32_t clen, isize; size_t ooff; } zblk;
typedef struct { zblk *b; uint8_t *out; int err; } inf_job;
static void inf_one(void *v, int i) {
    inf_job *j = (inf_job *)v;
    const zblk *k = &j->b[i];
    if (!k->isize) return;
    z_stream zs; memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) { j->err = 1; return; }
    zs.next_in = (Bytef *)k->c; zs.avail_in = k->clen;
    zs.next_out = j->out + k->ooff; zs.avail_out = k->isize;
    int r = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (r != Z_STREAM_END) j->err = 1;
}
/* complete BGZF blocks at the front of in[0,inlen), at most maxb of them: block table, bytes consumed, bytes they
 * inflate to.  Returns -1 on a malformed block. */
static int bgzf_scan(const uint8_t *in, size_t inlen, int maxb, zblk **bv, int *bcap, int *nb, size_t *used, size_t *osize) {
    size_t p = 0, o = 0; int n = 0;
    while (p + 18 <= inlen && n < maxb) {
        if (in[p] != 0x1f || in[p + 1] != 0x8b) return -1;
        unsigned xlen = in[p + 10] | (in[p + 11] << 8);
        size_t x = p + 12, xend = x + xlen;
        if (xend > inlen) break;
        int bsize = -1;
        while (x + 4 <= xend) {
            unsigned slen = in[x + 2] | (in[x + 3] << 8);
            if (in[x] == 'B' && in[x + 1] == 'C' && slen == 2 && x + 6 <= xend) bsize = in[x + 4] | (in[x + 5] << 8);
            x += 4 + slen;
        }
        if (bsize < 0) return -1;
        size_t blk = (size_t)bsize + 1;
        if (blk < 12 + (size_t)xlen + 8) return -1;
        if (p + blk > inlen) break;
        if (n == *bcap) {
            int nc = *bcap ? *bcap * 2 : 256;
            zblk *nv = (zblk *)realloc(*bv, (size_t)nc * sizeof(zblk));
            if (!nv) return -1;
            *bv = nv; *bcap = nc;
        }
        zblk *k = &(*bv)[n++];
        k->c = in + p + 12 + xlen; k->clen = (uint32_t)(blk - 12 - xlen - 8);
        k->isize = in[p + blk - 4] | (in[p + blk - 3] << 8) | (in[p + blk - 2] << 16) | ((uint32_t)in[p + blk - 1] << 24);
        k->ooff = o; o += k->isize;
        p += blk;
    }
    *nb = n; *used = p; *osize = o;
    return 0;
}

/* a whole BGZF stream held in memory (mem: files) */
static int bgzf_inflate_all(ppool *pool, const uint8_t *in, size_t inlen, uint8_t **out, size_t *outlen) {
    zblk *bv = NULL; int bcap = 0, nb = 0; size_t used = 0, osize = 0;
    if (bgzf_scan(in, inlen, INT32_MAX, &bv, &bcap, &nb, &used, &osize) < 0) { free(bv); return -1; }
    uint8_t *o = (uint8_t *)malloc(osize ? osize : 1);
    if (!o) { free(bv); return -1; }
    inf_job j = { bv, o, 0 };
    ppool_run(pool, nb, inf_one, &j);
    free(bv);
    if (j.err) { free(o); return -1; }
    *out = o; *outlen = osize;
    return 0;
}

/* ---- streaming reader: keep [rpos, rlen) and append the next chunk of decoded input ---------- */
#define RCHUNK (8u << 20)
static int rfill(struct hts_lite_file *fp) {
    if (!fp->in || !fp->rbuf_owned) return 0;                /* memory-backed: everything is there already */
    if (fp->in_eof && !(fp->in_bgzf && fp->cbuf.n > 0)) return 0;
    if (fp->rpos > 0) {                                       /* compact */
        memmove(fp->rbuf, fp->rbuf + fp->rpos, fp->rlen - fp->rpos);
        fp->rlen -= fp->rpos; fp->rpos = 0;
    }
    bbuf rb = { fp->rbuf, fp->rlen, fp->rcap };
    int got = 0;
    if (!fp->in_bgzf) {
        if (bb_reserve(&rb, RCHUNK) < 0) return -1;
        size_t r = fread(rb.p + rb.n, 1, RCHUNK, fp->in);
        if (r < RCHUNK) fp->in_eof = 1;
        rb.n += r; got = r > 0;
    } else {
        while (!got) {
            if (!fp->in_eof && fp->cbuf.n < RCHUNK) {
                if (bb_reserve(&fp->cbuf, RCHUNK) < 0) return -1;
                size_t r = fread(fp->cbuf.p + fp->cbuf.n, 1, RCHUNK, fp->in);
                if (r < RCHUNK) fp->in_eof = 1;
                fp->cbuf.n += r;
            }
            zblk *bv = NULL; int bcap = 0, nb = 0; size_t used = 0, osize = 0;
            if (bgzf_scan(fp->cbuf.p, fp->cbuf.n, 512, &bv, &bcap, &nb, &used, &osize) < 0) { free(bv); return -1; }
            if (nb == 0) { free(bv); if (fp->in_eof) { fp->cbuf.n = 0; break; } continue; }
            if (bb_reserve(&rb, osize + 1) < 0) { free(bv); return -1; }
            inf_job j = { bv, rb.p + rb.n, 0 };
            if (!fp->pool && fp->nthreads > 1) fp->pool = ppool_create(fp->nthreads - 1);
            ppool_run(fp->pool, nb, inf_one, &j);
            free(bv);
            if (j.err) return -1;
            rb.n += osize; got = osize > 0;
            memmove(fp->cbuf.p, fp->cbuf.p + used, fp->cbuf.n - used);
            fp->cbuf.n -= used;
            if (!got && fp->in_eof && fp->cbuf.n == 0) break;   /* only empty blocks (EOF marker) were left */
        }
    }
    fp->rbuf = rb.p; fp->rlen = rb.n; fp->rcap = rb.cap;
    return got;
}
/* bytes available at rpos after trying to have at least n there */
static size_t rneed(struct hts_lite_file *fp, size_t n) {
    while (fp->rlen - fp->rpos < n) { if (rfill(fp) <= 0) break; }
    return fp->rlen - fp->rpos;
}

/* one BGZF block (gzip member with the BC extra field) from src[0,len) into o (>= ZSLOT bytes); returns its size or -1 */
#define ZSLOT 66560
static long bgzf_pack(uint8_t *o, const uint8_t *src, size_t len, int level) {
    static const uint8_t hdr[12] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0 };
    memcpy(o, hdr, 12);
    o[12] = 'B'; o[13] = 'C'; o[14] = 2; o[15] = 0;
    z_stream zs; memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    zs.next_in = (Bytef *)src; zs.avail_in = (uInt)len;
    zs.next_out = o + 18; zs.avail_out = ZSLOT - 26;
    int r = deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out;
    deflateEnd(&zs);
    if (r != Z_STREAM_END) return -1;
    size_t blk = 18 + clen + 8;
    if (blk > 65536) return -1;
    o[16] = (uint8_t)((blk - 1) & 0xff); o[17] = (uint8_t)((blk - 1) >> 8);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)len);
    uint8_t *t = o + 18 + clen;
    t[0] = crc & 0xff; t[1] = (crc >> 8) & 0xff; t[2] = (crc >> 16) & 0xff; t[3] = (crc >> 24) & 0xff;
    t[4] = len & 0xff; t[5] = (len >> 8) & 0xff; t[6] = (len >> 16) & 0xff; t[7] = (len >> 24) & 0xff;
    return (long)blk;
}
typedef struct { const uint8_t *src; size_t total; uint8_t *slots; long *size; int level; } def_job;
#define BGZF_PAYLOAD 0xff00
#define WBLOCKS 256           /* blocks compressed per parallel round by a multi-threaded writer */
static void def_one(void *v, int i) {
    def_job *j = (def_job *)v;
    size_t off = (size_t)i * BGZF_PAYLOAD, l = j->total - off; if (l > BGZF_PAYLOAD) l = BGZF_PAYLOAD;
    j->size[i] = bgzf_pack(j->slots + (size_t)i * ZSLOT, j->src + off, l, j->level);
}

static int sink_bytes(struct hts_lite_file *fp, const uint8_t *p, size_t n) {
    if (fp->mem_name) return 0;                 /* kept in zbuf/wbuf until close */
    if (n && fwrite(p, 1, n, fp->out) != n) return -1;
    return 0;
}

/* push completed wbuf content downstream; final: 1 at close (drain + EOF block), 2 = drain only */
static int wflush(struct hts_lite_file *fp, int final) {
    if (fp->format == bam && !fp->raw) {
        /* a multi-threaded writer waits for WBLOCKS payloads and deflates them side by side */
        const size_t round = (fp->nthreads > 1 ? WBLOCKS : 1) * (size_t)BGZF_PAYLOAD;
        size_t off = 0;
        while (fp->wbuf.n - off >= round || (final && fp->wbuf.n > off)) {
            size_t tot = fp->wbuf.n - off; if (tot > round) tot = round;
            int nb = (int)((tot + BGZF_PAYLOAD - 1) / BGZF_PAYLOAD);
            uint8_t *slots = (uint8_t *)malloc((size_t)nb * ZSLOT);
            long *size = (long *)malloc((size_t)nb * sizeof(long));
            if (!slots || !size) { free(slots); free(size); return -1; }
            def_job j = { fp->wbuf.p + off, tot, slots, size, fp->level };
            if (!fp->pool && fp->nthreads > 1) fp->pool = ppool_create(fp->nthreads - 1);
            ppool_run(fp->pool, nb, def_one, &j);
            int bad = 0;
            for (int i = 0; i < nb && !bad; i++) { if (size[i] < 0 || bb_put(&fp->zbuf, slots + (size_t)i * ZSLOT, (size_t)size[i]) < 0) bad = 1; }
            free(slots); free(size);
            if (bad) return -1;
            off += tot;
            if (!fp->mem_name && fp->zbuf.n > (4u << 20)) { if (sink_bytes(fp, fp->zbuf.p, fp->zbuf.n) < 0) return -1; fp->zbuf.n = 0; }
        }
        memmove(fp->wbuf.p, fp->wbuf.p + off, fp->wbuf.n - off);
        fp->wbuf.n -= off;
        if (final == 1) {
            static const uint8_t eof[28] = { 0x1f,0x8b,8,4,0,0,0,0,0,0xff,6,0,'B','C',2,0,0x1b,0,3,0,0,0,0,0,0,0,0,0 };
            if (bb_put(&fp->zbuf, eof, 28) < 0) return -1;
        }
        if (!fp->mem_name && (final || fp->zbuf.n > (4u << 20))) {
            if (sink_bytes(fp, fp->zbuf.p, fp->zbuf.n) < 0) return -1;
            fp->zbuf.n = 0;
        }
    } else {
        if (!fp->mem_name && (final || fp->wbuf.n > (4u << 20))) {
            if (sink_bytes(fp, fp->wbuf.p, fp->wbuf.n) < 0) return -1;
            fp->wbuf.n = 0;
        }
    }
    return 0;
}

samFile *sam_open_format(const char *fn, const char *mode, const htsFormat *fmt) {
    struct hts_lite_file *fp = (struct hts_lite_file *)calloc(1, sizeof(*fp));
    if (!fp) return NULL;
    fp->level = Z_DEFAULT_COMPRESSION;
    if (strchr(mode, 'w')) {
        fp->is_write = 1;
        fp->format = sam;
        if (strchr(mode, 'b')) fp->format = bam;
        if (strchr(mode, 'c')) { fprintf(stderr, "hts_lite: CRAM output needs the real htslib\n"); free(fp); errno = ENOTSUP; return NULL; }
        if (strchr(mode, 'u')) fp->raw = 1;
        for (const char *m = mode; *m; m++) if (isdigit((unsigned char)*m)) fp->level = *m - '0';
        if (fmt && fmt->format != unknown_format) {
            if (fmt->format == cram) { fprintf(stderr, "hts_lite: CRAM output needs the real htslib\n"); free(fp); errno = ENOTSUP; return NULL; }
            fp->format = fmt->format;
            if (fmt->raw) fp->raw = 1;
            if (fmt->level >= 0) fp->level = fmt->level;
        }
        fp->nthreads = (fmt && fmt->nthreads > 0) ? fmt->nthreads : default_threads();
        if (!strncmp(fn, "mem:", 4)) fp->mem_name = strdup(fn + 4);
        else if (!strcmp(fn, "-")) fp->out = stdout;
        else if (!(fp->out = fopen(fn, "wb"))) { free(fp); return NULL; }
        return fp;
    }
    /* read */
    {
        int t = (fmt && fmt->nthreads > 0) ? fmt->nthreads : default_threads();
        fp->nthreads = t > 4 ? 4 : t;                        /* inflate is cheap: a few threads keep up with any writer */
    }
    if (!strncmp(fn, "mem:", 4)) {
        size_t len = 0;
        const uint8_t *buf = (const uint8_t *)hts_lite_mem_get(fn + 4, &len);
        if (!buf) { free(fp); errno = ENOENT; return NULL; }
        fp->rbuf = (uint8_t *)buf; fp->rlen = len; fp->rbuf_owned = 0;
        if (len >= 2 && buf[0] == 0x1f && buf[1] == 0x8b) {
            uint8_t *o; size_t ol;
            ppool *pool = fp->nthreads > 1 ? ppool_create(fp->nthreads - 1) : NULL;
            int r = bgzf_inflate_all(pool, buf, len, &o, &ol);
            ppool_destroy(pool);
            if (r < 0) { fprintf(stderr, "hts_lite: not a valid BGZF stream: %s\n", fn); free(fp); return NULL; }
            fp->rbuf = o; fp->rlen = fp->rcap = ol; fp->rbuf_owned = 1;
        }
    } else {
        fp->in = !strcmp(fn, "-") ? stdin : fopen(fn, "rb");
        if (!fp->in) { free(fp); return NULL; }
        fp->rbuf_owned = 1;
        /* the first bytes decide between BGZF and plain (SAM text, raw BAM) */
        if (bb_reserve(&fp->cbuf, 65536) < 0) { sam_close(fp); return NULL; }
        fp->cbuf.n = fread(fp->cbuf.p, 1, 65536, fp->in);
        if (fp->cbuf.n < 65536) fp->in_eof = 1;
        if (fp->cbuf.n >= 2 && fp->cbuf.p[0] == 0x1f && fp->cbuf.p[1] == 0x8b) fp->in_bgzf = 1;
        else { fp->rbuf = fp->cbuf.p; fp->rlen = fp->cbuf.n; fp->rcap = fp->cbuf.cap; memset(&fp->cbuf, 0, sizeof fp->cbuf); }
        if (fp->in_bgzf && rfill(fp) < 0) { fprintf(stderr, "hts_lite: not a valid BGZF stream: %s\n", fn); sam_close(fp); return NULL; }
    }
    rneed(fp, 4);
    fp->format = (fp->rlen >= 4 && !memcmp(fp->rbuf, "BAM\1", 4)) ? bam : sam;
    if (fp->rlen >= 4 && !memcmp(fp->rbuf, "CRAM", 4)) {
        fprintf(stderr, "hts_lite: CRAM input needs the real htslib\n");
        sam_close(fp); errno = ENOTSUP; return NULL;
    }
    return fp;
}

samFile *sam_open(const char *fn, const char *mode) { return sam_open_format(fn, mode, NULL); }

int sam_close(samFile *fp) {
    int ret = 0;
    if (!fp) return 0;
    if (g_t_first != 0) g_t_last = now_s();
    if (fp->is_write) {
        if (wflush(fp, 1) < 0) ret = -1;
        if (fp->mem_name) {
            bbuf *src = (fp->format == bam && !fp->raw) ? &fp->zbuf : &fp->wbuf;
            hts_lite_mem_put(fp->mem_name, src->p, src->n, 1);
        } else if (fp->out) {
            if (fflush(fp->out) != 0) ret = -1;
            if (fp->out != stdout && fclose(fp->out) != 0) ret = -1;
        }
        free(fp->wbuf.p); free(fp->zbuf.p); free(fp->mem_name);
    } else {
        if (fp->rbuf_owned) free(fp->rbuf);
        if (fp->in && fp->in != stdin) fclose(fp->in);
        free(fp->cbuf.p);
    }
    ppool_destroy(fp->pool);
    free(fp->line);
    free(fp);
    return ret;
}

/* ---- header ---------------------------------------------------------------- */
bam_hdr_t *bam_hdr_init(void) { return (bam_hdr_t *)calloc(1, sizeof(bam_hdr_t)); }

void bam_hdr_destroy(bam_hdr_t *h) {
    if (!h) return;
    for (int i = 0; i < h->n_targets; i++) free(h->target_name[i]);
    free(h->target_name); free(h->target_len); free(h->text); free(h->cigar_tab);
    free(h);
}

int bam_name2id(bam_hdr_t *h, const char *ref) {
    for (int i = 0; i < h->n_targets; i++)
        if (!strcmp(h->target_name[i], ref)) return i;
    return -1;
}

static int hdr_add_target(bam_hdr_t *h, const char *name, size_t nl, uint32_t len) {
    int n = h->n_targets;
    if ((n & (n + 1)) == 0 || n == 0) {   /* grow at powers of two minus one */
        int cap = n ? 2 * (n + 1) : 2;
        char **tn = (char **)realloc(h->target_name, sizeof(char *) * (size_t)cap);
        uint32_t *tl = (uint32_t *)realloc(h->target_len, sizeof(uint32_t) * (size_t)cap);
        if (!tn || !tl) return -1;
        h->target_name = tn; h->target_len = tl;
    }
    h->target_name[n] = strndup(name, nl);
    h->target_len[n] = len;
    h->n_targets = n + 1;
    return 0;
}

static int hdr_targets_from_text(bam_hdr_t *h) {
    const char *p = h->text, *end = h->text + h->l_text;
    while (p < end) {
        const char *eol = memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        if (eol - p > 3 && !memcmp(p, "@SQ", 3)) {
            const char *sn = NULL; size_t snl = 0; uint32_t ln = 0;
            const char *f = p + 3;
            while (f < eol) {
                if (*f == '\t') f++;
                const char *fe = memchr(f, '\t', (size_t)(eol - f));
                if (!fe) fe = eol;
                if (fe - f > 3 && f[2] == ':') {
                    if (f[0] == 'S' && f[1] == 'N') { sn = f + 3; snl = (size_t)(fe - sn); }
                    else if (f[0] == 'L' && f[1] == 'N') ln = (uint32_t)strtoul(f + 3, NULL, 10);
                }
                f = fe;
            }
            if (sn && hdr_add_target(h, sn, snl, ln) < 0) return -1;
        }
        p = eol + 1;
    }
    return 0;
}

static inline uint32_t rd_u32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline void wr_u32(uint8_t *p, uint32_t v) { p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = (uint8_t)(v >> 24); }
static inline void wr_u16(uint8_t *p, uint16_t v) { p[0] = v & 0xff; p[1] = (uint8_t)(v >> 8); }

bam_hdr_t *sam_hdr_read(samFile *fp) {
    bam_hdr_t *h = bam_hdr_init();
    if (!h) return NULL;
    if (fp->format == bam) {
        /* have the whole header in the window before parsing it (a streaming reader holds one chunk) */
        if (rneed(fp, 12) < 12) goto fail;
        {
            size_t want = 8 + (size_t)rd_u32(fp->rbuf + fp->rpos + 4) + 4;
            if (rneed(fp, want) < want) goto fail;
            uint32_t nr = rd_u32(fp->rbuf + fp->rpos + want - 4);
            for (uint32_t i = 0; i < nr; i++) {
                if (rneed(fp, want + 4) < want + 4) goto fail;
                want += 4 + (size_t)rd_u32(fp->rbuf + fp->rpos + want) + 4;
                if (rneed(fp, want) < want) goto fail;
            }
        }
        const uint8_t *p = fp->rbuf + 4;
        uint32_t lt = rd_u32(p); p += 4;
        if ((size_t)(p - fp->rbuf) + lt + 4 > fp->rlen) goto fail;
        h->text = (char *)malloc((size_t)lt + 1);
        memcpy(h->text, p, lt); h->text[lt] = 0;
        /* header text may carry trailing NULs */
        h->l_text = (uint32_t)strnlen(h->text, lt);
        p += lt;
        uint32_t nref = rd_u32(p); p += 4;
        for (uint32_t i = 0; i < nref; i++) {
            uint32_t ln = rd_u32(p); p += 4;
            const char *nm = (const char *)p; p += ln;
            uint32_t rl = rd_u32(p); p += 4;
            if (hdr_add_target(h, nm, ln ? ln - 1 : 0, rl) < 0) goto fail;
        }
        fp->rpos = (size_t)(p - fp->rbuf);
    } else {
        size_t p = 0;
        for (;;) {
            if (p >= fp->rlen && rneed(fp, fp->rlen - fp->rpos + 1) <= p) break;
            if (fp->rbuf[p] != '@') break;
            uint8_t *eol;
            while (!(eol = (uint8_t *)memchr(fp->rbuf + p, '\n', fp->rlen - p))) { size_t have = fp->rlen; if (rneed(fp, have + 1) <= have) break; }
            p = eol ? (size_t)(eol - fp->rbuf) + 1 : fp->rlen;
        }
        h->text = (char *)malloc(p + 1);
        memcpy(h->text, fp->rbuf, p); h->text[p] = 0;
        h->l_text = (uint32_t)p;
        fp->rpos = p;
        if (hdr_targets_from_text(h) < 0) goto fail;
    }
    fp->hdr = h;
    return h;
fail:
    bam_hdr_destroy(h);
    return NULL;
}

int sam_hdr_write(samFile *fp, const bam_hdr_t *h) {
    if (fp->format == bam) {
        uint8_t t[4];
        bb_put(&fp->wbuf, "BAM\1", 4);
        wr_u32(t, h->l_text); bb_put(&fp->wbuf, t, 4);
        bb_put(&fp->wbuf, h->text, h->l_text);
        wr_u32(t, (uint32_t)h->n_targets); bb_put(&fp->wbuf, t, 4);
        for (int i = 0; i < h->n_targets; i++) {
            uint32_t ln = (uint32_t)strlen(h->target_name[i]) + 1;
            wr_u32(t, ln); bb_put(&fp->wbuf, t, 4);
            bb_put(&fp->wbuf, h->target_name[i], ln);
            wr_u32(t, h->target_len[i]); bb_put(&fp->wbuf, t, 4);
        }
        /* htslib flushes the header into its own BGZF block(s) */
        if (!fp->raw && wflush(fp, 2) < 0) return -1;
    } else {
        if (bb_put(&fp->wbuf, h->text, h->l_text) < 0) return -1;
        if (h->l_text && h->text[h->l_text - 1] != '\n') bb_putc(&fp->wbuf, '\n');
    }
    return wflush(fp, 0);
}

/* ---- records --------------------------------------------------------------- */
bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if (b) { free(b->data); free(b); } }

static int bam_reserve(bam1_t *b, size_t n) {
    if (n <= b->m_data) return 0;
    size_t m = b->m_data ? b->m_data : 256;
    while (m < n) m += m >> 1;
    /* +8 slack: the reference reads one nibble past the sequence (snp_score.c:1241) */
    uint8_t *d = (uint8_t *)realloc(b->data, m + 8);
    if (!d) return -1;
    b->data = d; b->m_data = (uint32_t)m;
    return 0;
}

bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src) {
    if (bam_reserve(dst, (size_t)src->l_data) < 0) return NULL;
    memcpy(dst->data, src->data, (size_t)src->l_data);
    dst->core = src->core; dst->l_data = src->l_data; dst->id = src->id;
    return dst;
}
bam1_t *bam_dup1(const bam1_t *b) {
    bam1_t *d = bam_init1();
    if (!d) return NULL;
    if (!bam_copy1(d, b)) { bam_destroy1(d); return NULL; }
    return d;
}

int32_t bam_endpos(const bam1_t *b) {
    int32_t e = b->core.pos;
    if (!(b->core.flag & BAM_FUNMAP) && b->core.n_cigar) {
        const uint32_t *c = bam_get_cigar(b);
        for (uint32_t k = 0; k < b->core.n_cigar; k++)
            if (bam_cigar_type(bam_cigar_op(c[k])) & 2) e += (int32_t)bam_cigar_oplen(c[k]);
    }
    if (e == b->core.pos) e++;
    return e;
}

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

static int bam_read_rec(samFile *fp, bam1_t *b) {
    if (fp->rpos + 4 > fp->rlen && rneed(fp, 4) < 4) return -1;
    uint32_t bs = rd_u32(fp->rbuf + fp->rpos);
    if (bs < 32) return -2;
    if (fp->rpos + 4 + bs > fp->rlen && rneed(fp, 4 + (size_t)bs) < 4 + (size_t)bs) return -2;
    const uint8_t *p = fp->rbuf + fp->rpos;
    p += 4;
    bam1_core_t *c = &b->core;
    c->tid = (int32_t)rd_u32(p);
    c->pos = (int32_t)rd_u32(p + 4);
    uint8_t lq = p[8];
    c->qual = p[9];
    c->bin = (uint16_t)(p[10] | (p[11] << 8));
    c->n_cigar = (uint32_t)(p[12] | (p[13] << 8));
    c->flag = (uint16_t)(p[14] | (p[15] << 8));
    c->l_qseq = (int32_t)rd_u32(p + 16);
    c->mtid = (int32_t)rd_u32(p + 20);
    c->mpos = (int32_t)rd_u32(p + 24);
    c->isize = (int32_t)rd_u32(p + 28);
    c->unused1 = 0;
    /* a corrupt or truncated record must end in an error, not in an out-of-bounds copy: the fixed part, the name, the CIGAR,
     * the packed sequence and the qualities all have to fit inside block_size */
    if (c->l_qseq < 0 || (uint64_t)32 + lq + 4ull * c->n_cigar + ((uint64_t)c->l_qseq + 1) / 2 + (uint64_t)c->l_qseq > bs) return -2;
    unsigned pad = (4 - (lq & 3)) & 3;
    if (lq + pad > 255) pad = 0;                               /* htslib's rule: no extra NULs when they would not fit l_qname */
    c->l_extranul = (uint8_t)pad;
    c->l_qname = (uint8_t)(lq + pad);
    size_t rest = bs - 32 - lq;
    if (bam_reserve(b, (size_t)c->l_qname + rest) < 0) return -3;
    memcpy(b->data, p + 32, lq);
    memset(b->data + lq, 0, pad);
    memcpy(b->data + c->l_qname, p + 32 + lq, rest);
    b->l_data = (int)(c->l_qname + rest);
    fp->rpos += 4 + bs;
    return (int)bs;
}

static int bam_write_rec(samFile *fp, const bam1_t *b) {
    const bam1_core_t *c = &b->core;
    unsigned lq = c->l_qname - c->l_extranul;
    uint32_t bs = 32 + (uint32_t)b->l_data - c->l_extranul;
    if (bb_reserve(&fp->wbuf, bs + 4) < 0) return -1;
    uint8_t *p = fp->wbuf.p + fp->wbuf.n;
    wr_u32(p, bs);
    wr_u32(p + 4, (uint32_t)c->tid);
    wr_u32(p + 8, (uint32_t)c->pos);
    p[12] = (uint8_t)lq; p[13] = c->qual;
    wr_u16(p + 14, c->bin);
    wr_u16(p + 16, (uint16_t)c->n_cigar);
    wr_u16(p + 18, c->flag);
    wr_u32(p + 20, (uint32_t)c->l_qseq);
    wr_u32(p + 24, (uint32_t)c->mtid);
    wr_u32(p + 28, (uint32_t)c->mpos);
    wr_u32(p + 32, (uint32_t)c->isize);
    memcpy(p + 36, b->data, lq);
    memcpy(p + 36 + lq, b->data + c->l_qname, (size_t)b->l_data - c->l_qname);
    fp->wbuf.n += bs + 4;
    return (int)bs;
}

/* ---- SAM text --------------------------------------------------------------- */
static int aux_put_int(bbuf *o, long long v) {
    uint8_t t[5];
    if (v < 0) {
        if (v >= -128) { t[0] = 'c'; t[1] = (uint8_t)(int8_t)v; return bb_put(o, t, 2); }
        if (v >= -32768) { t[0] = 's'; wr_u16(t + 1, (uint16_t)(int16_t)v); return bb_put(o, t, 3); }
        t[0] = 'i'; wr_u32(t + 1, (uint32_t)(int32_t)v); return bb_put(o, t, 5);
    }
    if (v <= 255) { t[0] = 'C'; t[1] = (uint8_t)v; return bb_put(o, t, 2); }
    if (v <= 65535) { t[0] = 'S'; wr_u16(t + 1, (uint16_t)v); return bb_put(o, t, 3); }
    t[0] = 'I'; wr_u32(t + 1, (uint32_t)v); return bb_put(o, t, 5);
}

static char *next_field(char **cursor) {
    char *s = *cursor;
    if (!s) return NULL;
    char *t = strchr(s, '\t');
    if (t) { *t = 0; *cursor = t + 1; } else *cursor = NULL;
    return s;
}

int sam_parse_line(char *line, size_t len, bam_hdr_t *h, bam1_t *b) {
    (void)len;
    char *cur = line;
    char *qname = next_field(&cur), *flag = next_field(&cur), *rname = next_field(&cur),
         *pos = next_field(&cur), *mapq = next_field(&cur), *cigar = next_field(&cur),
         *rnext = next_field(&cur), *pnext = next_field(&cur), *tlen = next_field(&cur),
         *seq = next_field(&cur), *qual = next_field(&cur);
    if (!qual) return -2;
    bam1_core_t *c = &b->core;
    memset(c, 0, sizeof(*c));
    size_t lq = strlen(qname) + 1;
    if (lq > 255) return -2;
    unsigned pad = (4 - (lq & 3)) & 3;
    if (lq + pad > 255) pad = 0;                               /* htslib's rule: l_qname is 8 bit, so no extra NULs when they would not fit */
    c->l_qname = (uint8_t)(lq + pad); c->l_extranul = (uint8_t)pad;
    c->flag = (uint16_t)strtol(flag, NULL, 0);
    c->tid = strcmp(rname, "*") ? bam_name2id(h, rname) : -1;
    if (strcmp(rname, "*") && c->tid < 0) { fprintf(stderr, "hts_lite: unknown reference '%s'\n", rname); return -2; }
    c->pos = (int32_t)strtol(pos, NULL, 10) - 1;
    c->qual = (uint8_t)strtol(mapq, NULL, 10);
    c->mtid = !strcmp(rnext, "*") ? -1 : (!strcmp(rnext, "=") ? c->tid : bam_name2id(h, rnext));
    c->mpos = (int32_t)strtol(pnext, NULL, 10) - 1;
    c->isize = (int32_t)strtol(tlen, NULL, 10);
    /* count cigar ops */
    uint32_t nc = 0;
    if (strcmp(cigar, "*")) for (char *p = cigar; *p; p++) if (!isdigit((unsigned char)*p)) nc++;
    c->n_cigar = nc;
    size_t ls = strcmp(seq, "*") ? strlen(seq) : 0;
    c->l_qseq = (int32_t)ls;
    size_t need = c->l_qname + 4u * nc + ((ls + 1) >> 1) + ls;
    if (bam_reserve(b, need + 64) < 0) return -3;
    memcpy(b->data, qname, lq); memset(b->data + lq, 0, pad);
    uint32_t *cg = (uint32_t *)(b->data + c->l_qname);
    if (nc) {
        char *p = cigar; uint32_t k = 0;
        while (*p) {
            uint32_t l = (uint32_t)strtoul(p, &p, 10);
            const char *o = strchr(BAM_CIGAR_STR, *p);
            if (!o || !*p) return -2;
            cg[k++] = bam_cigar_gen(l, (uint32_t)(o - BAM_CIGAR_STR));
            p++;
        }
    }
    uint8_t *s = b->data + c->l_qname + 4u * nc;
    memset(s, 0, (ls + 1) >> 1);
    for (size_t i = 0; i < ls; i++)
        s[i >> 1] |= (uint8_t)(seq_nt16_table[(unsigned char)seq[i]] << ((~i & 1) << 2));
    uint8_t *q = s + ((ls + 1) >> 1);
    if (!strcmp(qual, "*")) memset(q, 0xff, ls);
    else {
        if (strlen(qual) != ls) return -2;
        for (size_t i = 0; i < ls; i++) q[i] = (uint8_t)(qual[i] - 33);
    }
    b->l_data = (int)need;
    /* aux */
    bbuf a = {0};
    char *f;
    while ((f = next_field(&cur))) {
        size_t fl = strlen(f);
        if (fl < 5 || f[2] != ':' || f[4] != ':') { free(a.p); return -2; }
        bb_put(&a, f, 2);
        char ty = f[3]; char *v = f + 5;
        switch (ty) {
        case 'A': bb_putc(&a, 'A'); bb_putc(&a, *v); break;
        case 'i': aux_put_int(&a, strtoll(v, NULL, 10)); break;
        case 'f': { float x = strtof(v, NULL); bb_putc(&a, 'f'); bb_put(&a, &x, 4); break; }
        case 'Z': case 'H': bb_putc(&a, ty); bb_put(&a, v, strlen(v) + 1); break;
        case 'B': {
            char st = *v++; uint32_t n = 0;
            for (char *p = v; *p; p++) if (*p == ',') n++;
            uint8_t t[4];
            bb_putc(&a, 'B'); bb_putc(&a, st); wr_u32(t, n); bb_put(&a, t, 4);
            char *p = v;
            for (uint32_t i = 0; i < n; i++) {
                p++;  /* skip ',' */
                if (st == 'f') { float x = strtof(p, &p); bb_put(&a, &x, 4); }
                else {
                    long long x = strtoll(p, &p, 10);
                    if (st == 'c' || st == 'C') bb_putc(&a, (int)(x & 0xff));
                    else if (st == 's' || st == 'S') { wr_u16(t, (uint16_t)x); bb_put(&a, t, 2); }
                    else { wr_u32(t, (uint32_t)x); bb_put(&a, t, 4); }
                }
            }
            break;
        }
        default: free(a.p); return -2;
        }
    }
    if (a.n) {
        if (bam_reserve(b, need + a.n + 64) < 0) { free(a.p); return -3; }
        memcpy(b->data + need, a.p, a.n);
        b->l_data += (int)a.n;
    }
    free(a.p);
    c->bin = (uint16_t)reg2bin(c->pos, bam_endpos(b));
    return 0;
}

static const uint8_t *fmt_aux(bbuf *o, const uint8_t *s, const uint8_t *end) {
    if (end - s < 3) return NULL;
    bb_putc(o, '\t'); bb_put(o, s, 2); bb_putc(o, ':');
    uint8_t ty = s[2]; s += 3;
    char tmp[64];
    switch (ty) {
    case 'A': bb_puts(o, "A:"); bb_putc(o, *s++); break;
    case 'c': bb_puts(o, "i:"); bb_putint(o, (int8_t)*s); s += 1; break;
    case 'C': bb_puts(o, "i:"); bb_putint(o, *s); s += 1; break;
    case 's': bb_puts(o, "i:"); bb_putint(o, (int16_t)(s[0] | (s[1] << 8))); s += 2; break;
    case 'S': bb_puts(o, "i:"); bb_putint(o, (uint16_t)(s[0] | (s[1] << 8))); s += 2; break;
    case 'i': bb_puts(o, "i:"); bb_putint(o, (int32_t)rd_u32(s)); s += 4; break;
    case 'I': bb_puts(o, "i:"); bb_putint(o, rd_u32(s)); s += 4; break;
    case 'f': { float x; memcpy(&x, s, 4); s += 4; snprintf(tmp, sizeof tmp, "f:%g", x); bb_puts(o, tmp); break; }
    case 'd': { double x; memcpy(&x, s, 8); s += 8; snprintf(tmp, sizeof tmp, "d:%g", x); bb_puts(o, tmp); break; }
    case 'Z': case 'H': bb_putc(o, ty); bb_putc(o, ':'); while (s < end && *s) bb_putc(o, *s++); s++; break;
    case 'B': {
        uint8_t st = *s++; uint32_t n = rd_u32(s); s += 4;
        bb_puts(o, "B:"); bb_putc(o, st);
        for (uint32_t i = 0; i < n; i++) {
            bb_putc(o, ',');
            switch (st) {
            case 'c': bb_putint(o, (int8_t)*s); s += 1; break;
            case 'C': bb_putint(o, *s); s += 1; break;
            case 's': bb_putint(o, (int16_t)(s[0] | (s[1] << 8))); s += 2; break;
            case 'S': bb_putint(o, (uint16_t)(s[0] | (s[1] << 8))); s += 2; break;
            case 'i': bb_putint(o, (int32_t)rd_u32(s)); s += 4; break;
            case 'I': bb_putint(o, rd_u32(s)); s += 4; break;
            case 'f': { float x; memcpy(&x, s, 4); s += 4; snprintf(tmp, sizeof tmp, "%g", x); bb_puts(o, tmp); break; }
            default: return NULL;
            }
        }
        break;
    }
    default: return NULL;
    }
    return s;
}

static int sam_format_into(bbuf *o, const bam_hdr_t *h, const bam1_t *b) {
    const bam1_core_t *c = &b->core;
    bb_puts(o, bam_get_qname(b)); bb_putc(o, '\t');
    bb_putint(o, c->flag); bb_putc(o, '\t');
    if (c->tid >= 0 && c->tid < h->n_targets) bb_puts(o, h->target_name[c->tid]); else bb_putc(o, '*');
    bb_putc(o, '\t');
    bb_putint(o, (long long)c->pos + 1); bb_putc(o, '\t');
    bb_putint(o, c->qual); bb_putc(o, '\t');
    if (c->n_cigar) {
        const uint32_t *cg = bam_get_cigar(b);
        for (uint32_t k = 0; k < c->n_cigar; k++) { bb_putint(o, bam_cigar_oplen(cg[k])); bb_putc(o, BAM_CIGAR_STR[bam_cigar_op(cg[k])]); }
    } else bb_putc(o, '*');
    bb_putc(o, '\t');
    if (c->mtid < 0) bb_putc(o, '*');
    else if (c->mtid == c->tid) bb_putc(o, '=');
    else if (c->mtid < h->n_targets) bb_puts(o, h->target_name[c->mtid]);
    else bb_putc(o, '*');
    bb_putc(o, '\t');
    bb_putint(o, (long long)c->mpos + 1); bb_putc(o, '\t');
    bb_putint(o, c->isize); bb_putc(o, '\t');
    int ls = c->l_qseq;
    if (ls) {
        if (bb_reserve(o, (size_t)ls * 2 + 2) < 0) return -1;
        const uint8_t *s = bam_get_seq(b);
        uint8_t *w = o->p + o->n;
        for (int i = 0; i < ls; i++) w[i] = (uint8_t)seq_nt16_str[bam_seqi(s, i)];
        w[ls] = '\t';
        const uint8_t *q = bam_get_qual(b);
        if (q[0] == 0xff) { w[ls + 1] = '*'; o->n += (size_t)ls + 2; }
        else { for (int i = 0; i < ls; i++) w[ls + 1 + i] = (uint8_t)(q[i] + 33); o->n += (size_t)ls * 2 + 1; }
    } else bb_puts(o, "*\t*");
    const uint8_t *a = bam_get_aux(b), *end = b->data + b->l_data;
    while (a && a < end) a = fmt_aux(o, a, end);
    if (!a) return -1;
    bb_putc(o, '\n');
    return 0;
}

size_t sam_format_line(const bam_hdr_t *h, const bam1_t *b, char **buf, size_t *cap) {
    bbuf o = { (uint8_t *)*buf, 0, *cap };
    sam_format_into(&o, h, b);
    bb_putc(&o, 0);
    *buf = (char *)o.p; *cap = o.cap;
    return o.n - 1;
}

int sam_read1(samFile *fp, bam_hdr_t *h, bam1_t *b) {
    span_touch();
    if (fp->format == bam) return bam_read_rec(fp, b);
    for (;;) {
        if (fp->rpos >= fp->rlen && rneed(fp, 1) < 1) return -1;
        uint8_t *eol;
        while (!(eol = (uint8_t *)memchr(fp->rbuf + fp->rpos, '\n', fp->rlen - fp->rpos))) {
            size_t have = fp->rlen - fp->rpos;
            if (rneed(fp, have + 1) <= have) break;              /* last line without a newline */
        }
        uint8_t *s = fp->rbuf + fp->rpos;
        size_t l = eol ? (size_t)(eol - s) : fp->rlen - fp->rpos;
        fp->rpos += l + (eol ? 1 : 0);
        if (l && s[l - 1] == '\r') l--;
        if (!l) continue;
        if (l + 1 > fp->line_cap) { fp->line_cap = (l + 1) * 2; fp->line = (char *)realloc(fp->line, fp->line_cap); }
        memcpy(fp->line, s, l); fp->line[l] = 0;
        int r = sam_parse_line(fp->line, l, h ? h : fp->hdr, b);
        if (r < 0) { fprintf(stderr, "hts_lite: malformed SAM line\n"); return -2; }
        return (int)l;
    }
}

int sam_write1(samFile *fp, const bam_hdr_t *h, const bam1_t *b) {
    span_touch();
    int r;
    if (fp->format == bam) r = bam_write_rec(fp, b);
    else r = sam_format_into(&fp->wbuf, h, b);
    if (r < 0) return -1;
    if (wflush(fp, 0) < 0) return -1;
    return 1;
}

/* ---- region iteration (no index: sequential scan with overlap filter) --------- */
hts_idx_t *sam_index_load(samFile *fp, const char *fn) { (void)fp; (void)fn; return (hts_idx_t *)calloc(1, sizeof(hts_idx_t)); }
void hts_idx_destroy(hts_idx_t *idx) { free(idx); }
void hts_itr_destroy(hts_itr_t *itr) { free(itr); }

hts_itr_t *sam_itr_querys(const hts_idx_t *idx, bam_hdr_t *hdr, const char *region) {
    (void)idx;
    hts_itr_t *it = (hts_itr_t *)calloc(1, sizeof(*it));
    if (!it) return NULL;
    it->beg = 0; it->end = INT32_MAX;
    char *name = strdup(region);
    it->tid = bam_name2id(hdr, name);
    if (it->tid < 0) {
        char *colon = strrchr(name, ':');
        if (colon) {
            *colon = 0;
            it->tid = bam_name2id(hdr, name);
            char *p = colon + 1, *q = p, *w = p;
            for (; *q; q++) if (*q != ',') *w++ = *q;   /* strip thousands separators */
            *w = 0;
            long b = strtol(p, &p, 10);
            it->beg = b > 0 ? (int)(b - 1) : 0;
            if (*p == '-') { long e = strtol(p + 1, NULL, 10); if (e > 0) it->end = (int)e; }
            else if (*p == 0 && b > 0) it->end = INT32_MAX;
        }
    }
    free(name);
    if (it->tid < 0) { free(it); return NULL; }
    return it;
}

int sam_itr_next(samFile *fp, hts_itr_t *itr, bam1_t *b) {
    if (itr->finished) return -1;
    for (;;) {
        int r = sam_read1(fp, NULL, b);   /* header only needed for SAM text */
        if (r < 0) { itr->finished = 1; return r; }
        if (b->core.tid < itr->tid && b->core.tid >= 0) continue;
        if (b->core.tid != itr->tid || b->core.pos >= itr->end) { itr->finished = 1; return -1; }
        if (bam_endpos(b) <= itr->beg) continue;
        return r;
    }
}
