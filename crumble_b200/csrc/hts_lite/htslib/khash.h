/*
 * hts_lite: tiny integer hash-set with the khash macro surface that crumble uses
 * (KHASH_SET_INIT_INT, khash_t, kh_init/put/get/end/destroy; reference call sites
 * snp_score.c:182-183,999,1016,2033,2038,2645).  Open addressing, linear probing.
 * Written for this repository; not derived from klib.
 */
#ifndef HTS_LITE_KHASH_H
#define HTS_LITE_KHASH_H
#include <stdint.h>
#include <stdlib.h>

typedef uint32_t khint_t;
typedef khint_t khiter_t;

typedef struct {
    khint_t n_buckets, size;
    uint32_t *keys;
    uint8_t  *used;
} hts_lite_intset;

static inline hts_lite_intset *hts_lite_intset_init(void) {
    return (hts_lite_intset *)calloc(1, sizeof(hts_lite_intset));
}
static inline void hts_lite_intset_destroy(hts_lite_intset *h) {
    if (h) { free(h->keys); free(h->used); free(h); }
}
static inline khint_t hts_lite_intset_get(const hts_lite_intset *h, uint32_t key) {
    if (!h->n_buckets) return 0;
    khint_t mask = h->n_buckets - 1, i = (key * 2654435761u) & mask, n = 0;
    while (h->used[i] && n++ < h->n_buckets) {
        if (h->keys[i] == key) return i;
        i = (i + 1) & mask;
    }
    return h->n_buckets;
}
static inline khint_t hts_lite_intset_put(hts_lite_intset *h, uint32_t key, int *ret) {
    if ((h->size + 1) * 2 > h->n_buckets) {
        khint_t nb = h->n_buckets ? h->n_buckets * 2 : 16;
        uint32_t *ok = h->keys; uint8_t *ou = h->used; khint_t on = h->n_buckets;
        h->keys = (uint32_t *)calloc(nb, sizeof(uint32_t));
        h->used = (uint8_t *)calloc(nb, 1);
        h->n_buckets = nb; h->size = 0;
        for (khint_t j = 0; j < on; j++) if (ou[j]) { int r; hts_lite_intset_put(h, ok[j], &r); }
        free(ok); free(ou);
    }
    khint_t mask = h->n_buckets - 1, i = (key * 2654435761u) & mask;
    while (h->used[i]) {
        if (h->keys[i] == key) { if (ret) *ret = 0; return i; }
        i = (i + 1) & mask;
    }
    h->used[i] = 1; h->keys[i] = key; h->size++;
    if (ret) *ret = 1;
    return i;
}

#define KHASH_SET_INIT_INT(name) typedef hts_lite_intset kh_##name##_t;
#define khash_t(name) kh_##name##_t
#define kh_init(name) hts_lite_intset_init()
#define kh_destroy(name, h) hts_lite_intset_destroy(h)
#define kh_get(name, h, k) hts_lite_intset_get(h, (uint32_t)(k))
#define kh_put(name, h, k, r) hts_lite_intset_put(h, (uint32_t)(k), r)
#define kh_end(h) ((h)->n_buckets)
#define kh_size(h) ((h)->size)
#endif
